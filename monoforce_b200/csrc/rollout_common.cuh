// Shared device code of the DPhysics rollout kernels (forward + adjoint), sm_100a.
//
// Algorithm: SURVEY.md appendix A, i.e. the reference's step loop
//   monoforce/src/monoforce/models/traj_predictor/dphysics.py:172-288 (forces + integration),
//   :385-455 (grid sampling), :467-497 (record order), :530-594 (pre/post-processing).
// Mapping: one warp per trajectory, contact point p -> (lane = p % 32, slot j = p / 32),
// PPL = ceil(N/32) points per lane, rigid-body state replicated in every lane's registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mfb {

constexpr int kMaxPointsPerLane = 8;          // N <= 256 contact points
constexpr unsigned kFull = 0xffffffffu;

enum Variant : int { kStepLoop = 0, kOdeintEuler = 1 };

// Plain-old-data argument block shared by host and device (typed on the scalar).
template <typename T>
struct RolloutArgs {
    // geometry / problem size
    int B, nT, N, H, W, n_tracks;
    long long map_stride;        // elements between two consecutive maps, 0 = one shared map
    int map_group;               // consecutive trajectories that read the same map (>= 1): trajectory b uses map b / map_group
    // constants (dphys_config.py:77-153), already converted to T the way torch converts python scalars
    T mass, inv_mass, mg, stiffness, damping, res, inv_res, d_max, dt, omega_max, half_Ly, delta_h;
    T Iinv[9];                   // inverse inertia tensor, row-major (dphysics.py:152-153)
    // inputs
    const T* z;                  // (B|1, H, W)
    const T* mu;                 // (B|1, H, W)
    const T* controls;           // (B, T, 2)
    const T* x0;                 // (B,3)
    const T* xd0;                // (B,3)
    const T* R0;                 // (B,3,3)
    const T* om0;                // (B,3)
    const T* pts;                // (N,3) body-frame contact points
    const int* part;             // (N,) driving part id or -1
    const T* ts;                 // (T,) solver time grid (odeint variant only)
    const T* joint_angles;       // (B, T, 4) flipper angles or nullptr (static geometry)      dphysics.py:326-358
    T joint_pos[12];             // pivot of each driving part, row-major (4,3)               dphys_config.py:99-104
    const T* cells;              // (n_maps, H, W, 12) packed cell table built by build_cell_table (workspace)
    long long cell_stride;       // elements between two trajectories' tables, 0 = shared
    // outputs
    T* Xs;                       // (B,T,3)
    T* Xds;                      // (B,T,3)
    T* Rs;                       // (B,T,3,3)
    T* Oms;                      // (B,T,3)
    T* Fs;                       // (B,T,N,3) or nullptr
    T* Ff;                       // (B,T,N,3) or nullptr
    T* x0z;                      // (B,) snapped start height (dphysics.py:567-571)
    T* cost;                     // (B,) or nullptr: std_t(std_p |F_spring|), monoforce_node.py:91
    T* Csum;                     // (B,T) or nullptr: soft-contact normaliser sum_p c_p of every step (dphysics.py:231), the
                                 // tape the single-sweep adjoint reads back (written by the forward, read by the backward)
};

// ------------------------------------------------------------------------------------------
// scalar math: approx SFU ops (<= 2 ulp) for float, libm for double
// ------------------------------------------------------------------------------------------
template <typename T> struct Mth;

template <> struct Mth<float> {
    static __device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    static __device__ __forceinline__ float rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    // reciprocal for warp-uniform scalars: SFU seed + one Newton step (<= 1 ulp), no IEEE slow path
    static __device__ __forceinline__ float inv(float x) { const float y = rcp(x); return fmaf(fmaf(-x, y, 1.0f), y, y); }
    static __device__ __forceinline__ float sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    // sigmoid(-10 dh) = 1 / (1 + exp(10 dh))
    static __device__ __forceinline__ float contact(float dh) {
        float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(dh * 14.426950408889634f));
        return rcp(1.0f + e);
    }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float fmin_(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float fmax_(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float to_cells(float x, float d_max, float res, float inv_res) { return (x + d_max) * inv_res; }
    static __device__ __forceinline__ float sqrt_rn(float x) { return sqrtf(x); }
    // (|v|, 1 / max(|v|, 1e-6)) from |v|^2 with one rsqrt (2 ulp)
    static __device__ __forceinline__ void norm_and_inv(float sq, float* nrm, float* inv) {
        const float r = rsqrt(fmaxf(sq, 1e-12f));
        *nrm = sq * r;
        *inv = r;
    }
};

template <> struct Mth<double> {
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / ::sqrt(x); }
    static __device__ __forceinline__ double inv(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double contact(double dh) { return 1.0 / (1.0 + ::exp(10.0 * dh)); }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double fmin_(double a, double b) { return ::fmin(a, b); }
    static __device__ __forceinline__ double fmax_(double a, double b) { return ::fmax(a, b); }
    static __device__ __forceinline__ double to_cells(double x, double d_max, double res, double) { return (x + d_max) / res; }
    static __device__ __forceinline__ double sqrt_rn(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ void norm_and_inv(double sq, double* nrm, double* inv) {
        *nrm = ::sqrt(sq);
        *inv = 1.0 / ::fmax(*nrm, 1e-6);
    }
};

template <typename T>
__device__ __forceinline__ T clampT(T v, T lim) { return Mth<T>::fmax_(Mth<T>::fmin_(v, lim), -lim); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Sums v[0..8) over the warp; every lane ends with all eight totals.  Reduce-scatter (each butterfly stage
// halves the number of live values per lane) followed by an all-gather: 17 shuffles + 14 selects + 9 adds
// instead of 40 shuffles + 40 adds for eight independent butterflies.
template <typename T>
__device__ __forceinline__ void warp_sum8(T* v, int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    T w[4], u[2], t;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T keep = b4 ? v[i + 4] : v[i], send = b4 ? v[i] : v[i + 4];
        w[i] = keep + __shfl_xor_sync(kFull, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const T keep = b3 ? w[i + 2] : w[i], send = b3 ? w[i] : w[i + 2];
        u[i] = keep + __shfl_xor_sync(kFull, send, 8);
    }
    {
        const T keep = b2 ? u[1] : u[0], send = b2 ? u[0] : u[1];
        t = keep + __shfl_xor_sync(kFull, send, 4);
    }
    t += __shfl_xor_sync(kFull, t, 2);
    t += __shfl_xor_sync(kFull, t, 1);
    // lane with (bit4, bit3, bit2) = (k>>2, k>>1, k) & 1 holds total k
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __shfl_sync(kFull, t, ((k >> 2) & 1) * 16 + ((k >> 1) & 1) * 8 + (k & 1) * 4);
}

// lane id through a volatile asm so that the compiler keeps it in a register instead of re-deriving it
// from %tid.x (S2R has a long latency) all over the unrolled step loop
__device__ __forceinline__ int lane_id() {
    int l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

__device__ __forceinline__ int warp_id_pinned() {
    int t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    return t >> 5;
}

// ------------------------------------------------------------------------------------------
// grid sampling (dphysics.py:385-455) through a packed per-cell table
// ------------------------------------------------------------------------------------------
// The reference samples  v(fx,fy) = (1-fx)(1-fy) v00 + (1-fx) fy v10 + fx (1-fy) v01 + fx fy v11
// with v10 the x+1 ("front") and v01 the y+1 ("left") neighbour (weights swapped w.r.t. true
// bilinear, dphysics.py:442-445) and a normal n = normalize(-(v10-v00)/res, -(v01-v00)/res, 1)
// that is constant per cell.  Per cell (ix, iy) we therefore pre-compute once per launch
//   c0 = v00, cy = v10 - v00, cx = v01 - v00, cxy = v00 - v10 - v01 + v11   (height and friction)
//   n  = the cell normal
// so that  v = c0 + fy cy + fx (cx + fy cxy)  is 3 FMAs and one contact point costs three
// 16-byte loads from one 48-byte record instead of eight scattered 4-byte gathers.
// The neighbour indices use the reference's flat clamp (no per-axis clamp, :432-435).
constexpr int kCellRec = 12;     // scalars per cell record: c0 cy cx cxy | n0 n1 n2 pad | m0 my mx mxy
constexpr int kCellStride = 12;  // records are packed: padding them to 16 scalars (64 B, never straddling a 128-byte line) measured
                                 // +11 % on the forward kernel (3 MB -> 4 MB table, fewer neighbouring cells per L1 line)

struct Corners {
    int k00, k10, k01, k11;      // flat indices: centre, x+1 ("front"), y+1 ("left"), both
};

__device__ __forceinline__ Corners flat_corners(long long lx, long long ly, int H, int W) {
    const long long last = (long long)H * W - 1;
    auto cl = [last](long long v) { return (int)(v < 0 ? 0 : (v > last ? last : v)); };
    Corners c;
    c.k00 = cl(ly + (long long)H * lx);
    c.k10 = cl(ly + (long long)H * (lx + 1));
    c.k01 = cl((ly + 1) + (long long)H * lx);
    c.k11 = cl((ly + 1) + (long long)H * (lx + 1));
    return c;
}

// on-map cells (0 <= ix < H, 0 <= iy < W, H == W): only the upper clamp can trigger
__device__ __forceinline__ Corners on_map_corners(int cell, int H, int W) {
    const int last = H * W - 1;
    Corners c;
    c.k00 = cell;
    c.k10 = min(cell + H, last);
    c.k01 = min(cell + 1, last);
    c.k11 = min(cell + H + 1, last);
    return c;
}

template <typename T>
__device__ __forceinline__ void make_cell_record(const T* __restrict__ z, const T* __restrict__ mu, const Corners& c,
                                                 T inv_res, T* rec) {
    const T z00 = z[c.k00], z10 = z[c.k10], z01 = z[c.k01], z11 = z[c.k11];
    const T m00 = mu[c.k00], m10 = mu[c.k10], m01 = mu[c.k01], m11 = mu[c.k11];
    rec[0] = z00; rec[1] = z10 - z00; rec[2] = z01 - z00; rec[3] = (z00 - z10) - (z01 - z11);
    const T ax = (z00 - z10) * inv_res, ay = (z00 - z01) * inv_res;
    const T q = (T)1 / Mth<T>::sqrt_rn(ax * ax + ay * ay + (T)1);
    rec[4] = ax * q; rec[5] = ay * q; rec[6] = q; rec[7] = (T)0;
    rec[8] = m00; rec[9] = m10 - m00; rec[10] = m01 - m00; rec[11] = (m00 - m10) - (m01 - m11);
}

// one thread per cell of every map
template <typename T>
__global__ void build_cell_table_kernel(const T* __restrict__ z, const T* __restrict__ mu, T* __restrict__ cells,
                                        int n_maps, int H, int W, long long map_stride, T inv_res) {
    const long long total = (long long)n_maps * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / ((long long)H * W));
        const int k = (int)(i - (long long)m * H * W);
        const int ix = k / W, iy = k - ix * W;
        const Corners c = flat_corners(ix, iy, H, W);
        T rec[kCellRec];
        make_cell_record(z + m * map_stride, mu + m * map_stride, c, inv_res, rec);
        T* out = cells + i * kCellStride;
#pragma unroll
        for (int j = 0; j < kCellRec; ++j) out[j] = rec[j];
    }
}

template <typename T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int n = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int n = 2; };

// 16-byte vector loads of one record (read-only path)
template <typename T>
__device__ __forceinline__ void load_cell_record(const T* __restrict__ p, T* rec) {
    using V = typename Vec16<T>::type;
    constexpr int n = Vec16<T>::n;
    const V* q = reinterpret_cast<const V*>(p);
#pragma unroll
    for (int i = 0; i < kCellRec / n; ++i) {
        const V v = __ldg(q + i);
        const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int j = 0; j < n; ++j) rec[i * n + j] = e[j];
    }
}

// points outside [0,H) x [0,W): the reference's clamped flat indices, straight from the raw maps (rare)
template <typename T>
__device__ __noinline__ void sample_off_map(const T* __restrict__ z, const T* __restrict__ mu, T gx, T gy, int H, int W,
                                            T inv_res, T* out /* [kCellRec + 2] */) {
    const long long lx = (long long)gx, ly = (long long)gy;
    out[kCellRec + 0] = gx - (T)lx;
    out[kCellRec + 1] = gy - (T)ly;
    const Corners c = flat_corners(lx, ly, H, W);
    make_cell_record(z, mu, c, inv_res, out);
}

// ------------------------------------------------------------------------------------------
// rigid-body state held in registers (replicated across the warp)
// ------------------------------------------------------------------------------------------
template <typename T>
struct Body {
    T x[3], v[3], R[9], w[3];
};

template <typename T>
__device__ __forceinline__ void load_body(Body<T>& s, const RolloutArgs<T>& a, int b) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        s.x[i] = a.x0[b * 3 + i];
        s.v[i] = a.xd0[b * 3 + i];
        s.w[i] = a.om0[b * 3 + i];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) s.R[i] = a.R0[b * 9 + i];
}

// sin(x) and 1 - cos(x).  One integration step turns the body by |w| dt, a small angle: below 0.5 rad
// an odd/even Taylor pair (error < 1e-10) is both cheaper than libm's sincos and free of the
// cancellation in 1 - cos(x); larger angles (tumbling states) use libm.
template <typename T>
__device__ __forceinline__ void sin_versin(T x, T* sn, T* vs) {
    if (sizeof(T) == 4 && x < (T)0.5) {       // x = |w| dt >= 0; double precision always takes libm
        const T x2 = x * x;
        *sn = x * ((T)1 + x2 * ((T)(-1.0 / 6) + x2 * ((T)(1.0 / 120) + x2 * ((T)(-1.0 / 5040) + x2 * (T)(1.0 / 362880)))));
        *vs = x2 * ((T)0.5 + x2 * ((T)(-1.0 / 24) + x2 * ((T)(1.0 / 720) + x2 * ((T)(-1.0 / 40320) + x2 * (T)(1.0 / 3628800)))));
    } else {
        T cs;
        Mth<T>::sincos(x, sn, &cs);
        *vs = (T)1 - cs;
    }
}

// R <- R (I + K sin(th dt) + K K (1 - cos(th dt))),  K = [w]x / max(|w|, 1e-6)   (dphysics.py:290-324)
template <typename T>
__device__ __forceinline__ void rodrigues_right(T* R, const T* w, T dt) {
    T th, inv;
    Mth<T>::norm_and_inv(w[0] * w[0] + w[1] * w[1] + w[2] * w[2], &th, &inv);
    const T k0 = w[0] * inv, k1 = w[1] * inv, k2 = w[2] * inv;
    T sn, c1;
    sin_versin(th * dt, &sn, &c1);
    const T kk = k0 * k0 + k1 * k1 + k2 * k2;
    // E = I + sn K + c1 (k k^T - |k|^2 I)
    T E[9];
    E[0] = (T)1 + c1 * (k0 * k0 - kk);  E[1] = -sn * k2 + c1 * k0 * k1;     E[2] = sn * k1 + c1 * k0 * k2;
    E[3] = sn * k2 + c1 * k0 * k1;      E[4] = (T)1 + c1 * (k1 * k1 - kk);  E[5] = -sn * k0 + c1 * k1 * k2;
    E[6] = -sn * k1 + c1 * k0 * k2;     E[7] = sn * k0 + c1 * k1 * k2;      E[8] = (T)1 + c1 * (k2 * k2 - kk);
    T Rn[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            Rn[r * 3 + c] = R[r * 3 + 0] * E[0 * 3 + c] + R[r * 3 + 1] * E[1 * 3 + c] + R[r * 3 + 2] * E[2 * 3 + c];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
}

// ------------------------------------------------------------------------------------------
// phase 1 of a step for one contact point (shared by the forward and the adjoint kernels)
// ------------------------------------------------------------------------------------------
template <typename T>
struct StepFrame {          // warp-uniform quantities of the current step
    T R[9], x[3], v[3], w[3];
    T ox, oy;               // (x + d_max) / res : grid offset of the body origin
    T hd[3];                // thrust direction: first column of R, normalised (dphysics.py:237)
    T uv, uw;               // controls (v, w)
};

template <typename T>
__device__ __forceinline__ void make_frame(StepFrame<T>& f, const Body<T>& s, T uv, T uw, const T d_max, const T res,
                                           const T inv_res) {
#pragma unroll
    for (int i = 0; i < 9; ++i) f.R[i] = s.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { f.x[i] = s.x[i]; f.v[i] = s.v[i]; f.w[i] = s.w[i]; }
    f.ox = Mth<T>::to_cells(s.x[0], d_max, res, inv_res);
    f.oy = Mth<T>::to_cells(s.x[1], d_max, res, inv_res);
    T nn, inv;
    Mth<T>::norm_and_inv(s.R[0] * s.R[0] + s.R[3] * s.R[3] + s.R[6] * s.R[6], &nn, &inv);
    f.hd[0] = s.R[0] * inv; f.hd[1] = s.R[3] * inv; f.hd[2] = s.R[6] * inv;
    f.uv = uv; f.uw = uw;
}

template <typename T>
struct PointEval {
    T r[3];                 // lever arm R p (== P - x)
    T V[3];                 // point velocity v + w x r
    T rec[kCellRec];        // cell record (coefficients, normal)
    T fx, fy;               // position inside the cell
    T dz_dfx;               // d zv / d fx = cx + fy cxy
    T mu, dh, cw, vn, sp;   // friction coef., height above terrain, soft-contact weight, normal speed, -(k dh + b vn)
    T tau;                  // commanded track speed at the point
    T e[3], d[3], dn;       // slip: e = tau hd - V, d = mu e, dn = d . n
    T sl[3];                // tangential slip d - dn n
    int cell;               // index into the cell table, -1 when the point is off the map
};

// `cells` / `zmap` / `fmap` are already offset to this trajectory's map.
// PATCH_OFF_MAP = false: branch-free fast path; an off-map point (o.cell == -1) is evaluated on cell 0's record and the
// caller must redo the step with PATCH_OFF_MAP = true if any lane reports one (rare).
// Raw-map pointers for the off-map path; `Maps` is any type with z() / mu() so that a caller can fetch them lazily
// (the single-sweep adjoint keeps them in shared memory: they are needed by one point in thousands).
template <typename T>
struct MapPtrs {
    const T* zp; const T* mp;
    __device__ __forceinline__ const T* z() const { return zp; }
    __device__ __forceinline__ const T* mu() const { return mp; }
};

template <typename T, bool PATCH_OFF_MAP, class Maps>
__device__ __forceinline__ void eval_point_maps(PointEval<T>& o, const StepFrame<T>& f, T px, T py, T pz, T drv, T side,
                                                bool valid, const T* __restrict__ cells, const Maps& maps, int H, int W,
                                                T inv_res, T stiffness, T damping) {
    // r = R p ; V = v + w x r                                                  dphysics.py:200-204
    o.r[0] = f.R[0] * px + f.R[1] * py + f.R[2] * pz;
    o.r[1] = f.R[3] * px + f.R[4] * py + f.R[5] * pz;
    o.r[2] = f.R[6] * px + f.R[7] * py + f.R[8] * pz;
    // (written as v + (a - b): the fused form fma(w1, r2, fma(-w2, r1, v0)) saves 3 instructions per point but measured
    // SLOWER in both kernels, +2 %: longer dependent chain on the same registers)
    o.V[0] = f.v[0] + (f.w[1] * o.r[2] - f.w[2] * o.r[1]);
    o.V[1] = f.v[1] + (f.w[2] * o.r[0] - f.w[0] * o.r[2]);
    o.V[2] = f.v[2] + (f.w[0] * o.r[1] - f.w[1] * o.r[0]);
    // grid coordinates of P = r + x and the cell they fall in                  dphysics.py:419-424
    const T gx = o.r[0] * inv_res + f.ox;
    const T gy = o.r[1] * inv_res + f.oy;
    const int ix = (int)gx, iy = (int)gy;        // truncation toward zero == .long()
    const bool on_map = ((unsigned)ix < (unsigned)H) && ((unsigned)iy < (unsigned)W);
    o.fx = gx - (T)ix;
    o.fy = gy - (T)iy;
    // the record load is unconditional (cell 0 stands in for off-map points) so that the loads of all
    // the lane's points can be in flight together; off-map points are patched afterwards (rare)
    const int cell = on_map ? ix * W + iy : 0;
    load_cell_record(cells + (long long)cell * kCellStride, o.rec);
    o.cell = on_map ? cell : -1;
    if (PATCH_OFF_MAP && !on_map) {
        // the record goes through a local buffer so that o.rec itself stays in registers
        T tmp[kCellRec + 2];
        sample_off_map(maps.z(), maps.mu(), gx, gy, H, W, inv_res, tmp);
#pragma unroll
        for (int k = 0; k < kCellRec; ++k) o.rec[k] = tmp[k];
        o.fx = tmp[kCellRec]; o.fy = tmp[kCellRec + 1];
    }
    // height, friction, normal                                                 dphysics.py:211-216
    o.dz_dfx = o.rec[2] + o.fy * o.rec[3];
    const T zv = o.rec[0] + o.fy * o.rec[1] + o.fx * o.dz_dfx;
    o.mu = o.rec[8] + o.fy * o.rec[9] + o.fx * (o.rec[10] + o.fy * o.rec[11]);
    const T n0 = o.rec[4], n1 = o.rec[5], n2 = o.rec[6];
    // soft contact, spring-damper magnitude                                    dphysics.py:220-232
    o.dh = (o.r[2] + f.x[2]) - zv;
    o.cw = valid ? Mth<T>::contact(o.dh) : (T)0;
    o.vn = o.V[0] * n0 + o.V[1] * n1 + o.V[2] * n2;
    o.sp = -(stiffness * o.dh + damping * o.vn);
    // tangential slip w.r.t. the commanded track speed                         dphysics.py:237-249
    o.tau = drv * f.uv + side * f.uw;
    o.e[0] = o.tau * f.hd[0] - o.V[0]; o.e[1] = o.tau * f.hd[1] - o.V[1]; o.e[2] = o.tau * f.hd[2] - o.V[2];
    o.d[0] = o.mu * o.e[0]; o.d[1] = o.mu * o.e[1]; o.d[2] = o.mu * o.e[2];
    o.dn = o.d[0] * n0 + o.d[1] * n1 + o.d[2] * n2;
    o.sl[0] = o.d[0] - o.dn * n0; o.sl[1] = o.d[1] - o.dn * n1; o.sl[2] = o.d[2] - o.dn * n2;
}

template <typename T, bool PATCH_OFF_MAP = true>
__device__ __forceinline__ void eval_point(PointEval<T>& o, const StepFrame<T>& f, T px, T py, T pz, T drv, T side,
                                           bool valid, const T* __restrict__ cells, const T* __restrict__ zmap,
                                           const T* __restrict__ fmap, int H, int W, T inv_res, T stiffness, T damping) {
    eval_point_maps<T, PATCH_OFF_MAP>(o, f, px, py, pz, drv, side, valid, cells, MapPtrs<T>{zmap, fmap}, H, W, inv_res,
                                      stiffness, damping);
}

// ------------------------------------------------------------------------------------------
// moving flippers (dphysics.py:326-358) and the per-step inverse inertia tensor (:196-197, :107-141)
// ------------------------------------------------------------------------------------------
// Lane i < 4 holds (cos, sin) of flipper angle i; a point of driving part q is rotated about the y axis
// through the part's pivot:  p' = Ry(angle_q) (p - pivot_q) + pivot_q.
template <typename T>
__device__ __forceinline__ void articulate_point(T& px, T& py, T& pz, int part, T my_cos, T my_sin, const T* joint_pos) {
    const int src = part < 0 ? 0 : part;
    const T c = __shfl_sync(kFull, my_cos, src), s = __shfl_sync(kFull, my_sin, src);
    if (part >= 0) {
        const T ox = joint_pos[part * 3 + 0], oz = joint_pos[part * 3 + 2];
        const T dx = px - ox, dz = pz - oz;
        px = c * dx + s * dz + ox;
        pz = -s * dx + c * dz + oz;
    }
}

// inverse of the symmetric point-mass inertia tensor from the six warp-reduced second moments
// m6 = (sum y^2+z^2, sum x^2+z^2, sum x^2+y^2, sum xy, sum xz, sum yz) * (mass / N)
template <typename T>
__device__ __forceinline__ void invert_inertia(const T* m6, T* inv9) {
    const T a = m6[0], d = m6[1], f = m6[2], b = -m6[3], c = -m6[4], e = -m6[5];     // [[a,b,c],[b,d,e],[c,e,f]]
    const T A = d * f - e * e, B = c * e - b * f, Cc = b * e - c * d;
    const T det = a * A + b * B + c * Cc;
    const T r = (T)1 / det;
    inv9[0] = A * r;            inv9[1] = B * r;            inv9[2] = Cc * r;
    inv9[3] = B * r;            inv9[4] = (a * f - c * c) * r; inv9[5] = (b * c - a * e) * r;
    inv9[6] = Cc * r;           inv9[7] = (b * c - a * e) * r; inv9[8] = (a * d - b * b) * r;
}

// 4 scalars moved as one (float) or two (double) 16-byte shared-memory accesses
template <typename T> struct Quad;                       // 4 scalars moved as one or two 16-byte shared accesses
template <> struct __align__(16) Quad<float> { float v[4]; };
template <> struct __align__(16) Quad<double> { double v[4]; };

__device__ __forceinline__ Quad<float> quad_load(const Quad<float>* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);       // one LDS.128
    Quad<float> q; q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w;
    return q;
}
__device__ __forceinline__ void quad_store(Quad<float>* p, const Quad<float>& q) {
    *reinterpret_cast<float4*>(p) = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]);      // one STS.128
}
__device__ __forceinline__ Quad<double> quad_load(const Quad<double>* p) {
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    Quad<double> q; q.v[0] = a.x; q.v[1] = a.y; q.v[2] = b.x; q.v[3] = b.y;
    return q;
}
__device__ __forceinline__ void quad_store(Quad<double>* p, const Quad<double>& q) {
    reinterpret_cast<double2*>(p)[0] = make_double2(q.v[0], q.v[1]);
    reinterpret_cast<double2*>(p)[1] = make_double2(q.v[2], q.v[3]);
}

// shared-memory quad load the compiler may not hoist out of the point loop
__device__ __forceinline__ Quad<float> quad_load_pinned(const Quad<float>* p) {
    Quad<float> q;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(q.v[0]), "=f"(q.v[1]), "=f"(q.v[2]), "=f"(q.v[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return q;
}
__device__ __forceinline__ Quad<double> quad_load_pinned(const Quad<double>* p) {
    Quad<double> q;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q.v[0]), "=d"(q.v[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(q.v[2]), "=d"(q.v[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return q;
}

// 16-byte shared-memory loads at [base + OFF] (32-bit shared address, compile-time offset); volatile: re-read at every use
template <int OFF>
__device__ __forceinline__ Quad<float> lds_quad(uint32_t base, float) {
    Quad<float> q;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                 : "=f"(q.v[0]), "=f"(q.v[1]), "=f"(q.v[2]), "=f"(q.v[3]) : "r"(base), "n"(OFF));
    return q;
}
template <int OFF>
__device__ __forceinline__ Quad<double> lds_quad(uint32_t base, double) {
    Quad<double> q;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(q.v[0]), "=d"(q.v[1]) : "r"(base), "n"(OFF));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(q.v[2]), "=d"(q.v[3]) : "r"(base), "n"(OFF + 16));
    return q;
}
template <int OFF>
__device__ __forceinline__ uint4 lds_u4(uint32_t base) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base), "n"(OFF));
    return v;
}

__device__ __forceinline__ unsigned long long u64_of(unsigned lo, unsigned hi) { return ((unsigned long long)hi << 32) | lo; }

// packed body-point table: one 16-byte load + one scalar load per point instead of five scalar loads
template <typename T>
struct SweepPoints {
    Quad<T> pp[kMaxPointsPerLane * 32];     // (px, py, pz, side)   side: 0 not driven, -+half_Ly left / right track
    T drv[kMaxPointsPerLane * 32];          // 1 if the point belongs to a driving part else 0
};

template <typename T>
__device__ __forceinline__ void fill_sweep_points(SweepPoints<T>& tab, const RolloutArgs<T>& a, int slots) {
    for (int p = threadIdx.x; p < slots; p += blockDim.x) {
        Quad<T> q; q.v[0] = q.v[1] = q.v[2] = q.v[3] = (T)0;
        T d = (T)0;
        if (p < a.N) {
            q.v[0] = a.pts[p * 3 + 0]; q.v[1] = a.pts[p * 3 + 1]; q.v[2] = a.pts[p * 3 + 2];
            const int part = a.part[p];
            d = part >= 0 ? (T)1 : (T)0;
            q.v[3] = part < 0 ? (T)0 : ((part & 1) ? a.half_Ly : -a.half_Ly);     // dphysics.py:75-104
        }
        quad_store(&tab.pp[p], q);
        tab.drv[p] = d;
    }
}

// Body points staged once per block: slot = j*32 + lane == point index.  Padded slots (only in
// the last j) hold the body origin; the kernels zero their soft-contact weight explicitly.
template <typename T>
struct PointTable {
    T px[kMaxPointsPerLane * 32];
    T py[kMaxPointsPerLane * 32];
    T pz[kMaxPointsPerLane * 32];
    T side[kMaxPointsPerLane * 32];   // 0: not driven, -half_Ly: left track, +half_Ly: right track
    T driven[kMaxPointsPerLane * 32]; // 1 if the point belongs to a driving part else 0
    int part[kMaxPointsPerLane * 32]; // driving part id or -1 (needed only when flippers move)
};

template <typename T>
__device__ __forceinline__ void fill_point_table(PointTable<T>& tab, const RolloutArgs<T>& a, int slots) {
    for (int p = threadIdx.x; p < slots; p += blockDim.x) {
        if (p < a.N) {
            tab.px[p] = a.pts[p * 3 + 0];
            tab.py[p] = a.pts[p * 3 + 1];
            tab.pz[p] = a.pts[p * 3 + 2];
            const int q = a.part[p];
            // 2 tracks: (left, right); 4 tracks: (FL, FR, RL, RR) -> odd index = right (dphysics.py:75-104)
            tab.part[p] = q;
            tab.driven[p] = q >= 0 ? (T)1 : (T)0;
            tab.side[p] = q < 0 ? (T)0 : ((q & 1) ? a.half_Ly : -a.half_Ly);
        } else {
            tab.px[p] = (T)0; tab.py[p] = (T)0; tab.pz[p] = (T)0;
            tab.driven[p] = (T)0; tab.side[p] = (T)0; tab.part[p] = -1;
        }
    }
}

}  // namespace mfb
