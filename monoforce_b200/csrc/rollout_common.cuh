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
    long long map_stride;        // elements between two trajectories' maps, 0 = one shared map
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
    // outputs
    T* Xs;                       // (B,T,3)
    T* Xds;                      // (B,T,3)
    T* Rs;                       // (B,T,3,3)
    T* Oms;                      // (B,T,3)
    T* Fs;                       // (B,T,N,3) or nullptr
    T* Ff;                       // (B,T,N,3) or nullptr
    T* x0z;                      // (B,) snapped start height (dphysics.py:567-571)
    T* cost;                     // (B,) or nullptr: std_t(std_p |F_spring|), monoforce_node.py:91
};

// ------------------------------------------------------------------------------------------
// scalar math: approx SFU ops (<= 2 ulp) for float, libm for double
// ------------------------------------------------------------------------------------------
template <typename T> struct Mth;

template <> struct Mth<float> {
    static __device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    static __device__ __forceinline__ float rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
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
};

template <> struct Mth<double> {
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / ::sqrt(x); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double contact(double dh) { return 1.0 / (1.0 + ::exp(10.0 * dh)); }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double fmin_(double a, double b) { return ::fmin(a, b); }
    static __device__ __forceinline__ double fmax_(double a, double b) { return ::fmax(a, b); }
    static __device__ __forceinline__ double to_cells(double x, double d_max, double res, double) { return (x + d_max) / res; }
    static __device__ __forceinline__ double sqrt_rn(double x) { return ::sqrt(x); }
};

template <typename T>
__device__ __forceinline__ T clampT(T v, T lim) { return Mth<T>::fmax_(Mth<T>::fmin_(v, lim), -lim); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// grid sampling (dphysics.py:385-455)
// ------------------------------------------------------------------------------------------
struct Cell {
    int k00, k10, k01, k11;      // flat indices: centre, x+1 ("front"), y+1 ("left"), both
};

// Flat indices of the four neighbours with the reference's flat clamp.  Fast path when the
// cell is interior (no clamp can trigger); otherwise 64-bit arithmetic like torch's int64.
template <typename T>
__device__ __forceinline__ Cell locate(T gx, T gy, int H, int W, T& fx, T& fy) {
    const int ix = (int)gx;                      // truncation toward zero == .long()
    const int iy = (int)gy;
    fx = gx - (T)ix;
    fy = gy - (T)iy;
    Cell c;
    c.k00 = iy + H * ix;
    c.k10 = c.k00 + H;
    c.k01 = c.k00 + 1;
    c.k11 = c.k10 + 1;
    const bool interior = ((unsigned)ix < (unsigned)(H - 1)) && ((unsigned)iy < (unsigned)(W - 1));
    if (!interior) {
        const long long lx = (long long)gx, ly = (long long)gy;
        fx = gx - (T)lx;
        fy = gy - (T)ly;
        const long long last = (long long)H * W - 1;
        auto cl = [last](long long v) { return (int)(v < 0 ? 0 : (v > last ? last : v)); };
        c.k00 = cl(ly + (long long)H * lx);
        c.k10 = cl(ly + (long long)H * (lx + 1));
        c.k01 = cl((ly + 1) + (long long)H * lx);
        c.k11 = cl((ly + 1) + (long long)H * (lx + 1));
    }
    return c;
}

template <typename T>
__device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }

// value = (1-fx)(1-fy) v00 + (1-fx) fy v10 + fx (1-fy) v01 + fx fy v11   (weights as in the
// reference: the x+1 neighbour carries the y weight and vice versa, dphysics.py:442-445)
template <typename T>
__device__ __forceinline__ T blend(T fx, T fy, T v00, T v10, T v01, T v11) {
    const T gx = (T)1 - fx, gy = (T)1 - fy;
    return gx * gy * v00 + gx * fy * v10 + fx * gy * v01 + fx * fy * v11;
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

// R <- R (I + K sin(th dt) + K K (1 - cos(th dt))),  K = [w]x / max(|w|, 1e-6)   (dphysics.py:290-324)
template <typename T>
__device__ __forceinline__ void rodrigues_right(T* R, const T* w, T dt) {
    const T th = Mth<T>::sqrt_rn(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const T inv = (T)1 / Mth<T>::fmax_(th, (T)1e-6);
    const T k0 = w[0] * inv, k1 = w[1] * inv, k2 = w[2] * inv;
    T sn, cs;
    Mth<T>::sincos(th * dt, &sn, &cs);
    const T c1 = (T)1 - cs;
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

// Body points staged once per block: slot = j*32 + lane == point index.  Padded slots (only in
// the last j) hold the body origin; the kernels zero their soft-contact weight explicitly.
template <typename T>
struct PointTable {
    T px[kMaxPointsPerLane * 32];
    T py[kMaxPointsPerLane * 32];
    T pz[kMaxPointsPerLane * 32];
    T side[kMaxPointsPerLane * 32];   // 0: not driven, -half_Ly: left track, +half_Ly: right track
    T driven[kMaxPointsPerLane * 32]; // 1 if the point belongs to a driving part else 0
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
            tab.driven[p] = q >= 0 ? (T)1 : (T)0;
            tab.side[p] = q < 0 ? (T)0 : ((q & 1) ? a.half_Ly : -a.half_Ly);
        } else {
            tab.px[p] = (T)0; tab.py[p] = (T)0; tab.pz[p] = (T)0;
            tab.driven[p] = (T)0; tab.side[p] = (T)0;
        }
    }
}

}  // namespace mfb
