// K2s: single-sweep adjoint of the fused rollout (static geometry).  One warp walks one trajectory
// backwards in time and visits every contact point ONCE per step.
//
// The three-pass adjoint (rollout_bwd.cuh) evaluates phase 1 of a step twice because the adjoint of the
// soft-contact normaliser C = sum_p c_p is only known after every point's phase 2 has been reversed.
// Two observations remove the second evaluation:
//   (1) the forward kernel records C per (trajectory, step) (4 bytes per step, `contact_sum` tape), so
//       phase 2 of a point can be re-evaluated and reversed straight after its phase 1;
//   (2) everything downstream of  c_bar = (point's own term) + C_bar  is LINEAR in C_bar, which is the
//       same scalar for all points.  Each lane therefore accumulates a second set of sums ("kappa
//       channel": what one unit of C_bar contributes to x_bar and R_bar), and the state adjoint is
//       closed after the warp reduction as  base + C_bar * kappa-channel.  The map-gradient part of the
//       kappa channel, -C_bar kappa_p w_ij(p), is applied one visit later: every point parks
//       (kappa, fx, fy, cell) in shared memory and the next visit (step t-1) folds it into the point's
//       private corner accumulators before anything else touches them.
// The point loop is rolled (per-point state lives in shared memory, not in registers indexed by an unrolled
// loop counter), so the kernel body is several times smaller than the three-pass kernel's: no instruction-cache
// stalls.  Registers: the 36 running sums stay in registers; the carried state adjoint is parked in shared memory
// while the point loop runs (MFB_SWEEP_PARK) and the warp-uniform operands of the step are re-read from shared
// memory at every point (MFB_SWEEP_FRAME_SMEM) => 128 registers, 4 CTAs x 4 warps per SM.
//
// Map gradients: per sampling cell one 8-scalar record (d/dz and d/dfriction of the cell's four corners)
// accumulated with two aligned 16-byte vector reductions; finalize_map_grads_kernel adds the records
// into the caller's maps with the reference's corner indexing (dphysics.py:427-435).
#pragma once
#include "rollout_bwd.cuh"

namespace mfb {

#ifndef MFB_SWEEP_WARPS
#define MFB_SWEEP_WARPS 4
#endif
#ifndef MFB_SWEEP_PARK
#define MFB_SWEEP_PARK 1       // 1: the carried state adjoint waits in shared memory while the point loop runs (-18 registers)
#endif
#ifndef MFB_SWEEP_FRAME_SMEM
#define MFB_SWEEP_FRAME_SMEM 2  // warp-uniform per-step operands are re-read from shared memory at every point instead of living in
                                // registers.  1: v, w, controls, thrust direction, tq_bar, fs_bar, 1/C (-20 registers, +5 loads / point);
                                // 2: also R, x.z and the grid offsets (-12 more, +3 loads).  Measured at config 3 (B200): 0 -> 8.25 ms
                                // (168 regs, 12 warps/SM), 1 -> 7.9 ms and 2 -> 7.5 ms (128 regs, 16 warps/SM)
#endif
#ifndef MFB_SWEEP_PREFETCH
#define MFB_SWEEP_PREFETCH 0    // 1: prefetch.global.L1 of the next point's cell record (from the cell of its previous visit); measured
                                // slower (7.86 vs 7.60 ms): the extra shared-memory read + address math cost more than the L2 latency hidden
#endif
constexpr int kSweepWarps = MFB_SWEEP_WARPS;
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
#ifndef MFB_SWEEP_MINB
#define MFB_SWEEP_MINB 4        // 4 CTAs x 4 warps = 16 warps / SM at 128 registers
#endif
#ifndef MFB_SWEEP_UNROLL
#define MFB_SWEEP_UNROLL 1
#endif
constexpr int kSweepUnroll = MFB_SWEEP_UNROLL;
constexpr int kGradRec = 8;      // scalars per cell of the gradient scratch: z00 z10 z01 z11 | m00 m10 m01 m11

__device__ __forceinline__ float pack_cell(int c, float) { return __int_as_float(c); }
__device__ __forceinline__ double pack_cell(int c, double) { return __longlong_as_double((long long)c); }
__device__ __forceinline__ int unpack_cell(float v) { return __float_as_int(v); }
__device__ __forceinline__ int unpack_cell(double v) { return (int)__double_as_longlong(v); }

// off-map samples use the reference's clamped flat indices; their gradients go straight to the caller's maps (rare)
template <typename T>
__device__ __noinline__ void scatter_off_map(T* __restrict__ gz, T* __restrict__ gm, T ggx, T ggy, int H, int W,
                                             T z00, T z10, T z01, T z11, T m00, T m10, T m01, T m11) {
    const Corners c = flat_corners((long long)ggx, (long long)ggy, H, W);
    if (gz) { atomicAdd(gz + c.k00, z00); atomicAdd(gz + c.k10, z10); atomicAdd(gz + c.k01, z01); atomicAdd(gz + c.k11, z11); }
    if (gm) { atomicAdd(gm + c.k00, m00); atomicAdd(gm + c.k10, m10); atomicAdd(gm + c.k01, m01); atomicAdd(gm + c.k11, m11); }
}

// kappa-channel fix-up of a point that was off the map at its previous visit: pend = (kappa, gx, gy, -1)
template <typename T>
__device__ __noinline__ void fixup_off_map(T* __restrict__ gz, T coef, T ggx, T ggy, int H, int W) {
    const T fx = ggx - (T)(long long)ggx, fy = ggy - (T)(long long)ggy;
    const T gx = (T)1 - fx, gy = (T)1 - fy;
    scatter_off_map<T>(gz, nullptr, ggx, ggy, H, W, coef * gx * gy, coef * gx * fy, coef * fx * gy, coef * fx * fy,
                       (T)0, (T)0, (T)0, (T)0);
}

// raw-map pointers fetched from the warp's shared area only when a point is off the map
template <typename T>
struct LazyMaps {
    uint32_t area_s;
    __device__ __forceinline__ const T* z() const {
        const uint4 mp = lds_u4<16>(area_s);
        return reinterpret_cast<const T*>(((unsigned long long)mp.y << 32) | mp.x);
    }
    __device__ __forceinline__ const T* mu() const {
        const uint4 mp = lds_u4<16>(area_s);
        return reinterpret_cast<const T*>(((unsigned long long)mp.w << 32) | mp.z);
    }
};

// everything one warp keeps in static shared memory, addressed from ONE pinned 32-bit base
template <typename T>
struct __align__(16) SweepWarpArea {
    uint4 wc[2];         // [0] cell-table pointer (lo, hi), shared address of the warp's cache, ppl | N << 8 ; [1] zmap, fmap pointers
    Quad<T> fr[8];       // warp-uniform operands of the current step (MFB_SWEEP_FRAME_SMEM)
    Quad<T> pk[5];       // carried state adjoint while the point loop runs (MFB_SWEEP_PARK)
};

template <typename T>
__device__ __forceinline__ void flush_cell(T* __restrict__ gcell, int cell, const Quad<T>& qz, const Quad<T>& qm) {
    T* p = gcell + (long long)cell * kGradRec;
    red_quad(p, qz.v[0], qz.v[1], qz.v[2], qz.v[3]);
    red_quad(p + 4, qm.v[0], qm.v[1], qm.v[2], qm.v[3]);
}

// one thread per cell: record -> the four corner entries of the caller's gradient maps (+=)
template <typename T>
__global__ void finalize_map_grads_kernel(const T* __restrict__ gcell, T* __restrict__ g_z, T* __restrict__ g_mu,
                                          int n_maps, int H, int W) {
    const long long HW = (long long)H * W, total = HW * n_maps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / HW;
        const int k = (int)(i - m * HW);
        const Corners c = on_map_corners(k, H, W);
        const T* r = gcell + i * kGradRec;
        const Quad<T> qz = quad_load(reinterpret_cast<const Quad<T>*>(r)), qm = quad_load(reinterpret_cast<const Quad<T>*>(r + 4));
        if (g_z) {
            T* o = g_z + m * HW;
            if (qz.v[0] != (T)0) atomicAdd(o + c.k00, qz.v[0]);
            if (qz.v[1] != (T)0) atomicAdd(o + c.k10, qz.v[1]);
            if (qz.v[2] != (T)0) atomicAdd(o + c.k01, qz.v[2]);
            if (qz.v[3] != (T)0) atomicAdd(o + c.k11, qz.v[3]);
        }
        if (g_mu) {
            T* o = g_mu + m * HW;
            if (qm.v[0] != (T)0) atomicAdd(o + c.k00, qm.v[0]);
            if (qm.v[1] != (T)0) atomicAdd(o + c.k10, qm.v[1]);
            if (qm.v[2] != (T)0) atomicAdd(o + c.k01, qm.v[2]);
            if (qm.v[3] != (T)0) atomicAdd(o + c.k11, qm.v[3]);
        }
    }
}

// Reverse of the state update of one step (same algebra as the three-pass kernel).  In: adjoint of the
// post-update state (xb, vb, wb, Rb), pre-update state s, post-update angular velocity w_post.  Out: adjoint
// of the pre-update state (without the force-model terms), wd_bar (already masked by the clamp) and vd_bar.
template <typename T, int VARIANT>
__device__ __forceinline__ void reverse_update(const RolloutArgs<T>& a, const Body<T>& s, const T* w_post, T h,
                                               T* xb, T* vb, T* wb, T* Rb, T* wdm, T* vd_b) {
    T wd_b[3];
    if (VARIANT == kStepLoop) {
        // R' = R E(w'):  E_bar = R^T R'_bar ; R_bar = R'_bar E^T                     dphysics.py:290-324
        T Eb[9], E[9];
        {
            T th, inv;
            Mth<T>::norm_and_inv(w_post[0] * w_post[0] + w_post[1] * w_post[1] + w_post[2] * w_post[2], &th, &inv);
            const T k0 = w_post[0] * inv, k1 = w_post[1] * inv, k2 = w_post[2] * inv;
            T sn, c1;
            sin_versin(th * a.dt, &sn, &c1);
            const T kk = k0 * k0 + k1 * k1 + k2 * k2;
            E[0] = (T)1 + c1 * (k0 * k0 - kk);  E[1] = -sn * k2 + c1 * k0 * k1;     E[2] = sn * k1 + c1 * k0 * k2;
            E[3] = sn * k2 + c1 * k0 * k1;      E[4] = (T)1 + c1 * (k1 * k1 - kk);  E[5] = -sn * k0 + c1 * k1 * k2;
            E[6] = -sn * k1 + c1 * k0 * k2;     E[7] = sn * k0 + c1 * k1 * k2;      E[8] = (T)1 + c1 * (k2 * k2 - kk);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                Eb[r * 3 + c] = s.R[0 + r] * Rb[0 + c] + s.R[3 + r] * Rb[3 + c] + s.R[6 + r] * Rb[6 + c];
        T Rn[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                Rn[r * 3 + c] = Rb[r * 3 + 0] * E[c * 3 + 0] + Rb[r * 3 + 1] * E[c * 3 + 1] + Rb[r * 3 + 2] * E[c * 3 + 2];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rb[i] = Rn[i];
        rodrigues_right_bwd(w_post, a.dt, Eb, wb);
        // w' = w + wd dt ; x' = x + v' dt ; v' = v + vd dt
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            wd_b[i] = a.dt * wb[i];
            vb[i] += a.dt * xb[i];
            vd_b[i] = a.dt * vb[i];
        }
    } else {
        // x' = x + h v ; v' = v + h vd ; w' = w + h wd ; R' = R + h [w]x R         dphysics.py:499-528
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            wd_b[i] = h * wb[i];
            vd_b[i] = h * vb[i];
        }
        T M[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                M[r * 3 + c] = Rb[r * 3 + 0] * s.R[c * 3 + 0] + Rb[r * 3 + 1] * s.R[c * 3 + 1] + Rb[r * 3 + 2] * s.R[c * 3 + 2];
        T Rn[9];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Rn[0 + c] = Rb[0 + c] - h * (s.w[1] * Rb[6 + c] - s.w[2] * Rb[3 + c]);
            Rn[3 + c] = Rb[3 + c] - h * (s.w[2] * Rb[0 + c] - s.w[0] * Rb[6 + c]);
            Rn[6 + c] = Rb[6 + c] - h * (s.w[0] * Rb[3 + c] - s.w[1] * Rb[0 + c]);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) Rb[i] = Rn[i];
        const T ax0 = M[7] - M[5], ax1 = M[2] - M[6], ax2 = M[3] - M[1];
        vb[0] += h * xb[0]; vb[1] += h * xb[1]; vb[2] += h * xb[2];
        wb[0] += h * ax0; wb[1] += h * ax1; wb[2] += h * ax2;
    }
    // clamp mask of the angular acceleration: active <=> w' == fma(+-omega_max, h, w) bit-for-bit
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const bool hi = fma(a.omega_max, h, s.w[i]) == w_post[i];
        const bool lo = fma(-a.omega_max, h, s.w[i]) == w_post[i];
        wdm[i] = (hi || lo) ? (T)0 : wd_b[i];
    }
}

// SPLIT = 1: one warp per trajectory (the throughput shape).  SPLIT = kSweepWarps: the four warps of a CTA share ONE trajectory
// (small batches: training steps with 16 maps, terrain fitting with a single one, where the call is bound by the latency of
// one warp walking ~3200 instructions per step).  Warp `sub` visits the point slots j = sub, sub + SPLIT, ...; the per-step
// sums are combined across the warps through shared memory (two __syncthreads per step, buffers alternating with the step
// parity) and every warp closes the state adjoint redundantly, so all of them carry identical xb / vb / wb / Rb.
template <typename T, int VARIANT, bool HAS_FGRAD, int SPLIT = 1>
__global__ void __launch_bounds__(kSweepWarps * 32, sizeof(T) == 4 ? MFB_SWEEP_MINB : 1)
rollout_bwd_sweep_kernel(const RolloutArgs<T> a, const AdjointArgs<T> g) {
    static_assert(SPLIT == 1 || SPLIT == kSweepWarps, "a trajectory owns one warp or the whole CTA");
    // SPLIT > 1: per-warp partials of sum f_bar f and of the 24 running sums (one scalar each otherwise: the throughput
    // instantiation sits exactly at a shared-memory carve-out step with 4 CTAs per SM)
    __shared__ T xw_c[SPLIT > 1 ? 2 : 1][SPLIT > 1 ? kSweepWarps : 1];
    __shared__ T xw_s[SPLIT > 1 ? 2 : 1][SPLIT > 1 ? kSweepWarps : 1][SPLIT > 1 ? 24 : 1];
    static_assert(!(VARIANT == kOdeintEuler && HAS_FGRAD), "odeint + force gradients: use the three-pass kernel");
    __shared__ SweepPoints<T> tab;
    __shared__ SweepWarpArea<T> area_all[kSweepWarps];
    constexpr int kOffFr = (int)offsetof(SweepWarpArea<T>, fr), kQ = (int)sizeof(Quad<T>);
    const int ppl = (a.N + 31) >> 5;
    const int slots = ppl * 32;
    fill_sweep_points(tab, a, slots);
    __syncthreads();

    const int lane = lane_id();
    const int warp = warp_id_pinned();     // opaque: or the compiler re-derives it from S2R %tid all over the point loop
    const int sub = SPLIT > 1 ? warp : 0;                       // which share of the point slots this warp visits
    const int b = SPLIT > 1 ? (int)blockIdx.x : (int)blockIdx.x * kSweepWarps + warp;
    if (b >= a.B) return;

    const long long mi = b / a.map_group;                       // map of this trajectory (groups of consecutive trajectories share one)
    const T* __restrict__ zmap = a.z + mi * a.map_stride;
    const T* __restrict__ fmap = a.mu + mi * a.map_stride;
    const T* __restrict__ cells = a.cells + mi * a.cell_stride;
    T* __restrict__ gcell = g.g_cells ? g.g_cells + mi * g.g_cells_stride : nullptr;
    T* __restrict__ gz_dir = g.g_z ? g.g_z + mi * g.g_dir_stride : nullptr;
    T* __restrict__ gm_dir = g.g_mu ? g.g_mu + mi * g.g_dir_stride : nullptr;

    extern __shared__ __align__(16) unsigned char cache_raw[];
    // per contact point one 3-quad record: [0] d/dz of the corners of the point's current cell, [1] d/dfriction of the
    // same corners, [2] (kappa, fx, fy, cell) of the last visit.  48-byte lane stride: conflict-free 16-byte accesses.
    Quad<T>* const cache = reinterpret_cast<Quad<T>*>(cache_raw) + (size_t)(SPLIT > 1 ? 0 : warp) * 3 * slots;
    if (gcell) {
        Quad<T> zero; zero.v[0] = zero.v[1] = zero.v[2] = zero.v[3] = (T)0;
        Quad<T> none = zero; none.v[3] = pack_cell(-1, (T)0);
        for (int j = sub; j < ppl; j += SPLIT) {
            Quad<T>* rec = cache + 3 * (j * 32 + lane);
            quad_store(rec, zero); quad_store(rec + 1, zero); quad_store(rec + 2, none);
        }
    }

    const T* __restrict__ ctrl = a.controls + (long long)b * a.nT * 2;
    const T* __restrict__ csum = a.Csum + (long long)b * a.nT;
    const int H = a.H, W = a.W;
    const long long rowF = (long long)a.N * 3;
    const T* __restrict__ Xs_b = a.Xs + (long long)b * a.nT * 3;
    const T* __restrict__ Xd_b = a.Xds + (long long)b * a.nT * 3;
    const T* __restrict__ Rs_b = a.Rs + (long long)b * a.nT * 9;
    const T* __restrict__ Om_b = a.Oms + (long long)b * a.nT * 3;

    auto load_state = [&](Body<T>& s, int idx) {
        if (idx < 0) {
            load_body(s, a, b);
            s.x[2] = a.x0z[b];
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) { s.v[i] = Xd_b[idx * 3 + i]; s.w[i] = Om_b[idx * 3 + i]; }
#pragma unroll
            for (int i = 0; i < 9; ++i) s.R[i] = Rs_b[idx * 9 + i];
            s.x[0] = Xs_b[idx * 3 + 0] - s.R[2] * a.delta_h;                       // undo Xs = x + R[:,2] delta_h
            s.x[1] = Xs_b[idx * 3 + 1] - s.R[5] * a.delta_h;
            s.x[2] = Xs_b[idx * 3 + 2] - s.R[8] * a.delta_h;
        }
    };

    T xb[3] = {0, 0, 0}, vb[3] = {0, 0, 0}, wb[3] = {0, 0, 0}, Rb[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Rb[i] = (T)0;

    auto add_output_grads = [&](int idx) {
        if (g.g_Xs) {
            const T* p = g.g_Xs + ((long long)b * a.nT + idx) * 3;
            const T g0 = p[0], g1 = p[1], g2 = p[2];
            xb[0] += g0; xb[1] += g1; xb[2] += g2;
            Rb[2] += a.delta_h * g0; Rb[5] += a.delta_h * g1; Rb[8] += a.delta_h * g2;
        }
        if (g.g_Xds) {
            const T* p = g.g_Xds + ((long long)b * a.nT + idx) * 3;
            vb[0] += p[0]; vb[1] += p[1]; vb[2] += p[2];
        }
        if (g.g_Oms) {
            const T* p = g.g_Oms + ((long long)b * a.nT + idx) * 3;
            wb[0] += p[0]; wb[1] += p[1]; wb[2] += p[2];
        }
        if (g.g_Rs) {
            const T* p = g.g_Rs + ((long long)b * a.nT + idx) * 9;
#pragma unroll
            for (int i = 0; i < 9; ++i) Rb[i] += p[i];
        }
    };

    const int n_steps = (VARIANT == kOdeintEuler) ? a.nT - 1 : a.nT;
    if (VARIANT == kOdeintEuler && g.g_controls && lane == 0) {
        g.g_controls[((long long)b * a.nT + a.nT - 1) * 2 + 0] = (T)0;     // never used by the fixed-grid solver
        g.g_controls[((long long)b * a.nT + a.nT - 1) * 2 + 1] = (T)0;
    }

    T w_post[3] = {0, 0, 0};
    if (n_steps > 0) {
        const int last = (VARIANT == kOdeintEuler) ? n_steps : n_steps - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) w_post[i] = Om_b[last * 3 + i];
    }

    T Cb_prev = (T)0;        // C_bar of the step visited last: its map-gradient share is applied at this visit

    // Loop-invariant addresses the point loop needs.  At 128 registers the compiler re-derives them from the kernel
    // parameters at every point (64-bit multiply-adds, ~25 instructions per point); one 16-byte shared load is cheaper.
    if (lane == 0) {
        const unsigned long long cp = (unsigned long long)cells;
        const unsigned long long zp = (unsigned long long)zmap, fp = (unsigned long long)fmap;
        area_all[warp].wc[0] = make_uint4((unsigned)cp, (unsigned)(cp >> 32), (unsigned)__cvta_generic_to_shared(cache),
                                          (unsigned)ppl | ((unsigned)a.N << 8));
        area_all[warp].wc[1] = make_uint4((unsigned)zp, (unsigned)(zp >> 32), (unsigned)fp, (unsigned)(fp >> 32));
    }
    __syncwarp();
    unsigned area_s = (unsigned)__cvta_generic_to_shared(&area_all[warp]);
    asm volatile("mov.u32 %0, %0;" : "+r"(area_s));        // opaque: keep the base in a register, do not re-derive it

    for (int t = n_steps - 1; t >= 0; --t) {
        Body<T> s;
        load_state(s, VARIANT == kOdeintEuler ? t : t - 1);
        const T uv = ctrl[t * 2], uw = ctrl[t * 2 + 1];
        const T invC = Mth<T>::rcp(csum[t]);
        const int rec = (VARIANT == kOdeintEuler) ? t + 1 : t;
        add_output_grads(rec);

        T h = a.dt;
        if (VARIANT == kOdeintEuler) h = a.ts[t + 1] - a.ts[t];

        T wdm[3], vd_b[3], tq_b[3];
        reverse_update<T, VARIANT>(a, s, w_post, h, xb, vb, wb, Rb, wdm, vd_b);
        const T fs_b0 = vd_b[0] * a.inv_mass, fs_b1 = vd_b[1] * a.inv_mass, fs_b2 = vd_b[2] * a.inv_mass;
#pragma unroll
        for (int c = 0; c < 3; ++c) tq_b[c] = a.Iinv[0 + c] * wdm[0] + a.Iinv[3 + c] * wdm[1] + a.Iinv[6 + c] * wdm[2];

        StepFrame<T> f;
        make_frame(f, s, uv, uw, a.d_max, a.res, a.inv_res);
        const T hd_norm = Mth<T>::sqrt_rn(s.R[0] * s.R[0] + s.R[3] * s.R[3] + s.R[6] * s.R[6]);

        if (MFB_SWEEP_PARK) {
            Quad<T>* pk = area_all[warp].pk;
            Quad<T> q;
            q.v[0] = xb[0]; q.v[1] = xb[1]; q.v[2] = xb[2]; q.v[3] = vb[0]; quad_store(pk + 0, q);
            q.v[0] = vb[1]; q.v[1] = vb[2]; q.v[2] = wb[0]; q.v[3] = wb[1]; quad_store(pk + 1, q);
            q.v[0] = wb[2]; q.v[1] = Rb[0]; q.v[2] = Rb[1]; q.v[3] = Rb[2]; quad_store(pk + 2, q);
            q.v[0] = Rb[3]; q.v[1] = Rb[4]; q.v[2] = Rb[5]; q.v[3] = Rb[6]; quad_store(pk + 3, q);
            q.v[0] = Rb[7]; q.v[1] = Rb[8]; q.v[2] = hd_norm; q.v[3] = (T)0; quad_store(pk + 4, q);
            __syncwarp();
        }
        Quad<T>* const fr = area_all[warp].fr;
        if (MFB_SWEEP_FRAME_SMEM) {
            Quad<T> q;
            q.v[0] = f.v[0]; q.v[1] = f.v[1]; q.v[2] = f.v[2]; q.v[3] = f.w[0]; quad_store(fr + 0, q);
            q.v[0] = f.w[1]; q.v[1] = f.w[2]; q.v[2] = f.uv; q.v[3] = f.uw; quad_store(fr + 1, q);
            q.v[0] = f.hd[0]; q.v[1] = f.hd[1]; q.v[2] = f.hd[2]; q.v[3] = invC; quad_store(fr + 2, q);
            q.v[0] = tq_b[0]; q.v[1] = tq_b[1]; q.v[2] = tq_b[2]; q.v[3] = fs_b0; quad_store(fr + 3, q);
            q.v[0] = fs_b1; q.v[1] = fs_b2; q.v[2] = Cb_prev; q.v[3] = f.x[2]; quad_store(fr + 4, q);
            if (MFB_SWEEP_FRAME_SMEM >= 2) {
                q.v[0] = f.R[0]; q.v[1] = f.R[1]; q.v[2] = f.R[2]; q.v[3] = f.R[3]; quad_store(fr + 5, q);
                q.v[0] = f.R[4]; q.v[1] = f.R[5]; q.v[2] = f.R[6]; q.v[3] = f.R[7]; quad_store(fr + 6, q);
                q.v[0] = f.R[8]; q.v[1] = f.ox; q.v[2] = f.oy; q.v[3] = (T)0; quad_store(fr + 7, q);
            }
            __syncwarp();
        }
        // acc: 0-2 x_bar, 3-5 v_bar, 6-8 w_bar, 9-17 R_bar, 18-20 hd_bar, 21-22 controls, 23 sum f_bar f (-> C_bar)
        // kap: 0-2 x_bar per unit C_bar, 3-11 R_bar per unit C_bar
        T acc[24], kap[12];
#pragma unroll
        for (int k = 0; k < 24; ++k) acc[k] = (T)0;
#pragma unroll
        for (int k = 0; k < 12; ++k) kap[k] = (T)0;

#pragma unroll kSweepUnroll
        for (int j = sub; SPLIT > 1 ? j < ppl : true; j += SPLIT) {
            const uint4 wcst = lds_u4<0>(area_s);
            const T* __restrict__ cells_w = reinterpret_cast<const T*>(((unsigned long long)wcst.y << 32) | wcst.x);
            const int slot = j * 32 + lane;
            const bool ok = slot < (int)(wcst.w >> 8);
            if (MFB_SWEEP_PREFETCH && gcell && j + 1 < ppl) {
                // the NEXT point's cell record: with 16 warps per SM the table does not stay in L1 between two visits, but
                // a point rarely leaves its cell within one step, so the cell it was in at its previous visit (parked in
                // shared memory) tells where to prefetch
                const int pc = unpack_cell(cache[3 * (slot + 32) + 2].v[3]);
                if (pc >= 0) prefetch_l1(cells + (long long)pc * kCellStride);
            }
            const Quad<T> pq = quad_load(&tab.pp[slot]);
            const T px = pq.v[0], py = pq.v[1], pz = pq.v[2], side = pq.v[3], drv = tab.drv[slot];
            // warp-uniform operands of this step: registers, or re-read from shared memory (MFB_SWEEP_FRAME_SMEM)
            StepFrame<T> fl = f;
            T L_invC = invC, L_Cb_prev = Cb_prev, L_tq[3] = {tq_b[0], tq_b[1], tq_b[2]}, L_fs[3] = {fs_b0, fs_b1, fs_b2};
            T L_w[3] = {s.w[0], s.w[1], s.w[2]};
            if (MFB_SWEEP_FRAME_SMEM) {
                const Quad<T> q0 = lds_quad<kOffFr + 0 * kQ>(area_s, (T)0), q1 = lds_quad<kOffFr + 1 * kQ>(area_s, (T)0), q2 = lds_quad<kOffFr + 2 * kQ>(area_s, (T)0);
                fl.v[0] = q0.v[0]; fl.v[1] = q0.v[1]; fl.v[2] = q0.v[2];
                fl.w[0] = q0.v[3]; fl.w[1] = q1.v[0]; fl.w[2] = q1.v[1];
                fl.uv = q1.v[2]; fl.uw = q1.v[3];
                fl.hd[0] = q2.v[0]; fl.hd[1] = q2.v[1]; fl.hd[2] = q2.v[2];
                L_invC = q2.v[3];
                L_w[0] = fl.w[0]; L_w[1] = fl.w[1]; L_w[2] = fl.w[2];
                if (MFB_SWEEP_FRAME_SMEM >= 2) {
                    const Quad<T> q5 = lds_quad<kOffFr + 5 * kQ>(area_s, (T)0), q6 = lds_quad<kOffFr + 6 * kQ>(area_s, (T)0), q7 = lds_quad<kOffFr + 7 * kQ>(area_s, (T)0);
                    fl.R[0] = q5.v[0]; fl.R[1] = q5.v[1]; fl.R[2] = q5.v[2]; fl.R[3] = q5.v[3];
                    fl.R[4] = q6.v[0]; fl.R[5] = q6.v[1]; fl.R[6] = q6.v[2]; fl.R[7] = q6.v[3];
                    fl.R[8] = q7.v[0]; fl.ox = q7.v[1]; fl.oy = q7.v[2];
                }
            }
            PointEval<T> e;
            eval_point_maps<T, true>(e, fl, px, py, pz, drv, side, ok, cells_w, LazyMaps<T>{area_s}, H, W, a.inv_res, a.stiffness,
                                     a.damping);
            const T n0 = e.rec[4], n1 = e.rec[5], n2 = e.rec[6];
            const T fx = e.fx, fy = e.fy;
            const T r0 = e.r[0], r1 = e.r[1], r2 = e.r[2];
            int cell_p = e.cell;
            asm volatile("mov.s32 %0, %0;" : "+r"(cell_p));     // opaque: or the cell index is re-derived (2 F2I + 7) at each of its 3 uses
            if (MFB_SWEEP_FRAME_SMEM) {
                const Quad<T> q3 = lds_quad<kOffFr + 3 * kQ>(area_s, (T)0), q4 = lds_quad<kOffFr + 4 * kQ>(area_s, (T)0);
                L_tq[0] = q3.v[0]; L_tq[1] = q3.v[1]; L_tq[2] = q3.v[2];
                L_fs[0] = q3.v[3]; L_fs[1] = q4.v[0]; L_fs[2] = q4.v[1];
                L_Cb_prev = q4.v[2];
                if (MFB_SWEEP_FRAME_SMEM >= 2) fl.x[2] = q4.v[3];
            }

            // ---- phase 2 forward (dphysics.py:228-251) ----
            const T fo = e.sp * e.cw * L_invC;
            const T G0 = fo * n0, G1 = fo * n1, G2 = fo * n2;
            const T Fr0 = clampT(G0, a.mg), Fr1 = clampT(G1, a.mg), Fr2 = clampT(G2, a.mg);
            const T Nf2 = Fr0 * Fr0 + Fr1 * Fr1 + Fr2 * Fr2;
            const T Nf_inv = Nf2 > (T)0 ? Mth<T>::rsqrt(Nf2) : (T)0;
            const T Nf = Nf2 * Nf_inv;
            const T Hh0 = Nf * e.sl[0], Hh1 = Nf * e.sl[1], Hh2 = Nf * e.sl[2];
            const T Ft0 = clampT(Hh0, a.mg), Ft1 = clampT(Hh1, a.mg), Ft2 = clampT(Hh2, a.mg);
            const T F0 = Fr0 + Ft0, F1 = Fr1 + Ft1, F2 = Fr2 + Ft2;

            // ---- phase 2 reversed ----
            // torque = sum r x F :  F_bar += tq_bar x r ;  r_bar += F x tq_bar
            T Frb0 = fma(L_tq[1], r2, fma(-L_tq[2], r1, L_fs[0]));
            T Frb1 = fma(L_tq[2], r0, fma(-L_tq[0], r2, L_fs[1]));
            T Frb2 = fma(L_tq[0], r1, fma(-L_tq[1], r0, L_fs[2]));
            T Ftb0 = Frb0, Ftb1 = Frb1, Ftb2 = Frb2;
            if (HAS_FGRAD) {
                const long long o = ((long long)b * a.nT + rec) * rowF + (long long)slot * 3;
                if (ok) {
                    if (g.g_Fs) { Frb0 += g.g_Fs[o]; Frb1 += g.g_Fs[o + 1]; Frb2 += g.g_Fs[o + 2]; }
                    if (g.g_Ff) { Ftb0 += g.g_Ff[o]; Ftb1 += g.g_Ff[o + 1]; Ftb2 += g.g_Ff[o + 2]; }
                }
            }
            const T ab0 = F1 * L_tq[2] - F2 * L_tq[1];
            const T ab1 = F2 * L_tq[0] - F0 * L_tq[2];
            const T ab2 = F0 * L_tq[1] - F1 * L_tq[0];
            // F_friction = clamp(Nf * slip)
            const T Hb0 = gate(Ftb0, Hh0, a.mg), Hb1 = gate(Ftb1, Hh1, a.mg), Hb2 = gate(Ftb2, Hh2, a.mg);
            const T Nf_b = Hb0 * e.sl[0] + Hb1 * e.sl[1] + Hb2 * e.sl[2];
            const T sb0 = Nf * Hb0, sb1 = Nf * Hb1, sb2 = Nf * Hb2;
            // Nf = |F_spring|  (zero gradient at the origin, like torch.norm)
            {
                const T k = Nf_b * Nf_inv;
                Frb0 += k * Fr0; Frb1 += k * Fr1; Frb2 += k * Fr2;
            }
            // F_spring = clamp(f n),  f = sp c / C
            const T Gb0 = gate(Frb0, G0, a.mg), Gb1 = gate(Frb1, G1, a.mg), Gb2 = gate(Frb2, G2, a.mg);
            const T f_b = Gb0 * n0 + Gb1 * n1 + Gb2 * n2;
            acc[23] += f_b * fo;

            // ---- phase 1 reversed (own terms; the C_bar share goes through the kappa channel) ----
            T nb0 = fo * Gb0, nb1 = fo * Gb1, nb2 = fo * Gb2;
            const T sc_b = f_b * L_invC;
            const T sp_b = sc_b * e.cw;
            const T cw_b = sc_b * e.sp;
            // slip = d - dn n ; dn = d . n
            const T dn_b = -(sb0 * n0 + sb1 * n1 + sb2 * n2);
            nb0 = fma(dn_b, e.d[0], fma(-e.dn, sb0, nb0)); nb1 = fma(dn_b, e.d[1], fma(-e.dn, sb1, nb1));
            nb2 = fma(dn_b, e.d[2], fma(-e.dn, sb2, nb2));
            const T db0 = sb0 + dn_b * n0, db1 = sb1 + dn_b * n1, db2 = sb2 + dn_b * n2;
            // d = mu e ; e = tau hd - V
            const T mu_b = db0 * e.e[0] + db1 * e.e[1] + db2 * e.e[2];
            const T eb0 = e.mu * db0, eb1 = e.mu * db1, eb2 = e.mu * db2;
            const T tau_b = eb0 * fl.hd[0] + eb1 * fl.hd[1] + eb2 * fl.hd[2];
            acc[18] += e.tau * eb0; acc[19] += e.tau * eb1; acc[20] += e.tau * eb2;
            acc[21] += drv * tau_b; acc[22] += side * tau_b;
            T Vb0 = -eb0, Vb1 = -eb1, Vb2 = -eb2;
            // sp = -(k dh + beta vn) ; vn = V . n
            T dh_b = -a.stiffness * sp_b;
            const T vn_b = -a.damping * sp_b;
            Vb0 += vn_b * n0; Vb1 += vn_b * n1; Vb2 += vn_b * n2;
            nb0 += vn_b * e.V[0]; nb1 += vn_b * e.V[1]; nb2 += vn_b * e.V[2];
            // cw = sigmoid(-10 dh):  d cw / d dh = kappa
            const T kappa = (T)-10 * e.cw * ((T)1 - e.cw);
            dh_b += cw_b * kappa;
            const T zv_b = -dh_b;
            // n = (ax q, ay q, q),  q = (ax^2 + ay^2 + 1)^(-1/2),  ax = -cy / res, ay = -cx / res
            const T q = n2;
            const T ax = -e.rec[1] * a.inv_res, ay = -e.rec[2] * a.inv_res;
            const T q_b = nb0 * ax + nb1 * ay + nb2;
            const T q3 = q * q * q;
            const T ax_b = (nb0 * q - q_b * ax * q3) * a.inv_res;
            const T ay_b = (nb1 * q - q_b * ay * q3) * a.inv_res;
            // bilinear weights (height and friction share them)
            const T gx = (T)1 - fx, gy = (T)1 - fy;
            const T w00 = gx * gy, w10 = gx * fy, w01 = fx * gy, w11 = fx * fy;
            const T dz_dfy = e.rec[1] + fx * e.rec[3];
            const T fx_b = zv_b * e.dz_dfx + mu_b * (e.rec[10] + fy * e.rec[11]);
            const T fy_b = zv_b * dz_dfy + mu_b * (e.rec[9] + fx * e.rec[11]);

            if (gcell && ok) {
                Quad<T>* const rec = reinterpret_cast<Quad<T>*>(__cvta_shared_to_generic(wcst.z)) + 3 * slot;
                const Quad<T> pend = quad_load(rec + 2);
                const bool on_map = cell_p >= 0;
                {   // park this visit's (kappa, fx, fy, cell); off-map points keep their raw grid coordinates instead
                    Quad<T> np;
                    np.v[0] = kappa;
                    np.v[1] = on_map ? fx : r0 * a.inv_res + fl.ox;
                    np.v[2] = on_map ? fy : r1 * a.inv_res + fl.oy;
                    np.v[3] = pack_cell(cell_p, (T)0);
                    quad_store(rec + 2, np);
                }
                const int cur = unpack_cell(pend.v[3]);
                const bool same = cur == cell_p;
                Quad<T> qz = quad_load(rec), qm = quad_load(rec + 1);
                // (1) the previous visit's kappa-channel share, now that its C_bar is known
                const T coef = -L_Cb_prev * pend.v[0];
                if (cur >= 0) {
                    const T pfx = pend.v[1], pfy = pend.v[2];
                    const T c0 = coef - coef * pfx, c1 = coef * pfx;
                    qz.v[0] += c0 - c0 * pfy; qz.v[1] += c0 * pfy; qz.v[2] += c1 - c1 * pfy; qz.v[3] += c1 * pfy;
                    if (!same) flush_cell(gcell, cur, qz, qm);
                } else if (coef != (T)0) {
                    fixup_off_map(gz_dir, coef, pend.v[1], pend.v[2], H, W);
                }
                // (2) this visit's own terms
                const T cgz0 = zv_b * w00 + (ax_b + ay_b), cgz1 = zv_b * w10 - ax_b, cgz2 = zv_b * w01 - ay_b, cgz3 = zv_b * w11;
                if (on_map) {
                    const T keep = same ? (T)1 : (T)0;
                    qz.v[0] = keep * qz.v[0] + cgz0; qz.v[1] = keep * qz.v[1] + cgz1;
                    qz.v[2] = keep * qz.v[2] + cgz2; qz.v[3] = keep * qz.v[3] + cgz3;
                    qm.v[0] = keep * qm.v[0] + mu_b * w00; qm.v[1] = keep * qm.v[1] + mu_b * w10;
                    qm.v[2] = keep * qm.v[2] + mu_b * w01; qm.v[3] = keep * qm.v[3] + mu_b * w11;
                    quad_store(rec, qz); quad_store(rec + 1, qm);
                } else {
                    scatter_off_map(gz_dir, gm_dir, r0 * a.inv_res + fl.ox, r1 * a.inv_res + fl.oy, H, W, cgz0, cgz1, cgz2, cgz3,
                                    mu_b * w00, mu_b * w10, mu_b * w01, mu_b * w11);
                }
            }

            // grid coordinate -> world point ; V = v + w x r ; P = r + x ; r = R p
            const T ires = a.inv_res;
            const T Pb0 = fx_b * ires, Pb1 = fy_b * ires, Pb2 = dh_b;
            T rb0 = fma(Vb1, L_w[2], fma(-Vb2, L_w[1], ab0 + Pb0));
            T rb1 = fma(Vb2, L_w[0], fma(-Vb0, L_w[2], ab1 + Pb1));
            T rb2 = fma(Vb0, L_w[1], fma(-Vb1, L_w[0], ab2 + Pb2));
            if (!ok) { rb0 = rb1 = rb2 = (T)0; Vb0 = Vb1 = Vb2 = (T)0; }
            const T okf = ok ? (T)1 : (T)0;
            acc[0] += okf * Pb0; acc[1] += okf * Pb1; acc[2] += okf * Pb2;
            acc[3] += Vb0; acc[4] += Vb1; acc[5] += Vb2;
            acc[6] = fma(r1, Vb2, fma(-r2, Vb1, acc[6])); acc[7] = fma(r2, Vb0, fma(-r0, Vb2, acc[7]));
            acc[8] = fma(r0, Vb1, fma(-r1, Vb0, acc[8]));
            acc[9] += rb0 * px;  acc[10] += rb0 * py; acc[11] += rb0 * pz;
            acc[12] += rb1 * px; acc[13] += rb1 * py; acc[14] += rb1 * pz;
            acc[15] += rb2 * px; acc[16] += rb2 * py; acc[17] += rb2 * pz;
            // kappa channel: one unit of C_bar adds kappa to dh_bar  ->  P_bar += kappa (-dz/dfx / res, -dz/dfy / res, 1)
            {
                const T kk = ok ? kappa : (T)0;
                const T nk = -kk * ires;
                const T u0 = nk * e.dz_dfx, u1 = nk * dz_dfy;
                kap[0] += u0; kap[1] += u1; kap[2] += kk;
                kap[3] += u0 * px; kap[4] += u0 * py; kap[5] += u0 * pz;
                kap[6] += u1 * px; kap[7] += u1 * py; kap[8] += u1 * pz;
                kap[9] += kk * px; kap[10] += kk * py; kap[11] += kk * pz;
            }
            if (SPLIT == 1 && j + 1 >= (int)(wcst.w & 0xffu)) break;
        }

        if (MFB_SWEEP_FRAME_SMEM) {
            // bring the operands the epilogue of the step needs back from shared memory (they did not occupy registers meanwhile)
            const Quad<T> q0 = lds_quad<kOffFr + 0 * kQ>(area_s, (T)0), q1 = lds_quad<kOffFr + 1 * kQ>(area_s, (T)0), q2 = lds_quad<kOffFr + 2 * kQ>(area_s, (T)0);
            s.w[0] = q0.v[3]; s.w[1] = q1.v[0]; s.w[2] = q1.v[1];
            f.hd[0] = q2.v[0]; f.hd[1] = q2.v[1]; f.hd[2] = q2.v[2];
        }
        T hd_nrm = hd_norm;
        if (MFB_SWEEP_PARK) {
            const Quad<T>* pk = area_all[warp].pk;
            Quad<T> q;
            q = quad_load(pk + 0); xb[0] = q.v[0]; xb[1] = q.v[1]; xb[2] = q.v[2]; vb[0] = q.v[3];
            q = quad_load(pk + 1); vb[1] = q.v[0]; vb[2] = q.v[1]; wb[0] = q.v[2]; wb[1] = q.v[3];
            q = quad_load(pk + 2); wb[2] = q.v[0]; Rb[0] = q.v[1]; Rb[1] = q.v[2]; Rb[2] = q.v[3];
            q = quad_load(pk + 3); Rb[3] = q.v[0]; Rb[4] = q.v[1]; Rb[5] = q.v[2]; Rb[6] = q.v[3];
            q = quad_load(pk + 4); Rb[7] = q.v[0]; Rb[8] = q.v[1]; hd_nrm = q.v[2];
            __syncwarp();
        }
        // C_bar first (one butterfly), so that every lane can close its kappa channel before the big reduction
        T ff_sum = warp_sum(acc[23]);
        if (SPLIT > 1) {
            if (lane == 0) xw_c[t & 1][sub] = ff_sum;
            __syncthreads();
            ff_sum = (xw_c[t & 1][0] + xw_c[t & 1][1]) + (xw_c[t & 1][2] + xw_c[t & 1][3]);
        }
        const T C_b = -ff_sum * (MFB_SWEEP_FRAME_SMEM ? lds_quad<kOffFr + 2 * kQ>(area_s, (T)0).v[3] : invC);
        Cb_prev = C_b;
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[i] += C_b * kap[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[9 + i] += C_b * kap[3 + i];
        warp_sum8(acc, lane); warp_sum8(acc + 8, lane); warp_sum8(acc + 16, lane);
        if (SPLIT > 1) {
            // every lane holds its warp's 24 totals: lane k < 24 publishes total k, then all lanes add the four warps' rows
            T mine = acc[0];
#pragma unroll
            for (int k = 1; k < 24; ++k) mine = lane == k ? acc[k] : mine;
            if (lane < 24) xw_s[t & 1][sub][lane] = mine;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 24; ++k) acc[k] = (xw_s[t & 1][0][k] + xw_s[t & 1][1][k]) + (xw_s[t & 1][2][k] + xw_s[t & 1][3][k]);
        }

        // fold the per-point sums into the adjoint of the pre-update state
#pragma unroll
        for (int i = 0; i < 3; ++i) { xb[i] += acc[i]; vb[i] += acc[3 + i]; wb[i] += acc[6 + i]; }
#pragma unroll
        for (int i = 0; i < 9; ++i) Rb[i] += acc[9 + i];
        // hd = R[:,0] / max(|R[:,0]|, eps)
        {
            T a0, a1, a2;
            if (hd_nrm >= (T)1e-6) {
                const T dot = f.hd[0] * acc[18] + f.hd[1] * acc[19] + f.hd[2] * acc[20];
                const T inv = Mth<T>::inv(hd_nrm);
                a0 = (acc[18] - f.hd[0] * dot) * inv; a1 = (acc[19] - f.hd[1] * dot) * inv; a2 = (acc[20] - f.hd[2] * dot) * inv;
            } else {
                a0 = acc[18] * (T)1e6; a1 = acc[19] * (T)1e6; a2 = acc[20] * (T)1e6;
            }
            Rb[0] += a0; Rb[3] += a1; Rb[6] += a2;
        }
        if (g.g_controls && lane == 0 && sub == 0) {
            g.g_controls[((long long)b * a.nT + t) * 2 + 0] = acc[21];
            g.g_controls[((long long)b * a.nT + t) * 2 + 1] = acc[22];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) w_post[i] = s.w[i];
    }

    // ---------------- initial state: recorded index 0 (odeint) and the start-height snap ----------------
    if (VARIANT == kOdeintEuler) add_output_grads(0);
    {
        Body<T> s;
        load_body(s, a, b);
        T zb = xb[2] + (g.g_x0z ? g.g_x0z[b] : (T)0);   // gradient reaching the snapped height
        zb /= (T)a.N;
        T sx = (T)0, sy = (T)0, rr[6] = {0, 0, 0, 0, 0, 0};
        StepFrame<T> f;
        make_frame(f, s, (T)0, (T)0, a.d_max, a.res, a.inv_res);
        for (int j = sub; j < ppl; j += SPLIT) {
            const int slot = j * 32 + lane;
            const bool ok = slot < a.N;
            const Quad<T> pq = quad_load(&tab.pp[slot]);
            const T px = pq.v[0], py = pq.v[1], pz = pq.v[2];
            PointEval<T> e;
            eval_point(e, f, px, py, pz, (T)0, (T)0, ok, cells, zmap, fmap, H, W, a.inv_res, a.stiffness, a.damping);
            const T fx = e.fx, fy = e.fy;
            const T gx = (T)1 - fx, gy = (T)1 - fy;
            if (ok) {
                if (gcell) {
                    // drain: last visit's kappa share, then the snap's own contribution
                    const Quad<T>* rec = cache + 3 * slot;
                    const Quad<T> pend = quad_load(rec + 2);
                    const int cur = unpack_cell(pend.v[3]);
                    const T coef = -Cb_prev * pend.v[0];
                    if (cur >= 0) {
                        Quad<T> qz = quad_load(rec);
                        const Quad<T> qm = quad_load(rec + 1);
                        const T pfx = pend.v[1], pfy = pend.v[2];
                        const T c0 = coef * ((T)1 - pfx), c1 = coef * pfx;
                        qz.v[0] += c0 * ((T)1 - pfy); qz.v[1] += c0 * pfy; qz.v[2] += c1 * ((T)1 - pfy); qz.v[3] += c1 * pfy;
                        flush_cell(gcell, cur, qz, qm);
                    } else if (coef != (T)0) {
                        fixup_off_map(gz_dir, coef, pend.v[1], pend.v[2], H, W);
                    }
                    if (e.cell >= 0) {
                        red_quad(gcell + (long long)e.cell * kGradRec, zb * gx * gy, zb * gx * fy, zb * fx * gy, zb * fx * fy);
                    } else {
                        const T ggx = e.r[0] * a.inv_res + f.ox, ggy = e.r[1] * a.inv_res + f.oy;
                        scatter_off_map<T>(gz_dir, nullptr, ggx, ggy, H, W, zb * gx * gy, zb * gx * fy, zb * fx * gy, zb * fx * fy,
                                           (T)0, (T)0, (T)0, (T)0);
                    }
                }
                const T Pb0 = zb * e.dz_dfx * a.inv_res;
                const T Pb1 = zb * (e.rec[1] + fx * e.rec[3]) * a.inv_res;
                sx += Pb0; sy += Pb1;
                rr[0] += Pb0 * px; rr[1] += Pb0 * py; rr[2] += Pb0 * pz;
                rr[3] += Pb1 * px; rr[4] += Pb1 * py; rr[5] += Pb1 * pz;
            }
        }
        {
            T red[8] = {sx, sy, rr[0], rr[1], rr[2], rr[3], rr[4], rr[5]};
            warp_sum8(red, lane);
            if (SPLIT > 1) {
                __syncthreads();                       // the last step's readers are done with xw_s
                if (lane < 8) {
                    T mine = red[0];
#pragma unroll
                    for (int k = 1; k < 8; ++k) mine = lane == k ? red[k] : mine;
                    xw_s[0][sub][lane] = mine;
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < 8; ++k) red[k] = (xw_s[0][0][k] + xw_s[0][1][k]) + (xw_s[0][2][k] + xw_s[0][3][k]);
            }
            sx = red[0]; sy = red[1];
#pragma unroll
            for (int k = 0; k < 6; ++k) rr[k] = red[2 + k];
        }
        if (lane == 0 && sub == 0) {
            if (g.g_x0) { g.g_x0[b * 3 + 0] = xb[0] + sx; g.g_x0[b * 3 + 1] = xb[1] + sy; g.g_x0[b * 3 + 2] = (T)0; }
            if (g.g_xd0) { g.g_xd0[b * 3 + 0] = vb[0]; g.g_xd0[b * 3 + 1] = vb[1]; g.g_xd0[b * 3 + 2] = vb[2]; }
            if (g.g_om0) { g.g_om0[b * 3 + 0] = wb[0]; g.g_om0[b * 3 + 1] = wb[1]; g.g_om0[b * 3 + 2] = wb[2]; }
            if (g.g_R0) {
#pragma unroll
                for (int k = 0; k < 6; ++k) g.g_R0[b * 9 + k] = Rb[k] + rr[k];
#pragma unroll
                for (int k = 6; k < 9; ++k) g.g_R0[b * 9 + k] = Rb[k];
            }
        }
    }
}

}  // namespace mfb
