// K1w: the forward rollout for SMALL batches (planner shooting: 64 control sequences, monoforce_ros/nodes/monoforce_node.py:75;
// the reference's published timing: examples/diff_physics.ipynb, 64 x 600 steps).
//
// K1 (rollout_fwd.cuh) gives every trajectory ONE warp, which is the right shape when thousands of trajectories fill the
// GPU.  With B = 64 the call is bound by the latency of one warp walking ~1250 instructions per step (ncu, B = 64: one warp per
// scheduler issues 0.38 instructions / cycle => 1.7 us per step), most of them the 7 contact points of each lane.
// Here ONE CTA integrates one trajectory and every thread owns TWO contact points (ceil(N/64) <= 4 warps, one per scheduler):
// the per-point work of a step (phase 1 and phase 2, dphysics.py:172-272) shrinks from 7 points to 2 per thread, the two
// reductions of a step (soft-contact normaliser, then force / torque sums) go warp shuffle -> shared memory -> every thread
// adds the per-warp partials in the same order, so all threads carry the identical rigid-body state and integrate it
// redundantly (dphysics.py:274-324).  Two __syncthreads per step; the partial buffers alternate with the step parity, so no
// third barrier is needed.  Same arithmetic per point as K1 (eval_point); the sums are merely associated differently.
#pragma once
#include "rollout_fwd.cuh"

namespace mfb {

constexpr int kWideWarps = 4;                        // 4 warps x 32 lanes x 2 points = 256 contact points
constexpr int kWidePts = 2;

// Reduce-scatter half of warp_sum8: after the five butterfly stages the lanes with (lane >> 2) & 7 == k hold the warp's total of v[k].
template <typename T>
__device__ __forceinline__ T warp_reduce_scatter8(const T* v, int lane) {
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
    return t;
}

template <typename T, int VARIANT, bool FORCES, bool COST>
__global__ void __launch_bounds__(kWideWarps * 32)
rollout_fwd_wide_kernel(const RolloutArgs<T> a) {
    // per-warp partial sums; warps the block does not have stay zero
    __shared__ Quad<T> part_c[2];                     // [parity] -> one scalar per warp
    __shared__ Quad<T> part_s[2][kWideWarps][2];      // [parity][warp] -> 8 scalars
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    if (tid < 2) part_c[tid].v[0] = part_c[tid].v[1] = part_c[tid].v[2] = part_c[tid].v[3] = (T)0;
    if (tid < 2 * kWideWarps * 2) {
        Quad<T>* q = &part_s[0][0][0] + tid;
        q->v[0] = q->v[1] = q->v[2] = q->v[3] = (T)0;
    }

    // contact points of this thread: tid and tid + blockDim.x
    T px[kWidePts], py[kWidePts], pz[kWidePts], side[kWidePts], drv[kWidePts];
    bool ok[kWidePts];
#pragma unroll
    for (int j = 0; j < kWidePts; ++j) {
        const int p = tid + j * blockDim.x;
        ok[j] = p < a.N;
        px[j] = py[j] = pz[j] = side[j] = drv[j] = (T)0;
        if (ok[j]) {
            px[j] = a.pts[p * 3 + 0]; py[j] = a.pts[p * 3 + 1]; pz[j] = a.pts[p * 3 + 2];
            const int part = a.part[p];
            drv[j] = part >= 0 ? (T)1 : (T)0;
            side[j] = part < 0 ? (T)0 : ((part & 1) ? a.half_Ly : -a.half_Ly);           // dphysics.py:75-104
        }
    }
    __syncthreads();

    const long long mi = b / a.map_group;
    const T* __restrict__ zmap = a.z + mi * a.map_stride;
    const T* __restrict__ fmap = a.mu + mi * a.map_stride;
    const T* __restrict__ cells = a.cells + mi * a.cell_stride;
    const T* __restrict__ ctrl = a.controls + (long long)b * a.nT * 2;
    const int H = a.H, W = a.W;

    // sum over the block: warp butterfly, one partial per warp in shared memory, every thread adds the four partials
    auto block_sum1 = [&](T v, int parity) -> T {
        v = warp_sum(v);
        if (lane == 0) part_c[parity].v[warp] = v;
        __syncthreads();
        const Quad<T> q = quad_load(&part_c[parity]);
        return (q.v[0] + q.v[1]) + (q.v[2] + q.v[3]);
    };

    Body<T> s;
    load_body(s, a, b);

    // ---- start-height snap: x.z = mean_p interp(z, (R p + x).xy)            dphysics.py:567-571
    {
        StepFrame<T> f;
        make_frame(f, s, (T)0, (T)0, a.d_max, a.res, a.inv_res);
        T zsum = (T)0;
#pragma unroll
        for (int j = 0; j < kWidePts; ++j) {
            PointEval<T> e;
            eval_point(e, f, px[j], py[j], pz[j], (T)0, (T)0, true, cells, zmap, fmap, H, W, a.inv_res, a.stiffness, a.damping);
            const T zv = e.rec[0] + e.fy * e.rec[1] + e.fx * e.dz_dfx;
            zsum += ok[j] ? zv : (T)0;
        }
        const T acc = block_sum1(zsum, 0);
        s.x[2] = acc / (T)a.N;
        if (tid == 0) a.x0z[b] = s.x[2];
        __syncthreads();                               // part_c[0] is reused by step 0
    }

    T cost_mean = (T)0, cost_m2 = (T)0;
    const T inv_n = (T)1 / (T)a.N, inv_nm1 = (T)1 / (T)(a.N - 1);

    const long long rowF = (long long)a.N * 3;
    T* __restrict__ Fs_b = FORCES ? a.Fs + (long long)b * a.nT * rowF : nullptr;
    T* __restrict__ Ff_b = FORCES ? a.Ff + (long long)b * a.nT * rowF : nullptr;
    T* __restrict__ Xs_b = a.Xs + (long long)b * a.nT * 3;
    T* __restrict__ Xd_b = a.Xds + (long long)b * a.nT * 3;
    T* __restrict__ Rs_b = a.Rs + (long long)b * a.nT * 9;
    T* __restrict__ Om_b = a.Oms + (long long)b * a.nT * 3;
    T* __restrict__ Cs_b = a.Csum ? a.Csum + (long long)b * a.nT : nullptr;

    // every thread holds the whole state; lane 0 of four different warps (when the block has them) writes one piece each
    const int w_x = 0, w_v = 1 % n_warps, w_w = 2 % n_warps, w_R = 3 % n_warps;
    auto record_state = [&](int t) {
        if (lane != 0) return;
        if (warp == w_x) {
            Xs_b[t * 3 + 0] = s.x[0] + s.R[2] * a.delta_h;                     // dphysics.py:587-589
            Xs_b[t * 3 + 1] = s.x[1] + s.R[5] * a.delta_h;
            Xs_b[t * 3 + 2] = s.x[2] + s.R[8] * a.delta_h;
        }
        if (warp == w_v) { Xd_b[t * 3 + 0] = s.v[0]; Xd_b[t * 3 + 1] = s.v[1]; Xd_b[t * 3 + 2] = s.v[2]; }
        if (warp == w_w) { Om_b[t * 3 + 0] = s.w[0]; Om_b[t * 3 + 1] = s.w[1]; Om_b[t * 3 + 2] = s.w[2]; }
        if (warp == w_R) {
#pragma unroll
            for (int i = 0; i < 9; ++i) Rs_b[t * 9 + i] = s.R[i];
        }
    };

    T accF[kWidePts][6];                               // odeint variant: time-integrated forces      dphysics.py:457-465
#pragma unroll
    for (int j = 0; j < kWidePts; ++j)
#pragma unroll
        for (int k = 0; k < 6; ++k) accF[j][k] = (T)0;
    if (VARIANT == kOdeintEuler) {
        record_state(0);
        if (FORCES) {
#pragma unroll
            for (int j = 0; j < kWidePts; ++j) {
                const int p = tid + j * blockDim.x;
                if (ok[j]) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { Fs_b[p * 3 + k] = (T)0; Ff_b[p * 3 + k] = (T)0; }
                }
            }
        }
    }

    const int n_steps = (VARIANT == kOdeintEuler) ? a.nT - 1 : a.nT;
    // controls and step length of the NEXT step are fetched one step ahead: the loads are off the dependent chain
    T uv = (T)0, uw = (T)0, h = (T)0;
    if (n_steps > 0) {
        uv = ctrl[0]; uw = ctrl[1];
        if (VARIANT == kOdeintEuler) h = a.ts[1] - a.ts[0];
    }
    for (int t = 0; t < n_steps; ++t) {
        const int par = t & 1;
        const int tn = t + 1 < n_steps ? t + 1 : t;
        const T uv_n = ctrl[tn * 2], uw_n = ctrl[tn * 2 + 1];
        T h_n = (T)0;
        if (VARIANT == kOdeintEuler) h_n = a.ts[tn + 1] - a.ts[tn];

        StepFrame<T> f;
        make_frame(f, s, uv, uw, a.d_max, a.res, a.inv_res);

        // ---- phase 1 (own points)
        PointEval<T> e[kWidePts];
        T cw = (T)0;
#pragma unroll
        for (int j = 0; j < kWidePts; ++j) {
            eval_point<T, true>(e[j], f, px[j], py[j], pz[j], drv[j], side[j], ok[j], cells, zmap, fmap, H, W, a.inv_res,
                                a.stiffness, a.damping);
            cw += e[j].cw;
        }
        const T C = block_sum1(cw, par);
        const T invC = Mth<T>::rcp(C);
        if (tid == 0 && Cs_b) Cs_b[t] = C;             // tape of the single-sweep adjoint

        // ---- phase 2 (own points)                                          dphysics.py:228-251
        T red[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) red[k] = (T)0;
#pragma unroll
        for (int j = 0; j < kWidePts; ++j) {
            const PointEval<T>& q = e[j];
            const T fo = q.sp * q.cw * invC;
            const T Fr0 = clampT(fo * q.rec[4], a.mg), Fr1 = clampT(fo * q.rec[5], a.mg), Fr2 = clampT(fo * q.rec[6], a.mg);
            const T Nf = Mth<T>::sqrt(Fr0 * Fr0 + Fr1 * Fr1 + Fr2 * Fr2);
            const T Ft0 = clampT(Nf * q.sl[0], a.mg), Ft1 = clampT(Nf * q.sl[1], a.mg), Ft2 = clampT(Nf * q.sl[2], a.mg);
            const T F0 = Fr0 + Ft0, F1 = Fr1 + Ft1, F2 = Fr2 + Ft2;
            red[0] += F0; red[1] += F1; red[2] += F2;
            red[3] += q.r[1] * F2 - q.r[2] * F1;                              // dphysics.py:255
            red[4] += q.r[2] * F0 - q.r[0] * F2;
            red[5] += q.r[0] * F1 - q.r[1] * F0;
            if (COST) { red[6] += Nf; red[7] += Nf * Nf; }
            if (FORCES && ok[j]) {
                const int p = tid + j * blockDim.x;
                const int rec = VARIANT == kOdeintEuler ? t + 1 : t;
                T* fs = Fs_b + (long long)rec * rowF + p * 3;
                T* ff = Ff_b + (long long)rec * rowF + p * 3;
                if (VARIANT == kOdeintEuler) {
                    accF[j][0] += h * Fr0; accF[j][1] += h * Fr1; accF[j][2] += h * Fr2;
                    accF[j][3] += h * Ft0; accF[j][4] += h * Ft1; accF[j][5] += h * Ft2;
                    fs[0] = accF[j][0]; fs[1] = accF[j][1]; fs[2] = accF[j][2];
                    ff[0] = accF[j][3]; ff[1] = accF[j][4]; ff[2] = accF[j][5];
                } else {
                    fs[0] = Fr0; fs[1] = Fr1; fs[2] = Fr2;
                    ff[0] = Ft0; ff[1] = Ft1; ff[2] = Ft2;
                }
            }
        }
        {
            const T tot = warp_reduce_scatter8(red, lane);
            if ((lane & 3) == 0) {
                const int k = lane >> 2;
                part_s[par][warp][k >> 2].v[k & 3] = tot;
            }
        }
        __syncthreads();
        T sum[8];
        {
            Quad<T> lo[kWideWarps], hi[kWideWarps];
#pragma unroll
            for (int w = 0; w < kWideWarps; ++w) { lo[w] = quad_load(&part_s[par][w][0]); hi[w] = quad_load(&part_s[par][w][1]); }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                sum[k] = (lo[0].v[k] + lo[1].v[k]) + (lo[2].v[k] + lo[3].v[k]);
                sum[k + 4] = (hi[0].v[k] + hi[1].v[k]) + (hi[2].v[k] + hi[3].v[k]);
            }
        }

        if (COST) {
            // unbiased std over the N points of |F_spring|, then Welford over steps
            const T mean = sum[6] * inv_n;
            T var = (sum[7] - sum[6] * mean) * inv_nm1;
            var = Mth<T>::fmax_(var, (T)0);
            const T sd = Mth<T>::sqrt_rn(var);
            const T d = sd - cost_mean;
            cost_mean += d * Mth<T>::inv((T)(t + 1));
            cost_m2 += d * (sd - cost_mean);
        }

        // ---- angular / linear acceleration and the state update, redundantly in every thread     dphysics.py:255-288
        T wd[3], vd[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            wd[r] = clampT(a.Iinv[r * 3 + 0] * sum[3] + a.Iinv[r * 3 + 1] * sum[4] + a.Iinv[r * 3 + 2] * sum[5], a.omega_max);
        vd[0] = sum[0] * a.inv_mass;
        vd[1] = sum[1] * a.inv_mass;
        vd[2] = (sum[2] - a.mg) * a.inv_mass;
        if (VARIANT == kStepLoop) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                s.v[i] = fma(vd[i], a.dt, s.v[i]);
                s.x[i] = fma(s.v[i], a.dt, s.x[i]);
                s.w[i] = fma(wd[i], a.dt, s.w[i]);
            }
            rodrigues_right(s.R, s.w, a.dt);
            record_state(t);
        } else {
            T Rn[9];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Rn[0 + c] = s.R[0 + c] + h * (s.w[1] * s.R[6 + c] - s.w[2] * s.R[3 + c]);
                Rn[3 + c] = s.R[3 + c] + h * (s.w[2] * s.R[0 + c] - s.w[0] * s.R[6 + c]);
                Rn[6 + c] = s.R[6 + c] + h * (s.w[0] * s.R[3 + c] - s.w[1] * s.R[0 + c]);
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) s.R[i] = Rn[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                s.x[i] = fma(h, s.v[i], s.x[i]);
                s.v[i] = fma(h, vd[i], s.v[i]);
                s.w[i] = fma(wd[i], h, s.w[i]);
            }
            record_state(t + 1);
        }
        uv = uv_n; uw = uw_n; h = h_n;
    }
    if (COST && tid == 0) {
        const int n = n_steps;
        a.cost[b] = n > 1 ? Mth<T>::sqrt_rn(cost_m2 / (T)(n - 1)) : (T)0;
    }
}

}  // namespace mfb
