// K1: fused T-step rollout, forward.  One warp integrates one trajectory for all T steps;
// nothing but the recorded outputs ever leaves the SM.
//
// Per step (reference dphysics.py:172-288, :467-497):
//   phase 1 (per contact point, PPL points per lane, all in registers):
//       world point, point velocity, bilinear height/friction sample + cell normal,
//       soft-contact weight c, un-normalised spring-damper magnitude, tangential slip
//   warp reduction: C = sum_p c
//   phase 2 (per point): F_spring = clamp(s c n / C), F_friction = clamp(|F_spring| slip_t),
//       accumulate force / torque partial sums, emit the two force rows
//   warp reduction of 9 partial sums, then the semi-implicit Euler + Rodrigues update
//   redundantly in every lane, record the post-update state.
#pragma once
#include <type_traits>
#include "rollout_common.cuh"

namespace mfb {

constexpr int kFwdWarps = 4;   // trajectories per CTA
#ifndef MFB_FWD_REDO
#define MFB_FWD_REDO 1      // 1: branch-free phase 1 + whole-step redo when a point is off the map
#endif
#ifndef MFB_FWD_PTRS_SMEM
#define MFB_FWD_PTRS_SMEM 0   // 1: per-trajectory output pointers wait in shared memory instead of being re-derived from the kernel
                              // parameters every step.  Measured: 0 -> 2.72 ms, 1 -> 2.86 ms (more spills), so off
#endif
#ifndef MFB_FWD_MINB
#define MFB_FWD_MINB 4        // resident CTAs per SM the fp32 kernel is compiled for (128 registers)
#endif

// Force rows leave the SM through TMA bulk stores: every lane drops its points' forces into a
// per-warp shared-memory image of the (N,3) row (stride-3 words across lanes: conflict free), the
// 16-byte aligned interior of the row then goes out as ONE cp.async.bulk per tensor and step, the
// <= 3 leading / trailing scalars with plain stores.  A row of N*3 scalars starts at an arbitrary
// 4-byte (8-byte) phase in global memory, so its shared image is placed at the same phase.
template <typename T> struct RowStage {
    static constexpr int kPer16 = 16 / (int)sizeof(T);            // scalars per 16 bytes
    __host__ __device__ static int stride(int N) { return (N * 3 + 2 * kPer16 + kPer16 - 1) / kPer16 * kPer16; }
    // rows per warp: the two row images, plus (odeint variant) the two running time integrals of the forces
    __host__ __device__ static constexpr int rows(int variant) { return variant == kOdeintEuler ? 4 : 2; }
    __host__ static size_t smem_bytes(int N, int warps, int variant) { return (size_t)warps * rows(variant) * stride(N) * sizeof(T); }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Emits the two staged rows (same phase): `img_*` are 16-byte aligned shared buffers, rows sit at img + phase.
template <typename T>
__device__ __forceinline__ void emit_rows(T* __restrict__ grow_s, T* __restrict__ grow_f, const T* img_s, const T* img_f,
                                          int phase, int n, int lane) {
    constexpr int P = RowStage<T>::kPer16;
    const int head = (P - phase) & (P - 1);                 // scalars before the first 16-byte boundary
    const int body = (n - head) & ~(P - 1);                 // scalars in the aligned interior
    const int tail = n - head - body;
    if (lane == 0) {
        if (body > 0) bulk_store(grow_s + head, img_s + phase + head, (uint32_t)(body * sizeof(T)));
        if (body > 0) bulk_store(grow_f + head, img_f + phase + head, (uint32_t)(body * sizeof(T)));
        bulk_commit();
    }
    // lanes 8..8+head-1 write the head scalars, lanes 16..16+tail-1 the tail scalars (at most 3 each)
    const int hk = lane - 8, tk = lane - 16;
    if ((unsigned)hk < (unsigned)head) { grow_s[hk] = img_s[phase + hk]; grow_f[hk] = img_f[phase + hk]; }
    if ((unsigned)tk < (unsigned)tail) {
        const int o = head + body + tk;
        grow_s[o] = img_s[phase + o]; grow_f[o] = img_f[phase + o];
    }
}

template <typename T, int PPL, int VARIANT, bool FORCES, bool COST, bool JOINTS = false>
__global__ void __launch_bounds__(kFwdWarps * 32, (sizeof(T) == 4 && PPL <= 7) ? MFB_FWD_MINB : 1)
rollout_fwd_kernel(const RolloutArgs<T> a) {
    __shared__ PointTable<T> tab;          // scalar arrays: snap + moving-flipper variant
    __shared__ SweepPoints<T> ptab;        // packed (px, py, pz, side) + driven flag: the step loop of the static variant
    if (JOINTS) fill_point_table(tab, a, PPL * 32);
    else fill_sweep_points(ptab, a, PPL * 32);
    __syncthreads();

    const int lane = lane_id();
    const int b = blockIdx.x * kFwdWarps + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const int n_last = a.N - (PPL - 1) * 32;            // valid lanes of the last slot
    const bool last_valid = lane < n_last;

    extern __shared__ __align__(16) unsigned char stage_raw[];
    const int row_stride = RowStage<T>::stride(a.N);
    T* const img_s = reinterpret_cast<T*>(stage_raw) + (size_t)(threadIdx.x >> 5) * RowStage<T>::rows(VARIANT) * row_stride;   // F_spring row image
    T* const img_f = img_s + row_stride;                                                                 // F_friction row image
    // odeint variant: the time-integrated forces (dphysics.py:457-465) accumulate in shared memory at fixed (phase 0) positions;
    // 42 accumulators per lane in registers spilled at 128 registers / thread (4.4 ms vs 2.8 ms for the step loop at config 3)
    T* const acc_s = img_f + row_stride;
    T* const acc_f = acc_s + row_stride;

    const long long mi = b / a.map_group;                       // map of this trajectory (groups of consecutive trajectories share one)
    const T* __restrict__ zmap = a.z + mi * a.map_stride;
    const T* __restrict__ fmap = a.mu + mi * a.map_stride;
    const T* __restrict__ cells = a.cells + mi * a.cell_stride;
    const T* __restrict__ ctrl = a.controls + (long long)b * a.nT * 2;
    const int H = a.H, W = a.W;

    Body<T> s;
    load_body(s, a, b);

    // ---- start-height snap: x.z = mean_p interp(z, (R p + x).xy)            dphysics.py:567-571
    {
        StepFrame<T> f;
        make_frame(f, s, (T)0, (T)0, a.d_max, a.res, a.inv_res);
        T acc = (T)0;
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const int slot = j * 32 + lane;
            T px, py, pz;
            if (JOINTS) {
                px = tab.px[slot]; py = tab.py[slot]; pz = tab.pz[slot];
            } else {
                const Quad<T> pq = quad_load(&ptab.pp[slot]);
                px = pq.v[0]; py = pq.v[1]; pz = pq.v[2];
            }
            PointEval<T> e;
            eval_point(e, f, px, py, pz, (T)0, (T)0, true, cells, zmap, fmap, H, W, a.inv_res, a.stiffness, a.damping);
            const T zv = e.rec[0] + e.fy * e.rec[1] + e.fx * e.dz_dfx;
            acc += (j == PPL - 1 && !last_valid) ? (T)0 : zv;
        }
        acc = warp_sum(acc);
        s.x[2] = acc / (T)a.N;
        if (lane == 0) a.x0z[b] = s.x[2];
    }

    // cost accumulators (Welford over steps of the per-step std over points)
    T cost_mean = (T)0, cost_m2 = (T)0;
    const T inv_n = (T)1 / (T)a.N, inv_nm1 = (T)1 / (T)(a.N - 1);    // N == 1: 0 * inf = NaN, like torch.std

    const long long rowF = (long long)a.N * 3;
    T* __restrict__ Fs_b = FORCES ? a.Fs + (long long)b * a.nT * rowF : nullptr;
    T* __restrict__ Ff_b = FORCES ? a.Ff + (long long)b * a.nT * rowF : nullptr;
    T* __restrict__ Xs_b = a.Xs + (long long)b * a.nT * 3;
    T* __restrict__ Xd_b = a.Xds + (long long)b * a.nT * 3;
    T* __restrict__ Rs_b = a.Rs + (long long)b * a.nT * 9;
    T* __restrict__ Om_b = a.Oms + (long long)b * a.nT * 3;
    T* __restrict__ Cs_b = a.Csum ? a.Csum + (long long)b * a.nT : nullptr;

    // [0] (Xs_b, Xd_b)  [1] (Rs_b, Om_b)  [2] (Fs_b, Ff_b)  [3] (controls, contact_sum) of this warp's trajectory
    __shared__ uint4 fptr_all[MFB_FWD_PTRS_SMEM ? kFwdWarps * 4 : 1];
    unsigned fptr_s = 0;
    if (MFB_FWD_PTRS_SMEM) {
        uint4* fp = fptr_all + (threadIdx.x >> 5) * 4;
        if (lane == 0) {
            auto pk = [](const void* p0, const void* p1) {
                const unsigned long long u0 = (unsigned long long)p0, u1 = (unsigned long long)p1;
                return make_uint4((unsigned)u0, (unsigned)(u0 >> 32), (unsigned)u1, (unsigned)(u1 >> 32));
            };
            fp[0] = pk(Xs_b, Xd_b); fp[1] = pk(Rs_b, Om_b); fp[2] = pk(Fs_b, Ff_b); fp[3] = pk(ctrl, Cs_b);
        }
        __syncwarp();
        fptr_s = (unsigned)__cvta_generic_to_shared(fp);
        asm volatile("mov.u32 %0, %0;" : "+r"(fptr_s));
    }

    auto record_state = [&](int t) {
        if (lane == 0) {
            T *pXs = Xs_b, *pXd = Xd_b, *pRs = Rs_b, *pOm = Om_b;
            if (MFB_FWD_PTRS_SMEM) {
                const uint4 p0 = lds_u4<0>(fptr_s), p1 = lds_u4<16>(fptr_s);
                pXs = reinterpret_cast<T*>(u64_of(p0.x, p0.y)); pXd = reinterpret_cast<T*>(u64_of(p0.z, p0.w));
                pRs = reinterpret_cast<T*>(u64_of(p1.x, p1.y)); pOm = reinterpret_cast<T*>(u64_of(p1.z, p1.w));
            }
            // Xs = x + R[:,2] * delta_h                                          dphysics.py:587-589
            pXs[t * 3 + 0] = s.x[0] + s.R[2] * a.delta_h;
            pXs[t * 3 + 1] = s.x[1] + s.R[5] * a.delta_h;
            pXs[t * 3 + 2] = s.x[2] + s.R[8] * a.delta_h;
#pragma unroll
            for (int i = 0; i < 3; ++i) { pXd[t * 3 + i] = s.v[i]; pOm[t * 3 + i] = s.w[i]; }
#pragma unroll
            for (int i = 0; i < 9; ++i) pRs[t * 9 + i] = s.R[i];
        }
    };

    if (VARIANT == kOdeintEuler) {
        record_state(0);
        if (FORCES) {
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                if (j < PPL - 1 || last_valid) {
                    const int o = (j * 32 + lane) * 3;         // also this lane's slots of the running integrals (never shared)
#pragma unroll
                    for (int k = 0; k < 3; ++k) { acc_s[o + k] = (T)0; acc_f[o + k] = (T)0; Fs_b[o + k] = (T)0; Ff_b[o + k] = (T)0; }
                }
            }
        }
    }

    const int n_steps = (VARIANT == kOdeintEuler) ? a.nT - 1 : a.nT;
    T uv = ctrl[0], uw = ctrl[1];

    for (int t = 0; t < n_steps; ++t) {
        // prefetch next controls
        T uv_n = uv, uw_n = uw;
        if (t + 1 < a.nT) {
            const T* cp = ctrl;
            if (MFB_FWD_PTRS_SMEM) { const uint4 p3 = lds_u4<48>(fptr_s); cp = reinterpret_cast<const T*>(u64_of(p3.x, p3.y)); }
            uv_n = cp[(t + 1) * 2]; uw_n = cp[(t + 1) * 2 + 1];
        }

        StepFrame<T> f;
        make_frame(f, s, uv, uw, a.d_max, a.res, a.inv_res);

        T nrm[PPL][3], sc[PPL], slip[PPL][3], arm[PPL][3];
        T C = (T)0;

        // moving flippers: lane i < 4 owns (cos, sin) of joint angle i at this step     dphysics.py:187-197
        T jc = (T)1, js = (T)0, mom[6] = {0, 0, 0, 0, 0, 0};
        if (JOINTS) {
            const T ang = lane < 4 ? a.joint_angles[((long long)b * a.nT + t) * 4 + lane] : (T)0;
            Mth<T>::sincos(ang, &js, &jc);
        }

        // phase 1 over the lane's points.  First without the off-map patch (branch free: the loads of all points are in
        // flight together); if any lane saw a point outside the map the step is redone with the patching version.
        auto phase1 = [&](auto patch) -> bool {
            constexpr bool kPatch = decltype(patch)::value;
            bool off = false;
            C = (T)0;
            if (JOINTS) {
#pragma unroll
                for (int k = 0; k < 6; ++k) mom[k] = (T)0;
            }
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                const int slot = j * 32 + lane;
                T px, py, pz, drv, side;
                if (JOINTS) {
                    px = tab.px[slot]; py = tab.py[slot]; pz = tab.pz[slot]; drv = tab.driven[slot]; side = tab.side[slot];
                } else {
                    const Quad<T> pq = quad_load(&ptab.pp[slot]);
                    px = pq.v[0]; py = pq.v[1]; pz = pq.v[2]; side = pq.v[3]; drv = ptab.drv[slot];
                }
                if (JOINTS) {
                    articulate_point(px, py, pz, tab.part[slot], jc, js, a.joint_pos);
                    if ((j < PPL - 1) || last_valid) {
                        mom[0] += py * py + pz * pz; mom[1] += px * px + pz * pz; mom[2] += px * px + py * py;
                        mom[3] += px * py; mom[4] += px * pz; mom[5] += py * pz;
                    }
                }
                PointEval<T> e;
                eval_point<T, kPatch>(e, f, px, py, pz, drv, side,
                                      (j < PPL - 1) || last_valid, cells, zmap, fmap, H, W, a.inv_res, a.stiffness, a.damping);
                off |= e.cell < 0;
                C += e.cw;
                sc[j] = e.sp * e.cw;
#pragma unroll
                for (int k = 0; k < 3; ++k) { slip[j][k] = e.sl[k]; nrm[j][k] = e.rec[4 + k]; arm[j][k] = e.r[k]; }
            }
            return off;
        };
#if MFB_FWD_REDO
        if (__any_sync(kFull, phase1(std::false_type{}))) phase1(std::true_type{});
#else
        phase1(std::true_type{});
#endif

        C = warp_sum(C);
        const T invC = Mth<T>::rcp(C);
        if (lane == 0) {                                                    // tape of the single-sweep adjoint
            T* cs = Cs_b;
            if (MFB_FWD_PTRS_SMEM) { const uint4 p3 = lds_u4<48>(fptr_s); cs = reinterpret_cast<T*>(u64_of(p3.z, p3.w)); }
            if (cs) cs[t] = C;
        }

        T sum[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) sum[k] = (T)0;
        T nf_sum = (T)0, nf_sq = (T)0;
        T h = (T)0;
        if (VARIANT == kOdeintEuler) h = a.ts[t + 1] - a.ts[t];

        // row pointers are re-derived from t every step: carrying them across the loop costs 5 registers and measured
        // +0.19 ms (spills) at 128 registers / thread
        T *Fs_t = nullptr, *Ff_t = nullptr;
        if (FORCES) {
            T *pFs = Fs_b, *pFf = Ff_b;
            if (MFB_FWD_PTRS_SMEM) {
                const uint4 p2 = lds_u4<32>(fptr_s);
                pFs = reinterpret_cast<T*>(u64_of(p2.x, p2.y)); pFf = reinterpret_cast<T*>(u64_of(p2.z, p2.w));
            }
            Fs_t = pFs + (long long)(VARIANT == kOdeintEuler ? t + 1 : t) * rowF;
            Ff_t = pFf + (long long)(VARIANT == kOdeintEuler ? t + 1 : t) * rowF;
        }
        // phase of the row start inside a 16-byte line (both tensors share it: same shape, 16-byte aligned bases)
        const int phase = FORCES ? (int)((((long long)b * a.nT + (VARIANT == kOdeintEuler ? t + 1 : t)) * rowF) &
                                         (RowStage<T>::kPer16 - 1)) : 0;
        if (FORCES) {
            // the previous step's bulk stores must have finished READING the row images
            if (lane == 0) bulk_wait_read();
            __syncwarp();
        }

#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const T f = sc[j] * invC;
            const T Fr0 = clampT(f * nrm[j][0], a.mg), Fr1 = clampT(f * nrm[j][1], a.mg), Fr2 = clampT(f * nrm[j][2], a.mg);
            const T Nf = Mth<T>::sqrt(Fr0 * Fr0 + Fr1 * Fr1 + Fr2 * Fr2);                   // dphysics.py:238
            const T Ft0 = clampT(Nf * slip[j][0], a.mg), Ft1 = clampT(Nf * slip[j][1], a.mg), Ft2 = clampT(Nf * slip[j][2], a.mg);
            const T F0 = Fr0 + Ft0, F1 = Fr1 + Ft1, F2 = Fr2 + Ft2;
            sum[0] += F0; sum[1] += F1; sum[2] += F2;                                        // dphysics.py:265 (F_spring + F_friction)
            sum[3] = fma(arm[j][1], F2, fma(-arm[j][2], F1, sum[3]));                        // dphysics.py:255
            sum[4] = fma(arm[j][2], F0, fma(-arm[j][0], F2, sum[4]));
            sum[5] = fma(arm[j][0], F1, fma(-arm[j][1], F0, sum[5]));
            if (COST) { nf_sum += Nf; nf_sq += Nf * Nf; }
            if (FORCES) {
                if (j < PPL - 1 || last_valid) {
                    const int o = phase + (j * 32 + lane) * 3;
                    if (VARIANT == kOdeintEuler) {
                        const int oa = (j * 32 + lane) * 3;
                        const T s0 = acc_s[oa + 0] + h * Fr0, s1 = acc_s[oa + 1] + h * Fr1, s2 = acc_s[oa + 2] + h * Fr2;
                        const T f0 = acc_f[oa + 0] + h * Ft0, f1 = acc_f[oa + 1] + h * Ft1, f2 = acc_f[oa + 2] + h * Ft2;
                        acc_s[oa + 0] = s0; acc_s[oa + 1] = s1; acc_s[oa + 2] = s2;
                        acc_f[oa + 0] = f0; acc_f[oa + 1] = f1; acc_f[oa + 2] = f2;
                        img_s[o + 0] = s0; img_s[o + 1] = s1; img_s[o + 2] = s2;
                        img_f[o + 0] = f0; img_f[o + 1] = f1; img_f[o + 2] = f2;
                    } else {
                        img_s[o + 0] = Fr0; img_s[o + 1] = Fr1; img_s[o + 2] = Fr2;
                        img_f[o + 0] = Ft0; img_f[o + 1] = Ft1; img_f[o + 2] = Ft2;
                    }
                }
            }
        }
        if (FORCES) {
            // generic-proxy writes -> visible to the async proxy, then one lane launches the two bulk stores
            fence_async_smem();
            __syncwarp();
            emit_rows(Fs_t, Ff_t, img_s, img_f, phase, (int)rowF, lane);

        }
        {
            T red[8] = {sum[0], sum[1], sum[2], sum[3], sum[4], sum[5], nf_sum, nf_sq};
            warp_sum8(red, lane);
#pragma unroll
            for (int k = 0; k < 6; ++k) sum[k] = red[k];
            nf_sum = red[6]; nf_sq = red[7];
        }

        if (COST) {
            // unbiased std over the N points of |F_spring|, then Welford over steps
            const T mean = nf_sum * inv_n;
            T var = (nf_sq - nf_sum * mean) * inv_nm1;
            var = Mth<T>::fmax_(var, (T)0);
            const T sd = Mth<T>::sqrt_rn(var);
            const T d = sd - cost_mean;
            cost_mean += d * Mth<T>::inv((T)(t + 1));
            cost_m2 += d * (sd - cost_mean);
        }

        // angular / linear acceleration                                         dphysics.py:255-266
        T wd[3], vd[3];
        T Iinv[9];
        if (JOINTS) {
            // inertia of the articulated point set about the body origin, inverted every step      dphysics.py:196-197
            const T mp = a.mass / (T)a.N;
#pragma unroll
            for (int k = 0; k < 6; ++k) mom[k] = warp_sum(mom[k]) * mp;
            invert_inertia(mom, Iinv);
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) Iinv[k] = a.Iinv[k];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            wd[r] = clampT(Iinv[r * 3 + 0] * sum[3] + Iinv[r * 3 + 1] * sum[4] + Iinv[r * 3 + 2] * sum[5], a.omega_max);
        }
        vd[0] = sum[0] * a.inv_mass;
        vd[1] = sum[1] * a.inv_mass;
        vd[2] = (sum[2] - a.mg) * a.inv_mass;

        if (VARIANT == kStepLoop) {
            // semi-implicit Euler                                               dphysics.py:274-288
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                s.v[i] = fma(vd[i], a.dt, s.v[i]);
                s.x[i] = fma(s.v[i], a.dt, s.x[i]);
                s.w[i] = fma(wd[i], a.dt, s.w[i]);      // explicit fma: the adjoint re-derives the clamp mask from it
            }
            rodrigues_right(s.R, s.w, a.dt);
            record_state(t);
        } else {
            // fixed-grid explicit Euler on the solver grid, dR = [w]x R          dphysics.py:258-259, :499-528
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
        uv = uv_n; uw = uw_n;
    }

    if (FORCES && lane == 0) bulk_wait_all();
    if (COST && lane == 0) {
        const int n = n_steps;
        a.cost[b] = n > 1 ? Mth<T>::sqrt_rn(cost_m2 / (T)(n - 1)) : (T)0;
    }
}

}  // namespace mfb
