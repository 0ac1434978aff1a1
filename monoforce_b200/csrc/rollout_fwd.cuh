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
#include "rollout_common.cuh"

namespace mfb {

constexpr int kFwdWarps = 4;   // trajectories per CTA

template <typename T, int PPL, int VARIANT, bool FORCES, bool COST>
__global__ void __launch_bounds__(kFwdWarps * 32)
rollout_fwd_kernel(const RolloutArgs<T> a) {
    __shared__ PointTable<T> tab;
    fill_point_table(tab, a, PPL * 32);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kFwdWarps + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const int n_last = a.N - (PPL - 1) * 32;            // valid lanes of the last slot
    const bool last_valid = lane < n_last;

    const T* __restrict__ zmap = a.z + (long long)b * a.map_stride;
    const T* __restrict__ fmap = a.mu + (long long)b * a.map_stride;
    const T* __restrict__ ctrl = a.controls + (long long)b * a.nT * 2;
    const int H = a.H, W = a.W;

    Body<T> s;
    load_body(s, a, b);

    // ---- start-height snap: x.z = mean_p interp(z, (R p + x).xy)            dphysics.py:567-571
    {
        T acc = (T)0;
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const int slot = j * 32 + lane;
            const T px = tab.px[slot], py = tab.py[slot], pz = tab.pz[slot];
            const T Px = s.R[0] * px + s.R[1] * py + s.R[2] * pz + s.x[0];
            const T Py = s.R[3] * px + s.R[4] * py + s.R[5] * pz + s.x[1];
            T fx, fy;
            const Cell c = locate(Mth<T>::to_cells(Px, a.d_max, a.res, a.inv_res),
                                  Mth<T>::to_cells(Py, a.d_max, a.res, a.inv_res), H, W, fx, fy);
            T zv = blend(fx, fy, ldg(zmap + c.k00), ldg(zmap + c.k10), ldg(zmap + c.k01), ldg(zmap + c.k11));
            if (j == PPL - 1 && !last_valid) zv = (T)0;
            acc += zv;
        }
        acc = warp_sum(acc);
        s.x[2] = acc / (T)a.N;
        if (lane == 0) a.x0z[b] = s.x[2];
    }

    // cost accumulators (Welford over steps of the per-step std over points)
    T cost_mean = (T)0, cost_m2 = (T)0;

    const long long rowF = (long long)a.N * 3;
    T* __restrict__ Fs_b = FORCES ? a.Fs + (long long)b * a.nT * rowF : nullptr;
    T* __restrict__ Ff_b = FORCES ? a.Ff + (long long)b * a.nT * rowF : nullptr;
    T* __restrict__ Xs_b = a.Xs + (long long)b * a.nT * 3;
    T* __restrict__ Xd_b = a.Xds + (long long)b * a.nT * 3;
    T* __restrict__ Rs_b = a.Rs + (long long)b * a.nT * 9;
    T* __restrict__ Om_b = a.Oms + (long long)b * a.nT * 3;

    auto record_state = [&](int t) {
        if (lane == 0) {
            // Xs = x + R[:,2] * delta_h                                          dphysics.py:587-589
            Xs_b[t * 3 + 0] = s.x[0] + s.R[2] * a.delta_h;
            Xs_b[t * 3 + 1] = s.x[1] + s.R[5] * a.delta_h;
            Xs_b[t * 3 + 2] = s.x[2] + s.R[8] * a.delta_h;
#pragma unroll
            for (int i = 0; i < 3; ++i) { Xd_b[t * 3 + i] = s.v[i]; Om_b[t * 3 + i] = s.w[i]; }
#pragma unroll
            for (int i = 0; i < 9; ++i) Rs_b[t * 9 + i] = s.R[i];
        }
    };

    // odeint variant: time-integrated forces live in registers                 dphysics.py:457-465
    T accF[VARIANT == kOdeintEuler ? PPL : 1][6];
    if (VARIANT == kOdeintEuler) {
#pragma unroll
        for (int j = 0; j < PPL; ++j)
#pragma unroll
            for (int k = 0; k < 6; ++k) accF[j][k] = (T)0;
        record_state(0);
        if (FORCES) {
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                if (j < PPL - 1 || last_valid) {
                    const long long o = (long long)(j * 32 + lane) * 3;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { Fs_b[o + k] = (T)0; Ff_b[o + k] = (T)0; }
                }
            }
        }
    }

    const int n_steps = (VARIANT == kOdeintEuler) ? a.nT - 1 : a.nT;
    T uv = ctrl[0], uw = ctrl[1];

    for (int t = 0; t < n_steps; ++t) {
        // prefetch next controls
        T uv_n = uv, uw_n = uw;
        if (t + 1 < a.nT) { uv_n = ctrl[(t + 1) * 2]; uw_n = ctrl[(t + 1) * 2 + 1]; }

        // thrust direction: first column of R, normalised                       dphysics.py:237
        T hd[3];
        {
            const T nn = Mth<T>::sqrt_rn(s.R[0] * s.R[0] + s.R[3] * s.R[3] + s.R[6] * s.R[6]);
            const T inv = (T)1 / Mth<T>::fmax_(nn, (T)1e-6);
            hd[0] = s.R[0] * inv; hd[1] = s.R[3] * inv; hd[2] = s.R[6] * inv;
        }

        T nrm[PPL][3], sc[PPL], slip[PPL][3], arm[PPL][3];
        T C = (T)0;

#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const int slot = j * 32 + lane;
            const T px = tab.px[slot], py = tab.py[slot], pz = tab.pz[slot];
            // r = R p ; P = r + x ; Pd = v + w x r                              dphysics.py:200-204
            const T r0 = s.R[0] * px + s.R[1] * py + s.R[2] * pz;
            const T r1 = s.R[3] * px + s.R[4] * py + s.R[5] * pz;
            const T r2 = s.R[6] * px + s.R[7] * py + s.R[8] * pz;
            const T Px = r0 + s.x[0], Py = r1 + s.x[1], Pz = r2 + s.x[2];
            const T V0 = s.v[0] + (s.w[1] * r2 - s.w[2] * r1);
            const T V1 = s.v[1] + (s.w[2] * r0 - s.w[0] * r2);
            const T V2 = s.v[2] + (s.w[0] * r1 - s.w[1] * r0);
            // terrain height, normal, friction at the point                     dphysics.py:211-216
            T fx, fy;
            const Cell c = locate(Mth<T>::to_cells(Px, a.d_max, a.res, a.inv_res),
                                  Mth<T>::to_cells(Py, a.d_max, a.res, a.inv_res), H, W, fx, fy);
            const T z00 = ldg(zmap + c.k00), z10 = ldg(zmap + c.k10), z01 = ldg(zmap + c.k01), z11 = ldg(zmap + c.k11);
            const T m00 = ldg(fmap + c.k00), m10 = ldg(fmap + c.k10), m01 = ldg(fmap + c.k01), m11 = ldg(fmap + c.k11);
            const T zv = blend(fx, fy, z00, z10, z01, z11);
            const T mu = blend(fx, fy, m00, m10, m01, m11);
            const T ax = (z00 - z10) * a.inv_res;       // -dz/dx
            const T ay = (z00 - z01) * a.inv_res;       // -dz/dy
            const T inv_n = Mth<T>::rsqrt(ax * ax + ay * ay + (T)1);
            const T n0 = ax * inv_n, n1 = ay * inv_n, n2 = inv_n;
            // soft contact + spring-damper magnitude                            dphysics.py:220-232
            const T dh = Pz - zv;
            T cw = Mth<T>::contact(dh);
            if (j == PPL - 1 && !last_valid) cw = (T)0;
            C += cw;
            const T vn = V0 * n0 + V1 * n1 + V2 * n2;
            sc[j] = -(a.stiffness * dh + a.damping * vn) * cw;
            // tangential slip of the driven point w.r.t. the commanded track speed   dphysics.py:237-249
            const T tau = tab.driven[slot] * uv + tab.side[slot] * uw;
            const T d0 = mu * (tau * hd[0] - V0), d1 = mu * (tau * hd[1] - V1), d2 = mu * (tau * hd[2] - V2);
            const T dn = d0 * n0 + d1 * n1 + d2 * n2;
            slip[j][0] = d0 - dn * n0; slip[j][1] = d1 - dn * n1; slip[j][2] = d2 - dn * n2;
            nrm[j][0] = n0; nrm[j][1] = n1; nrm[j][2] = n2;
            arm[j][0] = r0; arm[j][1] = r1; arm[j][2] = r2;
        }

        C = warp_sum(C);
        const T invC = Mth<T>::rcp(C);

        T sum[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) sum[k] = (T)0;
        T nf_sum = (T)0, nf_sq = (T)0;
        T h = (T)0;
        if (VARIANT == kOdeintEuler) h = a.ts[t + 1] - a.ts[t];

        T* __restrict__ Fs_t = FORCES ? Fs_b + (long long)(VARIANT == kOdeintEuler ? t + 1 : t) * rowF : nullptr;
        T* __restrict__ Ff_t = FORCES ? Ff_b + (long long)(VARIANT == kOdeintEuler ? t + 1 : t) * rowF : nullptr;

#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const T f = sc[j] * invC;
            const T Fr0 = clampT(f * nrm[j][0], a.mg), Fr1 = clampT(f * nrm[j][1], a.mg), Fr2 = clampT(f * nrm[j][2], a.mg);
            const T Nf = Mth<T>::sqrt(Fr0 * Fr0 + Fr1 * Fr1 + Fr2 * Fr2);                   // dphysics.py:238
            const T Ft0 = clampT(Nf * slip[j][0], a.mg), Ft1 = clampT(Nf * slip[j][1], a.mg), Ft2 = clampT(Nf * slip[j][2], a.mg);
            const T F0 = Fr0 + Ft0, F1 = Fr1 + Ft1, F2 = Fr2 + Ft2;
            sum[0] += Fr0; sum[1] += Fr1; sum[2] += Fr2;
            sum[3] += Ft0; sum[4] += Ft1; sum[5] += Ft2;
            sum[6] += arm[j][1] * F2 - arm[j][2] * F1;                                       // dphysics.py:255
            sum[7] += arm[j][2] * F0 - arm[j][0] * F2;
            sum[8] += arm[j][0] * F1 - arm[j][1] * F0;
            if (COST) { nf_sum += Nf; nf_sq += Nf * Nf; }
            if (FORCES) {
                if (j < PPL - 1 || last_valid) {
                    const long long o = (long long)(j * 32 + lane) * 3;
                    if (VARIANT == kOdeintEuler) {
                        accF[j][0] += h * Fr0; accF[j][1] += h * Fr1; accF[j][2] += h * Fr2;
                        accF[j][3] += h * Ft0; accF[j][4] += h * Ft1; accF[j][5] += h * Ft2;
                        Fs_t[o + 0] = accF[j][0]; Fs_t[o + 1] = accF[j][1]; Fs_t[o + 2] = accF[j][2];
                        Ff_t[o + 0] = accF[j][3]; Ff_t[o + 1] = accF[j][4]; Ff_t[o + 2] = accF[j][5];
                    } else {
                        Fs_t[o + 0] = Fr0; Fs_t[o + 1] = Fr1; Fs_t[o + 2] = Fr2;
                        Ff_t[o + 0] = Ft0; Ff_t[o + 1] = Ft1; Ff_t[o + 2] = Ft2;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) sum[k] = warp_sum(sum[k]);

        if (COST) {
            // unbiased std over the N points of |F_spring|, then Welford over steps
            nf_sum = warp_sum(nf_sum); nf_sq = warp_sum(nf_sq);
            const T mean = nf_sum / (T)a.N;
            T var = (nf_sq - nf_sum * mean) / (T)(a.N - 1);
            var = Mth<T>::fmax_(var, (T)0);
            const T sd = Mth<T>::sqrt_rn(var);
            const T d = sd - cost_mean;
            cost_mean += d / (T)(t + 1);
            cost_m2 += d * (sd - cost_mean);
        }

        // angular / linear acceleration                                         dphysics.py:255-266
        T wd[3], vd[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            wd[r] = clampT(a.Iinv[r * 3 + 0] * sum[6] + a.Iinv[r * 3 + 1] * sum[7] + a.Iinv[r * 3 + 2] * sum[8], a.omega_max);
        }
        vd[0] = ((T)0 + sum[0] + sum[3]) * a.inv_mass;
        vd[1] = ((T)0 + sum[1] + sum[4]) * a.inv_mass;
        vd[2] = (-a.mg + sum[2] + sum[5]) * a.inv_mass;

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

    if (COST && lane == 0) {
        const int n = n_steps;
        a.cost[b] = n > 1 ? Mth<T>::sqrt_rn(cost_m2 / (T)(n - 1)) : (T)0;
    }
}

}  // namespace mfb
