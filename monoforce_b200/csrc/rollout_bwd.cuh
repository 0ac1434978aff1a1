// K2: hand-written adjoint of the fused rollout (replaces autograd through DPhysics.dphysics,
// SURVEY.md section 8 row A11).  One warp walks one trajectory backwards in time.
//
// For step t the pre-update state is read back from the recorded outputs (index t-1, or the
// initial state), the step is re-evaluated and its vector-Jacobian product applied:
//   pass A  phase 1 of the forward step (contact weights, un-normalised forces)  -> C = sum c
//   pass B  forward phase 2 + reverse of phase 2 per point (clamps, |F|, friction, torque arm)
//           -> per-point (G_bar, slip_bar, arm_bar) and the partial sum of C_bar
//   pass C  phase 1 again, now reversed: sampling weights, normals, soft contact, point
//           kinematics; scatters d/dz_grid, d/dfriction with atomics; partial sums of the
//           state adjoint
//   warp reductions, thrust-direction and controls gradients.
// Non-differentiable pieces follow torch semantics: `.long()` passes gradient through the
// fractional part only, clamp passes gradient inside [min, max], norm at 0 has zero gradient.
#pragma once
#include "rollout_bwd_args.cuh"

namespace mfb {

constexpr int kBwdWarps = 4;
#ifndef MFB_BWD_MINB
#define MFB_BWD_MINB 3
#endif

// ------------------------------------------------------------------------------------------
// map-gradient accumulation
// ------------------------------------------------------------------------------------------
// A contact point stays in the same map cell for several steps (<= 0.2 cells / step at 1 m/s,
// 5 cm cells), so each lane keeps, per point, a private shared-memory accumulator of the eight
// corner gradients (4 x height, 4 x friction) of its current cell and only flushes it to global
// memory when the point moves to another cell: ~10x fewer global atomics than one per step.
// The global target interleaves (d/dz, d/dfriction) per cell so a flush is two 16-byte or four
// 8-byte vector reductions (red.global.add.v4.f32 / .v2.f32, sm_90+).
template <typename T>
struct MapGradCache {
    Quad<T> z[kMaxPointsPerLane * 32];               // d/dz of the point's current cell corners (00, 10, 01, 11)
    Quad<T> m[kMaxPointsPerLane * 32];               // d/dfriction of the same corners
    int cell[kMaxPointsPerLane * 32];
};

__device__ __forceinline__ void red_pair(float* p, float a, float b) { atomicAdd(reinterpret_cast<float2*>(p), make_float2(a, b)); }
__device__ __forceinline__ void red_pair(double* p, double a, double b) { atomicAdd(p, a); atomicAdd(p + 1, b); }
__device__ __forceinline__ void red_quad(float* p, float a, float b, float c, float d) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
}
__device__ __forceinline__ void red_quad(double* p, double a, double b, double c, double d) {
    atomicAdd(p, a); atomicAdd(p + 1, b); atomicAdd(p + 2, c); atomicAdd(p + 3, d);
}

// g: interleaved (z, mu) gradient map; v = {z00, z10, z01, z11, m00, m10, m01, m11}
template <typename T>
__device__ __forceinline__ void scatter_corners(T* __restrict__ g, const Corners& c, const T* v) {
    // (k00, k01) and (k10, k11) are neighbours along y unless the flat clamp collapsed them
    if (c.k01 == c.k00 + 1 && (c.k00 & 1) == 0) red_quad(g + 2 * (long long)c.k00, v[0], v[4], v[2], v[6]);
    else { red_pair(g + 2 * (long long)c.k00, v[0], v[4]); red_pair(g + 2 * (long long)c.k01, v[2], v[6]); }
    if (c.k11 == c.k10 + 1 && (c.k10 & 1) == 0) red_quad(g + 2 * (long long)c.k10, v[1], v[5], v[3], v[7]);
    else { red_pair(g + 2 * (long long)c.k10, v[1], v[5]); red_pair(g + 2 * (long long)c.k11, v[3], v[7]); }
}

// adds (z, mu) gradient scratch into the caller's separate gradient maps
template <typename T>
__global__ void scatter_map_grads_kernel(const T* __restrict__ g2, T* __restrict__ g_z, T* __restrict__ g_mu, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (g_z) g_z[i] += g2[2 * i];
        if (g_mu) g_mu[i] += g2[2 * i + 1];
    }
}

template <typename T>
__device__ __forceinline__ T gate(T grad, T val, T lim) {   // backward of clamp(val, -lim, lim): passes inside [-lim, lim]
    return (fabs(val) <= lim) ? grad : (T)0;
}

// adjoint of E(w) = I + sn K + c1 (k k^T - |k|^2 I) (see rodrigues_right); accumulates into wb
template <typename T>
__device__ __forceinline__ void rodrigues_right_bwd(const T* w, T dt, const T* Eb, T* wb) {
    T th, inv;         // |w| and 1 / max(|w|, 1e-6), the same way the forward forms them (one rsqrt in fp32)
    Mth<T>::norm_and_inv(w[0] * w[0] + w[1] * w[1] + w[2] * w[2], &th, &inv);
    const T k0 = w[0] * inv, k1 = w[1] * inv, k2 = w[2] * inv;
    T sn, c1;
    sin_versin(th * dt, &sn, &c1);
    const T cs = (T)1 - c1;
    const T kk = k0 * k0 + k1 * k1 + k2 * k2;
    const T tr = Eb[0] + Eb[4] + Eb[8];
    const T a0 = Eb[7] - Eb[5], a1 = Eb[2] - Eb[6], a2 = Eb[3] - Eb[1];   // axial part of Eb
    const T sn_b = k0 * a0 + k1 * a1 + k2 * a2;
    // k^T Eb k
    const T Ek0 = Eb[0] * k0 + Eb[1] * k1 + Eb[2] * k2;
    const T Ek1 = Eb[3] * k0 + Eb[4] * k1 + Eb[5] * k2;
    const T Ek2 = Eb[6] * k0 + Eb[7] * k1 + Eb[8] * k2;
    const T Etk0 = Eb[0] * k0 + Eb[3] * k1 + Eb[6] * k2;
    const T Etk1 = Eb[1] * k0 + Eb[4] * k1 + Eb[7] * k2;
    const T Etk2 = Eb[2] * k0 + Eb[5] * k1 + Eb[8] * k2;
    const T c1_b = (k0 * Ek0 + k1 * Ek1 + k2 * Ek2) - kk * tr;
    T kb0 = sn * a0 + c1 * (Ek0 + Etk0 - (T)2 * tr * k0);
    T kb1 = sn * a1 + c1 * (Ek1 + Etk1 - (T)2 * tr * k1);
    T kb2 = sn * a2 + c1 * (Ek2 + Etk2 - (T)2 * tr * k2);
    T th_b = dt * (sn_b * cs + c1_b * sn);
    wb[0] += kb0 * inv; wb[1] += kb1 * inv; wb[2] += kb2 * inv;
    const T inv_b = kb0 * w[0] + kb1 * w[1] + kb2 * w[2];
    if (th >= (T)1e-6) th_b -= inv_b * inv * inv;
    if (th > (T)0) {
        const T sc = th >= (T)1e-6 ? th_b * inv : th_b / th;
        wb[0] += sc * w[0]; wb[1] += sc * w[1]; wb[2] += sc * w[2];
    }
}

template <typename T, int PPL, int VARIANT, bool HAS_FGRAD, bool JOINTS = false>
__global__ void __launch_bounds__(kBwdWarps * 32, (sizeof(T) == 4 && PPL <= 7) ? MFB_BWD_MINB : 1)
rollout_bwd_kernel(const RolloutArgs<T> a, const AdjointArgs<T> g) {
    __shared__ PointTable<T> tab;
    fill_point_table(tab, a, PPL * 32);
    __syncthreads();

    const int lane = lane_id();
    const int b = blockIdx.x * kBwdWarps + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const int n_last = a.N - (PPL - 1) * 32;
    const bool last_valid = lane < n_last;

    const long long mi = b / a.map_group;                       // map of this trajectory (groups of consecutive trajectories share one)
    const T* __restrict__ zmap = a.z + mi * a.map_stride;
    const T* __restrict__ fmap = a.mu + mi * a.map_stride;
    const T* __restrict__ cells = a.cells + mi * a.cell_stride;
    T* __restrict__ gmap = g.g_maps ? g.g_maps + mi * g.g_maps_stride : nullptr;
    extern __shared__ __align__(16) unsigned char cache_raw[];
    MapGradCache<T>& wc = reinterpret_cast<MapGradCache<T>*>(cache_raw)[threadIdx.x >> 5];
    if (gmap) {
#pragma unroll
        for (int j = 0; j < PPL; ++j) wc.cell[j * 32 + lane] = -1;
    }
    const T* __restrict__ ctrl = a.controls + (long long)b * a.nT * 2;
    const int H = a.H, W = a.W;
    const long long rowF = (long long)a.N * 3;

    const T* __restrict__ Xs_b = a.Xs + (long long)b * a.nT * 3;
    const T* __restrict__ Xd_b = a.Xds + (long long)b * a.nT * 3;
    const T* __restrict__ Rs_b = a.Rs + (long long)b * a.nT * 9;
    const T* __restrict__ Om_b = a.Oms + (long long)b * a.nT * 3;

    // pre-update state of step t: recorded index t-1 (step loop) / t (odeint), else the initial state
    auto load_state = [&](Body<T>& s, int idx) {
        if (idx < 0) {
            load_body(s, a, b);
            s.x[2] = a.x0z[b];
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) { s.v[i] = Xd_b[idx * 3 + i]; s.w[i] = Om_b[idx * 3 + i]; }
#pragma unroll
            for (int i = 0; i < 9; ++i) s.R[i] = Rs_b[idx * 9 + i];
            // undo Xs = x + R[:,2] delta_h                                      dphysics.py:587-589
            s.x[0] = Xs_b[idx * 3 + 0] - s.R[2] * a.delta_h;
            s.x[1] = Xs_b[idx * 3 + 1] - s.R[5] * a.delta_h;
            s.x[2] = Xs_b[idx * 3 + 2] - s.R[8] * a.delta_h;
        }
    };

    // adjoint of the carried state
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

    // odeint: adjoint of the per-point force accumulators (suffix sums of the incoming force grads)
    T accb[(VARIANT == kOdeintEuler && HAS_FGRAD) ? PPL : 1][6];
    if (VARIANT == kOdeintEuler && HAS_FGRAD) {
#pragma unroll
        for (int j = 0; j < PPL; ++j)
#pragma unroll
            for (int k = 0; k < 6; ++k) accb[j][k] = (T)0;
    }

    const int n_steps = (VARIANT == kOdeintEuler) ? a.nT - 1 : a.nT;
    if (VARIANT == kOdeintEuler && g.g_controls && lane == 0) {
        // the last control sample is never used by the fixed-grid solver
        g.g_controls[((long long)b * a.nT + a.nT - 1) * 2 + 0] = (T)0;
        g.g_controls[((long long)b * a.nT + a.nT - 1) * 2 + 1] = (T)0;
    }

    if (JOINTS && VARIANT == kOdeintEuler && g.g_joint_angles && lane < 4)
        g.g_joint_angles[((long long)b * a.nT + a.nT - 1) * 4 + lane] = (T)0;

    T w_post[3] = {0, 0, 0};     // angular velocity after step t (== pre-state of step t+1)
    if (n_steps > 0) {
        const int last = (VARIANT == kOdeintEuler) ? n_steps : n_steps - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) w_post[i] = Om_b[last * 3 + i];
    }

    for (int t = n_steps - 1; t >= 0; --t) {
        Body<T> s;
        load_state(s, VARIANT == kOdeintEuler ? t : t - 1);
        const T uv = ctrl[t * 2], uw = ctrl[t * 2 + 1];
        const int rec = (VARIANT == kOdeintEuler) ? t + 1 : t;     // record index written by this step
        add_output_grads(rec);

        T h = a.dt;
        if (VARIANT == kOdeintEuler) h = a.ts[t + 1] - a.ts[t];

        // ---------------- reverse of the state update ----------------
        T wd_b[3], vd_b[3];
        if (VARIANT == kStepLoop) {
            // R' = R E(w'):  E_bar = R^T R'_bar ; R_bar = R'_bar E^T                     dphysics.py:290-324
            T Eb[9], E[9];
            {
                // rebuild E from w' (same formula as rodrigues_right)
                const T th = Mth<T>::sqrt_rn(w_post[0] * w_post[0] + w_post[1] * w_post[1] + w_post[2] * w_post[2]);
                const T inv = (T)1 / Mth<T>::fmax_(th, (T)1e-6);
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
            // w_bar += h * axial(R'_bar R^T) ; R_bar = R'_bar + h [w]x^T R'_bar
            T M[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    M[r * 3 + c] = Rb[r * 3 + 0] * s.R[c * 3 + 0] + Rb[r * 3 + 1] * s.R[c * 3 + 1] + Rb[r * 3 + 2] * s.R[c * 3 + 2];
            T Rn[9];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                // [w]x^T = -[w]x :  (-w x col)
                Rn[0 + c] = Rb[0 + c] - h * (s.w[1] * Rb[6 + c] - s.w[2] * Rb[3 + c]);
                Rn[3 + c] = Rb[3 + c] - h * (s.w[2] * Rb[0 + c] - s.w[0] * Rb[6 + c]);
                Rn[6 + c] = Rb[6 + c] - h * (s.w[0] * Rb[3 + c] - s.w[1] * Rb[0 + c]);
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) Rb[i] = Rn[i];
            // <Rb', h [dw]x R> = h dw . axial(M),  axial(M) = (M21 - M12, M02 - M20, M10 - M01)
            const T ax0 = M[7] - M[5], ax1 = M[2] - M[6], ax2 = M[3] - M[1];
            // v_bar gets h x_bar (x' = x + h v) AFTER vd_b was formed from the old v_bar
            vb[0] += h * xb[0]; vb[1] += h * xb[1]; vb[2] += h * xb[2];
            wb[0] += h * ax0; wb[1] += h * ax1; wb[2] += h * ax2;
        }

        // ---------------- clamp mask of the angular acceleration ----------------
        // active  <=>  w' == fma(+-omega_max, h, w) bit-for-bit (the forward uses the same fma)
        T tq_b[3], wdm[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const bool hi = fma(a.omega_max, h, s.w[i]) == w_post[i];
            const bool lo = fma(-a.omega_max, h, s.w[i]) == w_post[i];
            wdm[i] = (hi || lo) ? (T)0 : wd_b[i];
        }
        const T fs_b0 = vd_b[0] * a.inv_mass, fs_b1 = vd_b[1] * a.inv_mass, fs_b2 = vd_b[2] * a.inv_mass;

        StepFrame<T> f;
        make_frame(f, s, uv, uw, a.d_max, a.res, a.inv_res);
        const T hd_norm = Mth<T>::sqrt_rn(s.R[0] * s.R[0] + s.R[3] * s.R[3] + s.R[6] * s.R[6]);

        // ---------------- pass A: phase 1 forward ----------------
        T nrm[PPL][3], sc[PPL], slip[PPL][3], arm[PPL][3];
        T C = (T)0;
        // moving flippers: lane i < 4 owns (cos, sin) of joint angle i; second moments of the articulated body
        T jc = (T)1, js = (T)0, mom[6] = {0, 0, 0, 0, 0, 0};
        if (JOINTS) {
            const T ang = lane < 4 ? a.joint_angles[((long long)b * a.nT + t) * 4 + lane] : (T)0;
            Mth<T>::sincos(ang, &js, &jc);
        }
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const int slot = j * 32 + lane;
            T px = tab.px[slot], py = tab.py[slot], pz = tab.pz[slot];
            if (JOINTS) {
                articulate_point(px, py, pz, tab.part[slot], jc, js, a.joint_pos);
                if ((j < PPL - 1) || last_valid) {
                    mom[0] += py * py + pz * pz; mom[1] += px * px + pz * pz; mom[2] += px * px + py * py;
                    mom[3] += px * py; mom[4] += px * pz; mom[5] += py * pz;
                }
            }
            PointEval<T> e;
            eval_point(e, f, px, py, pz, tab.driven[slot], tab.side[slot],
                       (j < PPL - 1) || last_valid, cells, zmap, fmap, H, W, a.inv_res, a.stiffness, a.damping);
            C += e.cw;
            sc[j] = e.sp * e.cw;
#pragma unroll
            for (int k = 0; k < 3; ++k) { slip[j][k] = e.sl[k]; nrm[j][k] = e.rec[4 + k]; arm[j][k] = e.r[k]; }
        }
        T Iinv[9];
        if (JOINTS) {
            T red[8] = {C, mom[0], mom[1], mom[2], mom[3], mom[4], mom[5], (T)0};
            warp_sum8(red, lane);
            C = red[0];
            const T mp = a.mass / (T)a.N;
#pragma unroll
            for (int k = 0; k < 6; ++k) mom[k] = red[1 + k] * mp;
            invert_inertia(mom, Iinv);
        } else {
            C = warp_sum(C);
#pragma unroll
            for (int k = 0; k < 9; ++k) Iinv[k] = a.Iinv[k];
        }
        const T invC = Mth<T>::rcp(C);
        // tq_bar = Iinv^T (masked wd_bar)
#pragma unroll
        for (int c = 0; c < 3; ++c) tq_b[c] = Iinv[0 + c] * wdm[0] + Iinv[3 + c] * wdm[1] + Iinv[6 + c] * wdm[2];

        // ---------------- pass B: phase 2 forward + its reverse ----------------
        // reuses the per-point registers: nrm -> G_bar, slip -> slip_bar, arm -> arm_bar
        T Cb_part = (T)0, tq_p[3] = {0, 0, 0};
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const T f = sc[j] * invC;
            const T G0 = f * nrm[j][0], G1 = f * nrm[j][1], G2 = f * nrm[j][2];
            const T Fr0 = clampT(G0, a.mg), Fr1 = clampT(G1, a.mg), Fr2 = clampT(G2, a.mg);
            const T Nf2 = Fr0 * Fr0 + Fr1 * Fr1 + Fr2 * Fr2;
            const T Nf_inv = Nf2 > (T)0 ? Mth<T>::rsqrt(Nf2) : (T)0;
            const T Nf = Nf2 * Nf_inv;
            const T Hh0 = Nf * slip[j][0], Hh1 = Nf * slip[j][1], Hh2 = Nf * slip[j][2];
            const T Ft0 = clampT(Hh0, a.mg), Ft1 = clampT(Hh1, a.mg), Ft2 = clampT(Hh2, a.mg);
            const T F0 = Fr0 + Ft0, F1 = Fr1 + Ft1, F2 = Fr2 + Ft2;
            const T r0 = arm[j][0], r1 = arm[j][1], r2 = arm[j][2];
            // d/dF of torque = sum r x F :  F_bar += tq_bar x r ;  r_bar += F x tq_bar
            const T c0 = tq_b[1] * r2 - tq_b[2] * r1, c1 = tq_b[2] * r0 - tq_b[0] * r2, c2 = tq_b[0] * r1 - tq_b[1] * r0;
            T Frb0 = fs_b0 + c0, Frb1 = fs_b1 + c1, Frb2 = fs_b2 + c2;
            T Ftb0 = Frb0, Ftb1 = Frb1, Ftb2 = Frb2;
            if (HAS_FGRAD) {
                const bool ok = (j < PPL - 1 || last_valid);
                const long long o = ((long long)b * a.nT + rec) * rowF + (long long)(j * 32 + lane) * 3;
                T e[6] = {0, 0, 0, 0, 0, 0};
                if (ok) {
                    if (g.g_Fs) { e[0] = g.g_Fs[o]; e[1] = g.g_Fs[o + 1]; e[2] = g.g_Fs[o + 2]; }
                    if (g.g_Ff) { e[3] = g.g_Ff[o]; e[4] = g.g_Ff[o + 1]; e[5] = g.g_Ff[o + 2]; }
                }
                if (VARIANT == kOdeintEuler) {
                    // recorded forces are time integrals: A' = A + h F
#pragma unroll
                    for (int k = 0; k < 6; ++k) accb[j][k] += e[k];
                    Frb0 += h * accb[j][0]; Frb1 += h * accb[j][1]; Frb2 += h * accb[j][2];
                    Ftb0 += h * accb[j][3]; Ftb1 += h * accb[j][4]; Ftb2 += h * accb[j][5];
                } else {
                    Frb0 += e[0]; Frb1 += e[1]; Frb2 += e[2];
                    Ftb0 += e[3]; Ftb1 += e[4]; Ftb2 += e[5];
                }
            }
            if (JOINTS) {      // forward torque, needed for d wd / d Iinv
                tq_p[0] += r1 * F2 - r2 * F1; tq_p[1] += r2 * F0 - r0 * F2; tq_p[2] += r0 * F1 - r1 * F0;
            }
            arm[j][0] = F1 * tq_b[2] - F2 * tq_b[1];
            arm[j][1] = F2 * tq_b[0] - F0 * tq_b[2];
            arm[j][2] = F0 * tq_b[1] - F1 * tq_b[0];
            // F_friction = clamp(Nf * slip)
            const T Hb0 = gate(Ftb0, Hh0, a.mg), Hb1 = gate(Ftb1, Hh1, a.mg), Hb2 = gate(Ftb2, Hh2, a.mg);
            const T Nf_b = Hb0 * slip[j][0] + Hb1 * slip[j][1] + Hb2 * slip[j][2];
            slip[j][0] = Nf * Hb0; slip[j][1] = Nf * Hb1; slip[j][2] = Nf * Hb2;
            // Nf = |F_spring|
            {
                const T k = Nf_b * Nf_inv;          // d|F|/dF = F / |F| (0 at the origin, like torch.norm)
                Frb0 += k * Fr0; Frb1 += k * Fr1; Frb2 += k * Fr2;
            }
            // F_spring = clamp(f n)
            const T Gb0 = gate(Frb0, G0, a.mg), Gb1 = gate(Frb1, G1, a.mg), Gb2 = gate(Frb2, G2, a.mg);
            const T f_b = Gb0 * nrm[j][0] + Gb1 * nrm[j][1] + Gb2 * nrm[j][2];
            Cb_part += f_b * f;
            nrm[j][0] = Gb0; nrm[j][1] = Gb1; nrm[j][2] = Gb2;
        }
        T C_b, mom_b[6] = {0, 0, 0, 0, 0, 0};
        if (JOINTS) {
            T red[8] = {Cb_part, tq_p[0], tq_p[1], tq_p[2], (T)0, (T)0, (T)0, (T)0};
            warp_sum8(red, lane);
            C_b = -red[0] * invC;
            // wd = clamp(Iinv tq):  Iinv_bar = m tq^T  ->  I_bar = -Iinv^T Iinv_bar Iinv^T = -(tq_bar)(Iinv tq)^T
            T u[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) u[r] = Iinv[r * 3 + 0] * red[1] + Iinv[r * 3 + 1] * red[2] + Iinv[r * 3 + 2] * red[3];
            const T mp = a.mass / (T)a.N;
            // I = mp [[m0, -m3, -m4], [-m3, m1, -m5], [-m4, -m5, m2]]
            mom_b[0] = -mp * tq_b[0] * u[0];
            mom_b[1] = -mp * tq_b[1] * u[1];
            mom_b[2] = -mp * tq_b[2] * u[2];
            mom_b[3] = mp * (tq_b[0] * u[1] + tq_b[1] * u[0]);
            mom_b[4] = mp * (tq_b[0] * u[2] + tq_b[2] * u[0]);
            mom_b[5] = mp * (tq_b[1] * u[2] + tq_b[2] * u[1]);
        } else {
            C_b = -warp_sum(Cb_part) * invC;
        }

        // ---------------- pass C: phase 1 again, reversed ----------------
        T acc[24];
#pragma unroll
        for (int k = 0; k < 24; ++k) acc[k] = (T)0;
        // acc: 0-2 x_bar, 3-5 v_bar, 6-8 w_bar, 9-17 R_bar, 18-20 hd_bar, 21-22 controls
        T jg[4] = {0, 0, 0, 0};      // d loss / d joint angle, per driving part
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const int slot = j * 32 + lane;
            const bool ok = (j < PPL - 1 || last_valid);
            T px = tab.px[slot], py = tab.py[slot], pz = tab.pz[slot];
            if (JOINTS) articulate_point(px, py, pz, tab.part[slot], jc, js, a.joint_pos);
            const T drv = tab.driven[slot], side = tab.side[slot];
            PointEval<T> e;
            eval_point(e, f, px, py, pz, drv, side, ok, cells, zmap, fmap, H, W, a.inv_res, a.stiffness, a.damping);
            const T n0 = e.rec[4], n1 = e.rec[5], n2 = e.rec[6];
            const T fx = e.fx, fy = e.fy;

            const T Gb0 = nrm[j][0], Gb1 = nrm[j][1], Gb2 = nrm[j][2];
            const T fo = e.sp * e.cw * invC;
            const T f_b = Gb0 * n0 + Gb1 * n1 + Gb2 * n2;
            T nb0 = fo * Gb0, nb1 = fo * Gb1, nb2 = fo * Gb2;
            const T sc_b = f_b * invC;
            const T sp_b = sc_b * e.cw;
            const T cw_b = sc_b * e.sp + C_b;
            // slip = d - dn n ; dn = d . n
            const T sb0 = slip[j][0], sb1 = slip[j][1], sb2 = slip[j][2];
            const T dn_b = -(sb0 * n0 + sb1 * n1 + sb2 * n2);
            nb0 += -e.dn * sb0 + dn_b * e.d[0]; nb1 += -e.dn * sb1 + dn_b * e.d[1]; nb2 += -e.dn * sb2 + dn_b * e.d[2];
            const T db0 = sb0 + dn_b * n0, db1 = sb1 + dn_b * n1, db2 = sb2 + dn_b * n2;
            // d = mu e ; e = tau hd - V
            const T mu_b = db0 * e.e[0] + db1 * e.e[1] + db2 * e.e[2];
            const T eb0 = e.mu * db0, eb1 = e.mu * db1, eb2 = e.mu * db2;
            const T tau_b = eb0 * f.hd[0] + eb1 * f.hd[1] + eb2 * f.hd[2];
            acc[18] += e.tau * eb0; acc[19] += e.tau * eb1; acc[20] += e.tau * eb2;
            acc[21] += drv * tau_b; acc[22] += side * tau_b;
            T Vb0 = -eb0, Vb1 = -eb1, Vb2 = -eb2;
            // sp = -(k dh + beta vn) ; vn = V . n
            T dh_b = -a.stiffness * sp_b;
            const T vn_b = -a.damping * sp_b;
            Vb0 += vn_b * n0; Vb1 += vn_b * n1; Vb2 += vn_b * n2;
            nb0 += vn_b * e.V[0]; nb1 += vn_b * e.V[1]; nb2 += vn_b * e.V[2];
            // cw = sigmoid(-10 dh)
            dh_b += cw_b * ((T)-10 * e.cw * ((T)1 - e.cw));
            // dh = Pz - zv
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
            // d v / d fx = cx + fy cxy ; d v / d fy = cy + fx cxy
            const T fx_b = zv_b * e.dz_dfx + mu_b * (e.rec[10] + fy * e.rec[11]);
            const T fy_b = zv_b * (e.rec[1] + fx * e.rec[3]) + mu_b * (e.rec[9] + fx * e.rec[11]);
            if (ok && gmap) {
                const T cg[8] = {zv_b * w00 + (ax_b + ay_b), zv_b * w10 - ax_b, zv_b * w01 - ay_b, zv_b * w11,
                                 mu_b * w00, mu_b * w10, mu_b * w01, mu_b * w11};
                if (e.cell >= 0) {
                    const int cur = wc.cell[slot];
                    Quad<T> qz = quad_load(&wc.z[slot]), qm = quad_load(&wc.m[slot]);
                    if (cur == e.cell) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) { qz.v[k] += cg[k]; qm.v[k] += cg[4 + k]; }
                    } else {
                        if (cur >= 0) {
                            const T old[8] = {qz.v[0], qz.v[1], qz.v[2], qz.v[3], qm.v[0], qm.v[1], qm.v[2], qm.v[3]};
                            scatter_corners(gmap, on_map_corners(cur, H, W), old);
                        }
                        wc.cell[slot] = e.cell;
#pragma unroll
                        for (int k = 0; k < 4; ++k) { qz.v[k] = cg[k]; qm.v[k] = cg[4 + k]; }
                    }
                    quad_store(&wc.z[slot], qz); quad_store(&wc.m[slot], qm);
                } else {
                    const T ggx = e.r[0] * a.inv_res + f.ox, ggy = e.r[1] * a.inv_res + f.oy;
                    scatter_corners(gmap, flat_corners((long long)ggx, (long long)ggy, H, W), cg);
                }
            }
            // grid coordinate -> world point
            const T Pb0 = fx_b * a.inv_res, Pb1 = fy_b * a.inv_res, Pb2 = dh_b;
            // V = v + w x r ; P = r + x ; r = R p
            T rb0 = arm[j][0] + Pb0 + (Vb1 * s.w[2] - Vb2 * s.w[1]);
            T rb1 = arm[j][1] + Pb1 + (Vb2 * s.w[0] - Vb0 * s.w[2]);
            T rb2 = arm[j][2] + Pb2 + (Vb0 * s.w[1] - Vb1 * s.w[0]);
            if (!ok) { rb0 = rb1 = rb2 = (T)0; Vb0 = Vb1 = Vb2 = (T)0; }
            const T okf = ok ? (T)1 : (T)0;
            const T r0 = e.r[0], r1 = e.r[1], r2 = e.r[2];
            acc[0] += okf * Pb0; acc[1] += okf * Pb1; acc[2] += okf * Pb2;
            acc[3] += Vb0; acc[4] += Vb1; acc[5] += Vb2;
            acc[6] += r1 * Vb2 - r2 * Vb1; acc[7] += r2 * Vb0 - r0 * Vb2; acc[8] += r0 * Vb1 - r1 * Vb0;
            acc[9] += rb0 * px;  acc[10] += rb0 * py; acc[11] += rb0 * pz;
            acc[12] += rb1 * px; acc[13] += rb1 * py; acc[14] += rb1 * pz;
            acc[15] += rb2 * px; acc[16] += rb2 * py; acc[17] += rb2 * pz;
            if (JOINTS) {
                const int part = tab.part[slot];
                if (ok && part >= 0) {
                    // p'_bar = R^T r_bar + d(inertia moments)/dp' ; angle_bar += p'_bar . dp'/dangle
                    const T pbx = s.R[0] * rb0 + s.R[3] * rb1 + s.R[6] * rb2 +
                                  (T)2 * px * (mom_b[1] + mom_b[2]) + mom_b[3] * py + mom_b[4] * pz;
                    const T pbz = s.R[2] * rb0 + s.R[5] * rb1 + s.R[8] * rb2 +
                                  (T)2 * pz * (mom_b[0] + mom_b[1]) + mom_b[4] * px + mom_b[5] * py;
                    const T ga = pbx * (pz - a.joint_pos[part * 3 + 2]) - pbz * (px - a.joint_pos[part * 3 + 0]);
                    jg[0] += part == 0 ? ga : (T)0; jg[1] += part == 1 ? ga : (T)0;
                    jg[2] += part == 2 ? ga : (T)0; jg[3] += part == 3 ? ga : (T)0;
                }
            }
        }
        warp_sum8(acc, lane); warp_sum8(acc + 8, lane); warp_sum8(acc + 16, lane);
        if (JOINTS) {
            T red[8] = {jg[0], jg[1], jg[2], jg[3], (T)0, (T)0, (T)0, (T)0};
            warp_sum8(red, lane);
            if (g.g_joint_angles && lane < 4)
                g.g_joint_angles[((long long)b * a.nT + t) * 4 + lane] = lane == 0 ? red[0] : lane == 1 ? red[1] : lane == 2 ? red[2] : red[3];
        }

        // fold the per-point sums into the state adjoint (pre-update state)
#pragma unroll
        for (int i = 0; i < 3; ++i) { xb[i] += acc[i]; vb[i] += acc[3 + i]; wb[i] += acc[6 + i]; }
#pragma unroll
        for (int i = 0; i < 9; ++i) Rb[i] += acc[9 + i];
        // hd = R[:,0] / max(|R[:,0]|, eps)
        {
            T a0, a1, a2;
            if (hd_norm >= (T)1e-6) {
                const T dot = f.hd[0] * acc[18] + f.hd[1] * acc[19] + f.hd[2] * acc[20];
                const T inv = (T)1 / hd_norm;
                a0 = (acc[18] - f.hd[0] * dot) * inv; a1 = (acc[19] - f.hd[1] * dot) * inv; a2 = (acc[20] - f.hd[2] * dot) * inv;
            } else {
                a0 = acc[18] * (T)1e6; a1 = acc[19] * (T)1e6; a2 = acc[20] * (T)1e6;
            }
            Rb[0] += a0; Rb[3] += a1; Rb[6] += a2;
        }
        if (g.g_controls && lane == 0) {
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
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const int slot = j * 32 + lane;
            const bool ok = (j < PPL - 1 || last_valid);
            const T px = tab.px[slot], py = tab.py[slot], pz = tab.pz[slot];
            PointEval<T> e;
            eval_point(e, f, px, py, pz, (T)0, (T)0, ok, cells, zmap, fmap, H, W, a.inv_res, a.stiffness, a.damping);
            const T fx = e.fx, fy = e.fy;
            const T gx = (T)1 - fx, gy = (T)1 - fy;
            if (ok) {
                if (gmap) {
                    Corners c;
                    if (e.cell >= 0) {
                        c = on_map_corners(e.cell, H, W);
                    } else {
                        const T ggx = e.r[0] * a.inv_res + f.ox, ggy = e.r[1] * a.inv_res + f.oy;
                        c = flat_corners((long long)ggx, (long long)ggy, H, W);
                    }
                    const T cg[8] = {zb * gx * gy, zb * gx * fy, zb * fx * gy, zb * fx * fy, (T)0, (T)0, (T)0, (T)0};
                    scatter_corners(gmap, c, cg);
                }
                const T Pb0 = zb * e.dz_dfx * a.inv_res;
                const T Pb1 = zb * (e.rec[1] + fx * e.rec[3]) * a.inv_res;
                sx += Pb0; sy += Pb1;
                rr[0] += Pb0 * px; rr[1] += Pb0 * py; rr[2] += Pb0 * pz;
                rr[3] += Pb1 * px; rr[4] += Pb1 * py; rr[5] += Pb1 * pz;
            }
        }
        if (gmap) {
            // drain the lane-private accumulators
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                const int slot = j * 32 + lane;
                const int cur = wc.cell[slot];
                if (cur >= 0) {
                    const Quad<T> qz = quad_load(&wc.z[slot]), qm = quad_load(&wc.m[slot]);
                    const T old[8] = {qz.v[0], qz.v[1], qz.v[2], qz.v[3], qm.v[0], qm.v[1], qm.v[2], qm.v[3]};
                    scatter_corners(gmap, on_map_corners(cur, H, W), old);
                }
            }
        }
        sx = warp_sum(sx); sy = warp_sum(sy);
#pragma unroll
        for (int k = 0; k < 6; ++k) rr[k] = warp_sum(rr[k]);
        if (lane == 0) {
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
