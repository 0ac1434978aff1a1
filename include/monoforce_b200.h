/* monoforce_b200 C ABI  --  the drop-in boundary of the B200-native DPhysics rollout.
 *
 * The reference (ctu-vras/monoforce) is pure Python and has no FFI; the interface these
 * entry points replace is the tensor-level contract of
 *     DPhysics.forward / DPhysics.dphysics            monoforce/src/monoforce/models/traj_predictor/dphysics.py:596-605, :530-594
 *     its implicit autograd backward                  
 *     the per-trajectory path cost                    monoforce_ros/nodes/monoforce_node.py:91
 * One call == one whole rollout of B trajectories x T steps.  All pointers are plain device
 * (or, for the *_host entry points, host) pointers to dense row-major arrays of the scalar
 * type selected by `dtype`; no torch types cross this boundary.  monoforce_b200/_lib.py is
 * the ctypes binding; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Every function returns 0 on success and a negative mfb_status otherwise;
 * mfb_last_error() returns a thread-local, human-readable description of the last failure.
 */
#ifndef MONOFORCE_B200_H_
#define MONOFORCE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFB_ABI_VERSION 3

typedef enum mfb_status {
    MFB_OK = 0,
    MFB_ERR_INVALID_ARGUMENT = -1,
    MFB_ERR_UNSUPPORTED = -2,
    MFB_ERR_CUDA = -3
} mfb_status;

typedef enum mfb_dtype { MFB_F32 = 0, MFB_F64 = 1 } mfb_dtype;

/* Integrator variants of the reference. */
typedef enum mfb_variant {
    MFB_STEP_LOOP = 0,     /* DPhysics.dynamics, semi-implicit Euler + Rodrigues     dphysics.py:467-497 */
    MFB_ODEINT_EULER = 1   /* DPhysics.dynamics_odeint, torchdiffeq fixed-grid Euler dphysics.py:499-528 */
} mfb_variant;

/* Problem description: sizes + the constants the reference reads from DPhysConfig
 * (dphys_config.py:77-153).  Doubles are converted to the working precision inside the
 * library exactly the way torch converts Python scalars. */
typedef struct mfb_rollout_desc {
    int32_t B;            /* trajectories                                             */
    int32_t T;            /* recorded time indices (N_ts, dphysics.py:573)            */
    int32_t N;            /* contact points (<= 256)                                  */
    int32_t H, W;         /* height-map size; the reference requires H == W           */
    int32_t n_tracks;     /* 2 or 4 driving parts (dphysics.py:75-104)                */
    int32_t variant;      /* mfb_variant                                              */
    int32_t traj_per_map; /* with map_stride != 0: consecutive trajectories that read the same map (trajectory b uses map
                             b / traj_per_map: one map per scene, many control sequences per scene - the shooting call of
                             monoforce_ros/nodes/monoforce_node.py:75,156 batched over scenes); 0 or 1 = one map per trajectory */
    int64_t map_stride;   /* elements between consecutive maps; 0 = all trajectories share one map */
    double mass, gravity, stiffness, damping, grid_res, d_max, dt, omega_max;
    double robot_Ly;      /* robot_size[1] (track gauge), dphys_config.py:43          */
    double I_inv[9];      /* inverse of the body-frame inertia tensor, row-major (dphysics.py:152-153) */
    double joint_positions[12]; /* (4,3) pivots of the driving parts (dphys_config.py:99-104); used with joint_angles */
} mfb_rollout_desc;

/* Inputs and outputs of the forward rollout.  Shapes follow DPhysics.forward. */
typedef struct mfb_rollout_buffers {
    /* inputs */
    const void* z_grid;     /* (n_maps, H, W) height map(s), n_maps = 1, B or ceil(B / traj_per_map) */
    const void* friction;   /* (n_maps, H, W) friction map(s)                         */
    const void* controls;   /* (B, T, 2) (v, w) per step                              */
    const void* x0;         /* (B, 3)   initial position  (z component is overwritten by the start-height snap, see x0z) */
    const void* xd0;        /* (B, 3)   initial velocity                              */
    const void* R0;         /* (B, 3, 3) initial rotation                             */
    const void* omega0;     /* (B, 3)   initial angular velocity                      */
    const void* points;     /* (N, 3)   body-frame contact points                     */
    const int32_t* part_id; /* (N,)     driving part of each point or -1              */
    const void* ts;         /* (T,)     solver time grid; required for MFB_ODEINT_EULER, else may be NULL */
    const void* joint_angles; /* (B, T, 4) flipper angles of the 4 driving parts, or NULL for static geometry
                               (DPhysics.update_joints, dphysics.py:326-358) */
    /* outputs */
    void* Xs;               /* (B, T, 3)                                              */
    void* Xds;              /* (B, T, 3)                                              */
    void* Rs;               /* (B, T, 3, 3)                                           */
    void* Omegas;           /* (B, T, 3)                                              */
    void* F_springs;        /* (B, T, N, 3) or NULL: do not materialise forces        */
    void* F_frictions;      /* (B, T, N, 3) or NULL (must be NULL iff F_springs is)   */
    void* x0z;              /* (B,)  start height written by the snap (dphysics.py:567-571) */
    void* cost;             /* (B,)  or NULL: std_t(std_p |F_spring|), step-loop variant only */
    void* contact_sum;      /* (B, T) or NULL.  Adjoint tape: the forward writes the soft-contact normaliser
                               sum_p sigmoid(-10 dh_p) of every step (dphysics.py:231), mfb_rollout_backward reads it
                               back and then runs the single-sweep adjoint (each contact point visited once per
                               step).  With NULL the backward recomputes it (three-pass adjoint, slower). */
    /* scratch */
    void* workspace;        /* device scratch of at least mfb_rollout_workspace_bytes(); holds the packed
                               per-cell sampling table the kernels build from z_grid / friction on every call */
    int64_t workspace_bytes;
} mfb_rollout_buffers;

/* Gradients for the adjoint (reverse-time) pass.  NULL input gradients are treated as zero;
 * NULL output gradients are not computed. */
typedef struct mfb_rollout_grads {
    /* incoming: d loss / d output */
    const void* g_Xs;         /* (B, T, 3)    */
    const void* g_Xds;        /* (B, T, 3)    */
    const void* g_Rs;         /* (B, T, 3, 3) */
    const void* g_Omegas;     /* (B, T, 3)    */
    const void* g_F_springs;  /* (B, T, N, 3) */
    const void* g_F_frictions;/* (B, T, N, 3) */
    const void* g_x0z;        /* (B,) gradient w.r.t. the snapped start height output */
    /* outgoing: d loss / d input.  Map gradients are ACCUMULATED (+=) into zero-initialised buffers. */
    void* g_z_grid;           /* (n_maps, H, W) */
    void* g_friction;         /* (n_maps, H, W) */
    void* g_controls;         /* (B, T, 2)      */
    void* g_x0;               /* (B, 3)  (z component is always 0: the snap overwrites it) */
    void* g_xd0;              /* (B, 3)         */
    void* g_R0;               /* (B, 3, 3)      */
    void* g_omega0;           /* (B, 3)         */
    void* g_joint_angles;     /* (B, T, 4) only with io->joint_angles (moving flippers) */
} mfb_rollout_grads;

/* ---- device-pointer entry points (asynchronous on `stream`, a cudaStream_t) ------------ */

/* Bytes of device scratch mfb_rollout_forward / mfb_rollout_backward need for `desc` (16-byte
 * aligned base required): per distinct map and per cell one 12-scalar sampling record plus one
 * 8-scalar corner-gradient record. */
int64_t mfb_rollout_workspace_bytes(const mfb_rollout_desc* desc, int dtype);

/* Replaces DPhysics.dphysics (dphysics.py:530-594): snap, T fused steps, post-processing.
 * Two kernels implement it: one warp per trajectory (large batches) and one CTA per trajectory
 * (batches below 512 trajectories, 1024 for the odeint variant; never with joint_angles).  The
 * environment variable MFB_FWD_WIDE_MAX_B overrides that threshold (0 = always one warp per
 * trajectory); the two agree to fp32 round-off (the per-step sums are associated differently). */
int mfb_rollout_forward(const mfb_rollout_desc* desc, const mfb_rollout_buffers* io,
                        int dtype, void* stream);

/* Replaces autograd's backward through DPhysics.dphysics.  Like the forward, the
 * single-sweep adjoint runs in two shapes: one CTA per trajectory up to 512 trajectories, one warp per
 * trajectory above (environment variable MFB_BWD_WIDE_MAX_B moves the switch, 0 = always one warp).  `io` must
 * hold the forward inputs and the recorded states (Xs, Xds, Rs, Omegas, x0z and, if it was
 * requested, contact_sum) of the same call. */
int mfb_rollout_backward(const mfb_rollout_desc* desc, const mfb_rollout_buffers* io,
                         const mfb_rollout_grads* grads, int dtype, void* stream);

/* ---- host-pointer entry point (synchronous; copies in, launches, copies out) ----------- */

/* Same contract as mfb_rollout_forward with HOST buffers; NULL outputs are skipped.
 * Device scratch is cached inside the library between calls. */
int mfb_rollout_forward_host(const mfb_rollout_desc* desc, const mfb_rollout_buffers* io,
                             int dtype, int device);

/* ---- terrain encoder: fused lift + splat -------------------------------------------------
 * Replaces CamEncode.get_depth_feat's soft-max (x) feature outer product (terrain_encoder/lss.py:63-71)
 * together with LiftSplatShoot.voxel_pooling incl. the sort + QuickCumsum segment sum
 * (lss.py:238-280, terrain_encoder/utils.py:144-181).  fp32, device pointers, asynchronous on `stream`.
 *   logits  (B*N, fH, fW, D + C) channels-last rows: D depth logits then C = 64 camera features
 *   vox     (B*N, D, fH, fW) int32: flat BEV cell ix*Y + iy of each frustum point, -1 if outside the grid
 *   bev     (B, X, Y, C) channels-last, ZERO-INITIALISED by the caller, accumulated with vector atomics
 * The backward writes d loss / d logits (same layout as logits) from d loss / d bev. */
int mfb_lift_splat_forward(const void* logits, const void* vox, void* bev, int B, int N, int D, int C,
                           int fH, int fW, int X, int Y, void* stream);
int mfb_lift_splat_backward(const void* logits, const void* vox, const void* g_bev, void* g_logits,
                            int B, int N, int D, int C, int fH, int fW, int X, int Y, void* stream);
/* Forward on bf16 logits rows of `row_stride` scalars (>= D + C): the tensor-core depthnet writes 128-channel bf16 rows. */
int mfb_lift_splat_forward_bf16(const void* logits, int row_stride, const void* vox, void* bev, int B, int N, int D, int C,
                                int fH, int fW, int X, int Y, void* stream);

/* ---- terrain encoder: memory-bound NHWC bf16 layers between the tensor-core convolutions (csrc/encoder_ops.cu) ----------
 * All tensors NHWC bf16 unless stated, channel counts multiples of 8, 16-byte aligned, asynchronous on `stream`.
 *
 * mfb_upsample_concat_nhwc_bf16: out[n,h,w,:] = [skip[n,h,w,:C_skip], bilinear_upsample(low)[n,h,w,:C_low], 0...] - the
 *   `torch.cat([x2, self.up(x1)], dim=1)` of Up.forward (lss.py:44-46; nn.Upsample(bilinear, align_corners=True)) and, with
 *   C_skip = 0, the x2 up-sampling in front of the BEV heads (lss.py:118,126,133).
 * mfb_stem_conv_bf16: EfficientNet-B0 stem (efficientnet_pytorch 0.7.1 `_conv_stem` + `_bn0` + swish, lss.py:78):
 *   img (N,3,H,W) fp32 NCHW -> y (N,Ho,Wo,32); w (3,3,3,32) fp32 [dy][dx][ci][co] with the BatchNorm scale folded in and
 *   shift (32,) are HOST pointers (3.5 KB, passed to the kernel by value so that they feed the FMAs from the constant bank).
 * mfb_dwconv_bn_silu_bf16: MBConv depthwise K x K (3 | 5) convolution, stride 1 | 2, low-side padding (pad_h, pad_w), folded
 *   BatchNorm, swish; `pool` (N,C) fp32, if not NULL, accumulates the sum of the outputs over the pixels (the squeeze of the
 *   squeeze-excite block; zero it first).  w (K*K, C) fp32, shift (C,) fp32.
 * mfb_se_fold_bf16: s = sigmoid(W_expand swish(W_reduce (pool * inv_hw) + b_reduce) + b_expand) per image and
 *   out_w[n,co,c] = proj_w[co,c] * s[c]: the excite scale folded into that image's projection matrix (consumed by
 *   mfb_conv2d_bf16 with per_image_weights).  w_reduce (Sq,C_se) and the TRANSPOSE of w_expand, also (Sq,C_se), fp32; channels [C_se, C) are padding.
 *   `pool` (N,C) is OVERWRITTEN with the excite scale s (two launches: the MLP once per image, then the scaling).
 * mfb_cast_f32_to_bf16: n scalars (multiple of 8). */
int mfb_upsample_concat_nhwc_bf16(const void* skip, const void* low, void* out, int N, int H, int W, int C_skip, int Hl,
                                  int Wl, int C_low, int C_out, void* stream);
int mfb_stem_conv_bf16(const void* img, const void* w, const void* shift, void* y, int N, int H, int W, int Ho, int Wo,
                       int pad_h, int pad_w, void* stream);
int mfb_dwconv_bn_silu_bf16(const void* x, const void* w, const void* shift, void* y, void* pool, int N, int H, int W, int C,
                            int Ho, int Wo, int K, int stride, int pad_h, int pad_w, void* stream);
int mfb_se_fold_bf16(void* pool, float inv_hw, const void* w_reduce, const void* b_reduce, const void* w_expand,
                     const void* b_expand, const void* proj_w, void* out_w, int N, int C, int C_se, int Sq, int Cout, void* stream);
int mfb_cast_f32_to_bf16(const void* src, void* dst, long long n, void* stream);

/* ---- between encoder, physics and planner, fp32 ---------------------------------------------------
 * mfb_terrain_postproc: terrain = geom - diff (terrain_encoder/lss.py:158) and AvgPool2d(k) (scripts/train.py:96-99,234-235) of
 *   the two maps the rollout reads, in one pass.  geom / diff / friction (B,1,X,Y) with `batch_stride` scalars between scenes;
 *   terrain (B,1,X,Y), z_pooled / mu_pooled (B, X/k, Y/k); any output may be NULL.
 * mfb_path_postproc: 4x4 poses (B,T,4,4) from Xs (B,T,3), Rs (B,T,3,3) (monoforce_ros/nodes/monoforce_node.py:80-85) and the
 *   inclination cost mean_t|roll| + mean_t|pitch| (monoforce_ros/nodes/diff_physics.py:262-266, extrinsic xyz Euler angles). */
int mfb_terrain_postproc(const void* geom, const void* diff, const void* friction, long long batch_stride, void* terrain,
                         void* z_pooled, void* mu_pooled, int B, int X, int Y, int k, void* stream);
int mfb_path_postproc(const void* Xs, const void* Rs, void* poses, void* cost, int B, int T, void* stream);

/* ---- terrain encoder: dense convolution on tcgen05 tensor cores ----------------------------
 * Replaces the Conv2d -> BatchNorm2d(eval) [-> + residual] -> activation groups of the encoder's GEMM-shaped layers:
 *   Up.conv                                  terrain_encoder/lss.py:34-41
 *   BevEncode conv1 7x7/2, ResNet-18 layer1-3 (torchvision BasicBlock: 3x3, 3x3/2, 1x1/2 downsample, add, ReLU)  lss.py:104-116,140-151
 *   BevEncode heads incl. their 1x1 output convolution and ScaledTanh / ReLU   lss.py:117-139
 *   depthnet                                 lss.py:58
 *   EfficientNet-B0 MBConv 1x1 expand / project convolutions (efficientnet_pytorch 0.7.1, called from lss.py:73-94)
 *
 *   t[n,h,w,co] = scale[co] * sum_{dy,dx,ci} x[n, stride*h + dy - pad_h, stride*w + dx - pad_w, ci] * wgt[co,dy,dx,ci] + shift[co]
 *   y           = act(t + residual)                                                  (n_heads == 0)
 *   head_out[n,g,h,w] = head_act_g(sum_{c<G} act(t)[n,h,w,g*G+c] * head_w[g*G+c] + head_bias[g]),  G = Cout / n_heads   (n_heads > 0)
 * x (N,H,W,Cin), y / residual (N,Ho,Wo,Cout) NHWC bf16; wgt (Cout,KH,KW,Cin) bf16, or (N,Cout,KH,KW,Cin) with
 * per_image_weights (an MBConv block's squeeze-excite scale folded into its projection matrix per image);
 * scale / shift / head_w fp32; head_out (N,n_heads,Ho,Wo) fp32.  Reads outside the image are zeros (the convolution's
 * padding; pad_* is the LOW-side padding, the high side follows from Ho / Wo).
 * Cin % 8 == 0, Cout % 8 == 0 (channels beyond the tensors' extent are zero-filled by TMA, nothing is padded in memory),
 * KH, KW <= 7, stride 1 or 2; head mode needs G == 128 (G == 64 when Cout == 64). */
typedef enum mfb_act { MFB_ACT_NONE = 0, MFB_ACT_RELU = 1, MFB_ACT_GELU = 2, MFB_ACT_SILU = 3 } mfb_act;
typedef enum mfb_head_act { MFB_HEAD_NONE = 0, MFB_HEAD_RELU = 1, MFB_HEAD_SCALED_TANH = 2 /* lo + (hi-lo)(tanh x + 1)/2, lss.py:17-24 */ } mfb_head_act;
typedef struct mfb_conv_desc {
    int32_t N, H, W, Cin;
    int32_t Ho, Wo, Cout;
    int32_t KH, KW, stride, pad_h, pad_w;
    int32_t act;                 /* mfb_act */
    int32_t per_image_weights;   /* 0 / 1 */
    int32_t n_heads;             /* 0: write y; 1..4: fused 1x1 heads, write head_out */
    int32_t head_act[4];         /* mfb_head_act */
    float head_bias[4], head_lo[4], head_hi[4];
} mfb_conv_desc;
int mfb_conv2d_bf16(const mfb_conv_desc* desc, const void* x, const void* wgt, const void* scale, const void* shift,
                    const void* residual, void* y, const void* head_w, void* head_out, void* stream);
/* ABI v2 form: KS x KS (1 or 3), stride 1, zero padding KS/2; act: 0 none, 1 ReLU, 2 GELU(erf). */
int mfb_conv_bn_act_bf16(const void* x, const void* wgt, const void* scale, const void* shift, void* y,
                         int N, int H, int W, int Cin, int Cout, int KS, int act, void* stream);

/* ---- training objective fused with its gradient -------------------------------------------
 * Replaces physics_loss (monoforce/src/monoforce/losses.py:102-138) and its autograd backward:
 *   j(b,k) = argmin_j |pred_ts[b,j] - gt_ts[b,k]|,  w = 1 / (1 + gamma gt_ts[b,k]),
 *   loss   = mean_{b,k,c} (X_pred[b,j,c] w - X_gt[b,k,c] w)^2          -> loss (1 scalar)
 *   g_X_pred[b,j,c] (+)= d loss / d X_pred                             (NULL: not wanted)
 * X_pred (B,T1,3), X_gt (B,T2,3); pred_ts / gt_ts rows are `*_ts_stride` scalars apart (0: one row shared by the
 * batch).  same_time_grid != 0 promises pred_ts == gt_ts element for element (then j = k, pred_ts may be NULL and
 * g_X_pred is written, not accumulated); otherwise g_X_pred must be zero-initialised.  scratch: device buffer of
 * MFB_PHYSICS_LOSS_MAX_BLOCKS doubles.  The rotation term of the reference is not used by any shipped caller. */
#define MFB_PHYSICS_LOSS_MAX_BLOCKS 4096
int mfb_physics_loss(const void* X_pred, const void* X_gt, const void* pred_ts, const void* gt_ts,
                     int64_t pred_ts_stride, int64_t gt_ts_stride, int B, int T1, int T2, double gamma,
                     int same_time_grid, void* loss, void* g_X_pred, void* scratch, int dtype, void* stream);

/* ---- bookkeeping ------------------------------------------------------------------------- */
const char* mfb_last_error(void);
int mfb_abi_version(void);
/* Number of CUDA kernels this library has launched since load (monotonic). */
long long mfb_kernel_launches(void);
/* Frees the cached device scratch of mfb_rollout_forward_host. */
void mfb_release_scratch(void);

#ifdef __cplusplus
}
#endif
#endif /* MONOFORCE_B200_H_ */
