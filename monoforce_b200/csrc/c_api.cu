// C ABI of libmonoforce_b200.so (declared in include/monoforce_b200.h).
// Validates the descriptor, converts the constants to the working precision the way torch
// converts python scalars, and dispatches to the sm_100a kernels.  No torch types here.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/monoforce_b200.h"
#include "launch.h"
#include "rollout_bwd_args.cuh"

namespace mfb {

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int fail_status(int code, const std::string& msg) { return fail(code, msg); }   // for the other translation units

static const char* check_desc(const mfb_rollout_desc* d) {
    if (!d) return "desc is NULL";
    if (d->B < 1 || d->T < 1) return "B and T must be >= 1";
    if (d->N < 1 || d->N > kMaxPointsPerLane * 32) return "N must be in [1, 256]";
    if (d->H < 2 || d->W < 2) return "map must be at least 2x2";
    if (d->H != d->W) return "the reference indexes the map with H as the row stride: H must equal W";
    if ((long long)d->H * d->W > (1ll << 30)) return "map too large";
    if (d->n_tracks != 2 && d->n_tracks != 4) return "n_tracks must be 2 or 4";
    if (d->variant != MFB_STEP_LOOP && d->variant != MFB_ODEINT_EULER) return "unknown variant";
    if (d->map_stride != 0 && d->map_stride < (long long)d->H * d->W) return "map_stride must be 0 or >= H*W";
    if (d->traj_per_map < 0) return "traj_per_map must be >= 0";
    if (!(d->grid_res > 0) || !(d->dt > 0) || !(d->mass > 0)) return "grid_res, dt and mass must be positive";
    return nullptr;
}

// trajectories per map: 0 keeps the ABI-v2 meaning (map_stride == 0: all share one map, else one map per trajectory)
static int map_group(const mfb_rollout_desc& d) {
    if (d.map_stride == 0) return d.B;
    return d.traj_per_map > 0 ? d.traj_per_map : 1;
}
static long long map_count(const mfb_rollout_desc& d) {
    const int g = map_group(d);
    return ((long long)d.B + g - 1) / g;
}

template <typename T>
static RolloutArgs<T> make_args(const mfb_rollout_desc& d, const mfb_rollout_buffers& io) {
    RolloutArgs<T> a;
    a.B = d.B; a.nT = d.T; a.N = d.N; a.H = d.H; a.W = d.W; a.n_tracks = d.n_tracks;
    a.map_stride = d.map_stride;
    a.map_group = map_group(d);
    a.mass = (T)d.mass;
    a.inv_mass = (T)1 / (T)d.mass;
    a.mg = (T)(d.mass * d.gravity);                         // clamp bounds / gravity force: python double product, then cast
    a.stiffness = (T)d.stiffness;
    a.damping = (T)d.damping;
    a.res = (T)d.grid_res;
    a.inv_res = (T)1 / (T)d.grid_res;
    a.d_max = (T)d.d_max;
    a.dt = (T)d.dt;
    a.omega_max = (T)d.omega_max;
    a.half_Ly = (T)(d.robot_Ly / 2.0);
    a.delta_h = (T)(d.mass * d.gravity / (d.stiffness + 1e-6));   // dphysics.py:587
    for (int i = 0; i < 9; ++i) a.Iinv[i] = (T)d.I_inv[i];
    a.z = (const T*)io.z_grid; a.mu = (const T*)io.friction; a.controls = (const T*)io.controls;
    a.x0 = (const T*)io.x0; a.xd0 = (const T*)io.xd0; a.R0 = (const T*)io.R0; a.om0 = (const T*)io.omega0;
    a.pts = (const T*)io.points; a.part = io.part_id; a.ts = (const T*)io.ts;
    a.joint_angles = (const T*)io.joint_angles;
    for (int i = 0; i < 12; ++i) a.joint_pos[i] = (T)d.joint_positions[i];
    a.cells = (const T*)io.workspace;
    a.cell_stride = d.map_stride == 0 ? 0 : (long long)d.H * d.W * kCellStride;
    a.Xs = (T*)io.Xs; a.Xds = (T*)io.Xds; a.Rs = (T*)io.Rs; a.Oms = (T*)io.Omegas;
    a.Fs = (T*)io.F_springs; a.Ff = (T*)io.F_frictions; a.x0z = (T*)io.x0z; a.cost = (T*)io.cost;
    a.Csum = (T*)io.contact_sum;
    return a;
}

static const char* check_io_forward(const mfb_rollout_desc& d, const mfb_rollout_buffers* io) {
    if (!io) return "io is NULL";
    if (!io->z_grid || !io->friction || !io->controls || !io->x0 || !io->xd0 || !io->R0 || !io->omega0 ||
        !io->points || !io->part_id)
        return "a required input pointer is NULL";
    if (!io->Xs || !io->Xds || !io->Rs || !io->Omegas || !io->x0z) return "a required output pointer is NULL";
    if ((io->F_springs == nullptr) != (io->F_frictions == nullptr)) return "F_springs and F_frictions must both be set or both NULL";
    if (d.variant == MFB_ODEINT_EULER && !io->ts) return "ts is required for the odeint variant";
    if (d.variant == MFB_ODEINT_EULER && io->cost) return "cost output is defined for the step-loop variant only";
    return nullptr;
}

// workspace layout: [cell table: n_maps*H*W*12 scalars][map-gradient scratch: n_maps*H*W*8 scalars]
static long long table_elems(const mfb_rollout_desc& d) {
    const long long n_maps = map_count(d);
    return n_maps * d.H * d.W * kCellStride;
}
constexpr int kGradScratchPerCell = 8;   // per-cell corner records of the single-sweep adjoint (the three-pass kernel uses 2)
static long long workspace_bytes(const mfb_rollout_desc& d, int dtype) {
    const long long n_maps = map_count(d);
    return (table_elems(d) + n_maps * d.H * d.W * kGradScratchPerCell) * (dtype == MFB_F32 ? 4 : 8);
}

static const char* check_workspace(const mfb_rollout_desc& d, const mfb_rollout_buffers& io, int dtype) {
    if (!io.workspace) return "workspace is NULL (see mfb_rollout_workspace_bytes)";
    if (io.workspace_bytes < workspace_bytes(d, dtype)) return "workspace is smaller than mfb_rollout_workspace_bytes()";
    if (((uintptr_t)io.workspace & 15) != 0) return "workspace must be 16-byte aligned";
    return nullptr;
}

// K0: packed per-cell sampling table (one thread per cell), rebuilt on every call
template <typename T>
static int build_table(const mfb_rollout_desc& d, const mfb_rollout_buffers& io, cudaStream_t st) {
    const int n_maps = (int)map_count(d);
    const long long total = (long long)n_maps * d.H * d.W;
    const int block = 256;
    const int grid = (int)std::min<long long>((total + block - 1) / block, 148 * 32);
    build_cell_table_kernel<T><<<grid, block, 0, st>>>((const T*)io.z_grid, (const T*)io.friction, (T*)io.workspace, n_maps,
                                                       d.H, d.W, d.map_stride, (T)1 / (T)d.grid_res);
    count_launch();
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return fail(MFB_ERR_CUDA, std::string("build_cell_table launch: ") + cudaGetErrorString(ce));
    return MFB_OK;
}

template <typename T>
static int forward_typed(const mfb_rollout_desc& d, const mfb_rollout_buffers& io, cudaStream_t st) {
    RolloutArgs<T> a = make_args<T>(d, io);
    if (int rc = build_table<T>(d, io, st)) return rc;
    LaunchError e = d.variant == MFB_STEP_LOOP ? launch_rollout_fwd<T, kStepLoop>(a, st)
                                               : launch_rollout_fwd<T, kOdeintEuler>(a, st);
    if (e.msg) return fail(MFB_ERR_UNSUPPORTED, e.msg);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return fail(MFB_ERR_CUDA, std::string("rollout_fwd launch: ") + cudaGetErrorString(ce));
    return MFB_OK;
}

template <typename T>
static int backward_typed(const mfb_rollout_desc& d, const mfb_rollout_buffers& io, const mfb_rollout_grads& g,
                          cudaStream_t st) {
    RolloutArgs<T> a = make_args<T>(d, io);
    if (int rc = build_table<T>(d, io, st)) return rc;
    AdjointArgs<T> ga;
    ga.g_Xs = (const T*)g.g_Xs; ga.g_Xds = (const T*)g.g_Xds; ga.g_Rs = (const T*)g.g_Rs; ga.g_Oms = (const T*)g.g_Omegas;
    ga.g_Fs = (const T*)g.g_F_springs; ga.g_Ff = (const T*)g.g_F_frictions; ga.g_x0z = (const T*)g.g_x0z;
    const long long n_maps = map_count(d);
    const bool want_maps = g.g_z_grid || g.g_friction;
    if (want_maps && d.map_stride != 0 && d.map_stride != (long long)d.H * d.W)
        return fail(MFB_ERR_UNSUPPORTED, "map gradients need densely packed per-trajectory maps (map_stride == H*W)");
    T* scratch = (T*)io.workspace + table_elems(d);
    ga.g_scratch = want_maps ? scratch : nullptr;
    ga.n_maps = n_maps;
    ga.g_maps = ga.g_cells = nullptr;
    ga.g_maps_stride = ga.g_cells_stride = 0;
    ga.g_z = (T*)g.g_z_grid; ga.g_mu = (T*)g.g_friction;
    ga.g_dir_stride = d.map_stride == 0 ? 0 : (long long)d.H * d.W;
    ga.g_controls = (T*)g.g_controls;
    ga.g_joint_angles = (T*)g.g_joint_angles;
    ga.g_x0 = (T*)g.g_x0; ga.g_xd0 = (T*)g.g_xd0; ga.g_R0 = (T*)g.g_R0; ga.g_om0 = (T*)g.g_omega0;
    LaunchError e = d.variant == MFB_STEP_LOOP ? launch_rollout_bwd<T, kStepLoop>(a, ga, st)
                                               : launch_rollout_bwd<T, kOdeintEuler>(a, ga, st);
    if (e.msg) return fail(MFB_ERR_UNSUPPORTED, e.msg);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return fail(MFB_ERR_CUDA, std::string("rollout_bwd launch: ") + cudaGetErrorString(ce));
    return MFB_OK;
}

// ---- cached device scratch for the host entry point ----------------------------------------
struct Scratch {
    std::mutex mu;
    void* ptr = nullptr;
    size_t cap = 0;
    int device = -1;
    cudaStream_t stream = nullptr;
};
static Scratch g_scratch;

static size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace mfb

using namespace mfb;

extern "C" {

int mfb_rollout_forward(const mfb_rollout_desc* desc, const mfb_rollout_buffers* io, int dtype, void* stream) {
    if (const char* m = check_desc(desc)) return fail(MFB_ERR_INVALID_ARGUMENT, m);
    if (const char* m = check_io_forward(*desc, io)) return fail(MFB_ERR_INVALID_ARGUMENT, m);
    if (dtype != MFB_F32 && dtype != MFB_F64) return fail(MFB_ERR_INVALID_ARGUMENT, "dtype must be MFB_F32 or MFB_F64");
    if (const char* m = check_workspace(*desc, *io, dtype)) return fail(MFB_ERR_INVALID_ARGUMENT, m);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MFB_F32) return forward_typed<float>(*desc, *io, st);
    if (dtype == MFB_F64) return forward_typed<double>(*desc, *io, st);
    return fail(MFB_ERR_INVALID_ARGUMENT, "dtype must be MFB_F32 or MFB_F64");
}

int mfb_rollout_backward(const mfb_rollout_desc* desc, const mfb_rollout_buffers* io, const mfb_rollout_grads* grads,
                         int dtype, void* stream) {
    if (const char* m = check_desc(desc)) return fail(MFB_ERR_INVALID_ARGUMENT, m);
    if (!io || !grads) return fail(MFB_ERR_INVALID_ARGUMENT, "io / grads is NULL");
    if (!io->z_grid || !io->friction || !io->controls || !io->x0 || !io->xd0 || !io->R0 || !io->omega0 ||
        !io->points || !io->part_id || !io->Xs || !io->Xds || !io->Rs || !io->Omegas || !io->x0z)
        return fail(MFB_ERR_INVALID_ARGUMENT, "backward needs the forward inputs and the recorded states");
    if (desc->variant == MFB_ODEINT_EULER && !io->ts) return fail(MFB_ERR_INVALID_ARGUMENT, "ts is required for the odeint variant");
    if (dtype != MFB_F32 && dtype != MFB_F64) return fail(MFB_ERR_INVALID_ARGUMENT, "dtype must be MFB_F32 or MFB_F64");
    if (const char* m = check_workspace(*desc, *io, dtype)) return fail(MFB_ERR_INVALID_ARGUMENT, m);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MFB_F32) return backward_typed<float>(*desc, *io, *grads, st);
    if (dtype == MFB_F64) return backward_typed<double>(*desc, *io, *grads, st);
    return fail(MFB_ERR_INVALID_ARGUMENT, "dtype must be MFB_F32 or MFB_F64");
}

int mfb_rollout_forward_host(const mfb_rollout_desc* desc, const mfb_rollout_buffers* io, int dtype, int device) {
    if (const char* m = check_desc(desc)) return fail(MFB_ERR_INVALID_ARGUMENT, m);
    if (!io) return fail(MFB_ERR_INVALID_ARGUMENT, "io is NULL");
    if (dtype != MFB_F32 && dtype != MFB_F64) return fail(MFB_ERR_INVALID_ARGUMENT, "dtype must be MFB_F32 or MFB_F64");
    if (!io->z_grid || !io->friction || !io->controls || !io->x0 || !io->xd0 || !io->R0 || !io->omega0 ||
        !io->points || !io->part_id)
        return fail(MFB_ERR_INVALID_ARGUMENT, "a required input pointer is NULL");
    if ((io->F_springs == nullptr) != (io->F_frictions == nullptr))
        return fail(MFB_ERR_INVALID_ARGUMENT, "F_springs and F_frictions must both be set or both NULL");
    const mfb_rollout_desc& d = *desc;
    const size_t es = dtype == MFB_F32 ? 4 : 8;
    const size_t B = d.B, T = d.T, N = d.N, HW = (size_t)d.H * d.W;
    const size_t n_maps = (size_t)map_count(d);
    const size_t map_elems = (size_t)d.map_stride * (n_maps - 1) + HW;

    // layout of the scratch arena
    struct Seg { size_t off, bytes; };
    size_t cur = 0;
    auto seg = [&](size_t bytes) { Seg s{cur, bytes}; cur += align_up(bytes); return s; };
    Seg s_z = seg(map_elems * es), s_mu = seg(map_elems * es), s_ctrl = seg(B * T * 2 * es);
    Seg s_x0 = seg(B * 3 * es), s_xd0 = seg(B * 3 * es), s_R0 = seg(B * 9 * es), s_om0 = seg(B * 3 * es);
    Seg s_pts = seg(N * 3 * es), s_part = seg(N * 4), s_ts = seg(T * es);
    Seg s_ja = seg(io->joint_angles ? B * T * 4 * es : 0);
    Seg s_Xs = seg(B * T * 3 * es), s_Xds = seg(B * T * 3 * es), s_Rs = seg(B * T * 9 * es), s_Oms = seg(B * T * 3 * es);
    Seg s_x0z = seg(B * es), s_cost = seg(B * es);
    const bool forces = io->F_springs != nullptr;
    Seg s_Fs = seg(forces ? B * T * N * 3 * es : 0), s_Ff = seg(forces ? B * T * N * 3 * es : 0);
    Seg s_ws = seg((size_t)workspace_bytes(d, dtype));

    std::lock_guard<std::mutex> lock(g_scratch.mu);
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(MFB_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
    if (g_scratch.device != device || g_scratch.cap < cur) {
        // the arena and its stream belong to ONE device: release them under that device before moving on
        if (g_scratch.device >= 0 && g_scratch.device != device) {
            cudaSetDevice(g_scratch.device);
            if (g_scratch.stream) cudaStreamDestroy(g_scratch.stream);
            g_scratch.stream = nullptr;
            if (g_scratch.ptr) cudaFree(g_scratch.ptr);
            g_scratch.ptr = nullptr; g_scratch.cap = 0;
            cudaSetDevice(device);
        }
        if (g_scratch.ptr) cudaFree(g_scratch.ptr);
        g_scratch.ptr = nullptr; g_scratch.cap = 0; g_scratch.device = -1;
        ce = cudaMalloc(&g_scratch.ptr, cur);
        if (ce != cudaSuccess) return fail(MFB_ERR_CUDA, std::string("cudaMalloc scratch: ") + cudaGetErrorString(ce));
        g_scratch.cap = cur; g_scratch.device = device;
        if (!g_scratch.stream) {
            ce = cudaStreamCreateWithFlags(&g_scratch.stream, cudaStreamNonBlocking);
            if (ce != cudaSuccess) return fail(MFB_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(ce));
        }
    }
    char* base = (char*)g_scratch.ptr;
    cudaStream_t st = g_scratch.stream;
    auto h2d = [&](Seg s, const void* src) { if (src && s.bytes) cudaMemcpyAsync(base + s.off, src, s.bytes, cudaMemcpyHostToDevice, st); };
    auto d2h = [&](void* dst, Seg s) { if (dst && s.bytes) cudaMemcpyAsync(dst, base + s.off, s.bytes, cudaMemcpyDeviceToHost, st); };
    h2d(s_z, io->z_grid); h2d(s_mu, io->friction); h2d(s_ctrl, io->controls);
    h2d(s_x0, io->x0); h2d(s_xd0, io->xd0); h2d(s_R0, io->R0); h2d(s_om0, io->omega0);
    h2d(s_pts, io->points); h2d(s_part, io->part_id); h2d(s_ts, io->ts); h2d(s_ja, io->joint_angles);

    mfb_rollout_buffers dev = *io;
    dev.z_grid = base + s_z.off; dev.friction = base + s_mu.off; dev.controls = base + s_ctrl.off;
    dev.x0 = base + s_x0.off; dev.xd0 = base + s_xd0.off; dev.R0 = base + s_R0.off; dev.omega0 = base + s_om0.off;
    dev.points = base + s_pts.off; dev.part_id = (const int32_t*)(base + s_part.off);
    dev.ts = io->ts ? base + s_ts.off : nullptr;
    dev.joint_angles = io->joint_angles ? base + s_ja.off : nullptr;   // (B,T,4) host -> device like every other input
    dev.Xs = base + s_Xs.off; dev.Xds = base + s_Xds.off; dev.Rs = base + s_Rs.off; dev.Omegas = base + s_Oms.off;
    dev.x0z = base + s_x0z.off;
    dev.cost = io->cost ? base + s_cost.off : nullptr;
    dev.F_springs = forces ? base + s_Fs.off : nullptr;
    dev.F_frictions = forces ? base + s_Ff.off : nullptr;
    dev.contact_sum = nullptr;                 // adjoint tape: not part of the host entry point
    dev.workspace = base + s_ws.off;
    dev.workspace_bytes = (int64_t)s_ws.bytes;

    int rc = mfb_rollout_forward(desc, &dev, dtype, st);
    if (rc != MFB_OK) return rc;
    d2h(io->Xs, s_Xs); d2h(io->Xds, s_Xds); d2h(io->Rs, s_Rs); d2h(io->Omegas, s_Oms);
    d2h(io->x0z, s_x0z); d2h(io->cost, s_cost);
    if (forces) { d2h(io->F_springs, s_Fs); d2h(io->F_frictions, s_Ff); }
    ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return fail(MFB_ERR_CUDA, std::string("rollout (host entry): ") + cudaGetErrorString(ce));
    return MFB_OK;
}

int64_t mfb_rollout_workspace_bytes(const mfb_rollout_desc* desc, int dtype) {
    if (check_desc(desc) || (dtype != MFB_F32 && dtype != MFB_F64)) return -1;
    return workspace_bytes(*desc, dtype);
}

const char* mfb_last_error(void) { return g_err.c_str(); }
int mfb_abi_version(void) { return MFB_ABI_VERSION; }
long long mfb_kernel_launches(void) { return g_launches.load(); }

void mfb_release_scratch(void) {
    std::lock_guard<std::mutex> lock(g_scratch.mu);
    if (g_scratch.device >= 0) cudaSetDevice(g_scratch.device);
    if (g_scratch.stream) cudaStreamDestroy(g_scratch.stream);
    g_scratch.stream = nullptr;
    if (g_scratch.ptr) cudaFree(g_scratch.ptr);
    g_scratch.ptr = nullptr; g_scratch.cap = 0; g_scratch.device = -1;
}

}  // extern "C"
