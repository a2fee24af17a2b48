// c_abi.cu -- the C ABI of include/slrbs_b200.h: context, HBM allocation, staging of
// caller-owned host buffers, stage launches, CUDA-event profiling.  No CPU fallback: every
// entry point fails with SLRBS_E_CUDA if the device is not usable.
#include "../../include/slrbs_b200.h"
#include "kernels.cuh"
#include "pgs_math.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace slrbs;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(SLRBS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

}  // namespace

struct slrbs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    View v{};
    float mu = 0.8f;
    int solver_id = 0, solver_iters = 10, collisions = 1;
    int level_group = 8;                        // lanes per scene of the level-scheduled solver
    float* bpp_scratch = nullptr;               // global scratch of the pivoting solver for scenes beyond 288 rows (allocated on first use)
    int level_variant = 0;                      // 0 k_pgs_lv, 1 k_pgs_sp (two lanes per contact row, opt-in), 2 k_pgs_pipe (contact-only scenes, pipelined)
    int broadphase = -1;                        // -1 auto, 0 all pairs, 1 hashed grid (SLRBS_BROADPHASE)
    bool joints_dirty = true;                   // joint levels must be recomputed (bodies / joints uploaded)
    std::vector<int> h_nbodies;                 // host mirror of the body counts set through slrbs_upload_bodies (0 = unknown)
    std::vector<void*> allocs;
    char* stage = nullptr; size_t stage_cap = 0, stage_off = 0;
    int* h_flag = nullptr;                      // pinned copy of the overflow flag
    // profiling
    int profiling = 0;
    std::vector<cudaEvent_t> ev; size_t ev_used = 0;
    double solve_ms = 0.0; int64_t solve_launches = 0, total_launches = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    // one step captured as a CUDA graph (launch-bound small batches); re-captured when its inputs change
    cudaGraphExec_t graph = nullptr;
    struct { float dt, mu; int iters, solver, collisions, has_ext, launches; } gsig = { 0.f, 0.f, 0, 0, 0, 0, 0 };
    int use_graph = 1;
    // extension: mesh shapes as signed-distance grids (host mirror of the device table v.sdf_shapes)
    SdfShape shapes[MAX_SDF_SHAPES] = {};
    int n_shapes = 0;
    // halo maps (device copies of the index lists)
    int *halo_src = nullptr, *halo_dst = nullptr, *halo_send = nullptr, *halo_recv = nullptr;
    int n_halo_local = 0, n_halo_send = 0, n_halo_recv = 0;
    bool halo_borrowed = false;                 // the four lists above point into the pile's buffers (not freed singly)
    // one large pile re-partitioned on the device (pile.cu)
    struct {
        bool on = false;
        PileGrid g{};
        float *radius = nullptr, *mass = nullptr, *fixed10 = nullptr;
        int *count = nullptr, *keys = nullptr, *loc = nullptr, *info = nullptr;
        int *hsrc = nullptr, *hdst = nullptr, *ldst = nullptr, *lgid = nullptr, *rdst = nullptr, *rgid = nullptr, *recv = nullptr;
        int *send = nullptr; int send_cap = 0;
        int n_left = 0, n_right = 0;
    } pile;
};

namespace {

template <typename T>
int dalloc(slrbs_ctx* c, T** p, size_t n)
{
    void* q = nullptr;
    const size_t bytes = (n ? n : 1) * sizeof(T);
    CU(cudaMalloc(&q, bytes));
    CU(cudaMemsetAsync(q, 0, bytes, c->stream));
    c->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

// bump allocator over one device staging buffer; reset at the start of every staging call
int stage_reserve(slrbs_ctx* c, size_t bytes)
{
    c->stage_off = 0;
    if (bytes <= c->stage_cap) return 0;
    CU(cudaStreamSynchronize(c->stream));
    if (c->stage) CU(cudaFree(c->stage));
    c->stage = nullptr; c->stage_cap = 0;
    const size_t cap = bytes + bytes / 4 + 4096;
    CU(cudaMalloc((void**)&c->stage, cap));
    c->stage_cap = cap;
    return 0;
}
template <typename T>
T* stage_take(slrbs_ctx* c, size_t n)
{
    const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    T* p = (T*)(c->stage + c->stage_off);
    c->stage_off += bytes;
    return p;
}
inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

template <typename T>
int stage_in(slrbs_ctx* c, const T* host, size_t n, const T** dev)
{
    if (!host) { *dev = nullptr; return 0; }
    T* d = stage_take<T>(c, n);
    CU(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    *dev = d;
    return 0;
}

int check_launch(slrbs_ctx* c, int n = 1)
{
    c->total_launches += n;
    CU(cudaGetLastError());
    return 0;
}

int range_ok(const slrbs_ctx* c, int scene0, int ns)
{
    return c && scene0 >= 0 && ns > 0 && scene0 + ns <= c->v.B;
}

int fold_events(slrbs_ctx* c)
{
    if (c->ev_used == 0) return 0;
    CU(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]));
        c->solve_ms += ms;
    }
    c->ev_used = 0;
    return 0;
}

int solve_launch(slrbs_ctx* c, float dt)
{
    if ((c->solver_id == 1 || c->solver_id == 2) && c->v.NJ > 0 && !c->v.kry) {
        int r = dalloc(c, &c->v.kry, (size_t)7 * 5 * c->v.NJ * c->v.B);
        if (r) return r;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profiling) {
        if (c->ev_used + 2 > 8192) { int r = fold_events(c); if (r) return r; }
        while (c->ev.size() < c->ev_used + 2) { cudaEvent_t e; CU(cudaEventCreate(&e)); c->ev.push_back(e); }
        e0 = c->ev[c->ev_used]; e1 = c->ev[c->ev_used + 1]; c->ev_used += 2;
        CU(cudaEventRecord(e0, c->stream));
    }
    // solvers 1-3 write their impulses, then the PGS kernel runs with zero sweeps for the force scatter:
    // 1, 2 joint-only Krylov solves; 3 block principal pivoting over joints and contacts (extension)
    int launches = 1;
    int sweeps = c->solver_iters;
    if (c->solver_id == 3) {
        const int big = bpp_big_capacity(c->v);
        if (big && !c->bpp_scratch) { int r = dalloc(c, &c->bpp_scratch, bpp_big_scratch_floats(big) * (size_t)c->v.B); if (r) return r; }
        launches = 1 + launch_bpp(c->v, c->solver_iters, c->mu, c->bpp_scratch, c->stream); sweeps = 0;
    }
    else if (c->solver_id != 0) { launch_krylov(c->v, c->solver_id, c->solver_iters, c->stream); sweeps = 0; launches = 2; }
    if (c->v.color_mode) {
        const cudaError_t e = (cudaError_t)launch_pgs_color(c->v, dt, sweeps, c->mu, c->stream);
        if (e != cudaSuccess) return fail(SLRBS_E_CUDA, std::string("k_pgs_color (cooperative launch): ") + cudaGetErrorString(e));
    }
    else if (c->v.level_mode) launch_pgs_level(c->v, dt, sweeps, c->mu, c->level_group, c->level_variant, c->stream);
    else launch_pgs(c->v, dt, sweeps, c->mu, c->stream);
    if (c->profiling) { CU(cudaEventRecord(e1, c->stream)); c->solve_launches++; }
    return check_launch(c, launches);
}

}  // namespace

extern "C" {

const char* slrbs_last_error(void) { return g_err.c_str(); }
const char* slrbs_version(void) { return "slrbs_b200 0.1 (sm_100a)"; }

int slrbs_describe(slrbs_ctx* c, char* buf, int buf_len)
{
    if (!c || !buf || buf_len <= 0) return fail(SLRBS_E_ARG, "slrbs_describe: bad argument");
    char tmp[256];
    if (c->v.level_mode && c->level_variant == 5)
        std::snprintf(tmp, sizeof tmp, "solver=k_pgs_stage (level-scheduled, 1 warp/scene, next chunk's 208 B rows staged in shared memory by cp.async) detect=%s graph=%d",
                      detect_variant(c->v, c->broadphase), c->use_graph);
    else if (c->v.level_mode && c->level_variant == 4)
        std::snprintf(tmp, sizeof tmp, "solver=k_pgs_gacc (level-scheduled, 1 warp/scene, accumulators in L1/L2, 18 warps/SM, %s B rows) detect=%s graph=%d",
                      c->v.crow_compact ? "208" : "256", detect_variant(c->v, c->broadphase), c->use_graph);
    else if (c->v.color_mode)
        std::snprintf(tmp, sizeof tmp, "solver=k_pgs_color (graph-coloured, all SMs, accumulators in HBM) detect=%s graph=%d",
                      detect_variant(c->v, c->broadphase), c->use_graph);
    else if (c->v.level_mode && c->level_variant == 2)
        std::snprintf(tmp, sizeof tmp, "solver=k_pgs_pipe (level-scheduled, 1 warp/scene, chunk table, next chunk loaded under the solve, sliding L2 window, %s B rows) detect=%s graph=%d",
                      c->v.crow_compact ? "208" : "256", detect_variant(c->v, c->broadphase), c->use_graph);
    else if (c->v.level_mode && c->level_variant == 1)
        std::snprintf(tmp, sizeof tmp, "solver=k_pgs_sp<ROWS=%d> (level-scheduled, 2 lanes/contact row, %d threads/scene) detect=%s graph=%d",
                      c->level_group, 2 * c->level_group, detect_variant(c->v, c->broadphase), c->use_graph);
    else if (c->v.level_mode)
        std::snprintf(tmp, sizeof tmp, "solver=k_pgs_lv<G=%d> (level-scheduled, %d lanes/scene) detect=%s graph=%d", c->level_group,
                      c->level_group, detect_variant(c->v, c->broadphase), c->use_graph);
    else
        std::snprintf(tmp, sizeof tmp, "solver=k_pgs (thread/scene, cp.async row ring) detect=%s graph=%d",
                      detect_variant(c->v, c->broadphase), c->use_graph);
    std::snprintf(buf, (size_t)buf_len, "%s", tmp);
    return (int)std::strlen(buf);
}

int slrbs_create(slrbs_ctx** out, int device, int n_scenes, int max_bodies, int max_joints, int max_contacts)
{
    if (!out || n_scenes <= 0 || max_bodies <= 0 || max_bodies > MAX_BODIES_PER_SCENE || max_joints < 0 || max_contacts < 0)
        return fail(SLRBS_E_ARG, "slrbs_create: bad sizes (max_bodies must be 1..32767)");
    *out = nullptr;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(SLRBS_E_CUDA, "slrbs_create: no such CUDA device (no CPU fallback)");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(SLRBS_E_CUDA, std::string("slrbs_create: device is ") + prop.name + ", kernels are built for sm_100a only");
    CU(cudaSetDevice(device));
    init_detect_kernels();
    init_pair_kernels();
    init_level_kernels();
    init_bpp_kernel();
    init_pile_kernels();
    slrbs_ctx* c = new slrbs_ctx();
    { const char* g = std::getenv("SLRBS_GRAPH"); if (g) c->use_graph = std::atoi(g); }
    c->device = device;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    View& v = c->v;
    v.B = n_scenes; v.NB = max_bodies; v.NJ = max_joints; v.NC = max_contacts;
    // Solver variant. Small and medium batches use the level-scheduled kernel (pgs_level.cu: G lanes per
    // scene; measured 1.4x (cars) to 30x (1000-sphere piles) faster up to ~8k scenes); very large batches
    // use one thread per scene, which then has a GPU's worth of independent scenes to hide latency. Both
    // give the same impulses bit for bit; SLRBS_PGS_MODE=thread|level and SLRBS_LEVEL_GROUP=4|8|16|32 override.
    {
        const char* m = std::getenv("SLRBS_PGS_MODE");
        const bool fits = level_mode_fits(max_bodies) && detect_warp_fits(max_bodies) &&
                          contact_levels_smem(max_bodies, max_contacts) <= 200 * 1024 && max_contacts < 65535;   // (level, rank) are packed 16 : 16
        // scenes too large for one SM's shared memory, a few at a time: graph-coloured solver (pgs_color.cu; order differs from
        // the reference, validated on the constraint residual). SLRBS_PGS_MODE=color forces it for smaller scenes.
        const bool colorable = max_contacts < 65535 && max_contacts > 0;
        if (m && std::strcmp(m, "thread") == 0) v.level_mode = 0;
        else if (m && std::strcmp(m, "level") == 0) v.level_mode = fits ? 1 : 0;
        else if (m && std::strcmp(m, "color") == 0) { v.level_mode = colorable ? 1 : 0; v.color_mode = v.level_mode; }
        else if (!level_mode_fits(max_bodies) && n_scenes <= 64 && colorable) { v.level_mode = 1; v.color_mode = 1; }
        else v.level_mode = (n_scenes <= 8192 && fits) ? 1 : 0;
        const char* g = std::getenv("SLRBS_LEVEL_GROUP");
        c->level_group = level_group_size(v, g ? std::atoi(g) : 0);
        c->level_variant = !v.level_mode ? 0 : (v.color_mode ? 3 : (level_split_wanted(v, c->level_group) ? 1 : (level_gacc_wanted(v, c->level_group) ? 4 : (level_stage_wanted(v, c->level_group) ? 5 : (level_pipe_wanted(v, c->level_group) ? 2 : 0)))));
        v.crow_compact = (c->level_variant == 5 || ((c->level_variant == 2 || c->level_variant == 4) && pipe_rows_compact())) ? 1 : 0;
        if (v.color_mode) c->use_graph = 0;                     // the solver is one cooperative launch; such scenes are not launch bound
        const char* bp = std::getenv("SLRBS_BROADPHASE");
        if (bp) c->broadphase = std::atoi(bp) ? 1 : 0;
    }
    const size_t nb = (size_t)v.NB * v.B, nj = (size_t)v.NJ * v.B, nc = (size_t)v.NC * v.B;
    const size_t bpad = ((size_t)v.B + 31) / 32 * 32;      // row records are blocked 32 scenes at a time
    int r = 0;
#define A(p, n) if ((r = dalloc(c, &p, n))) { slrbs_destroy(c); return r; }
    A(v.nbodies, v.B) A(v.njoints, v.B) A(v.ncontacts, v.B) A(v.overflow, 1) A(v.visit_counters, 2) A(v.ext_mode, v.B)
    A(v.pos, nb) A(v.quat, nb) A(v.vel, nb) A(v.omg, nb) A(v.flags, nb) A(v.geomA, nb) A(v.geomB, nb)
    A(v.ibody, 9 * nb) A(v.ibodyinv, 9 * nb) A(v.cproxy, nb) A(v.bext, nb) A(v.force, nb) A(v.torque, nb)
    A(v.extf, nb) A(v.extt, nb) A(v.Iw, 9 * nb) A(v.Iinvw, 9 * nb) A(v.wl, nb) A(v.wa, nb) A(v.brec, 8 * nb)
    A(v.s_pos, nb) A(v.s_quat, nb) A(v.s_vel, nb) A(v.s_omg, nb)
    A(v.jdef, nj) A(v.jr0, nj) A(v.jq0, nj) A(v.jr1, nj) A(v.jq1, nj) A(v.jlam, nj) A(v.jlam4, nj)
    A(v.jrow, (size_t)JROW_FIELDS * v.NJ * bpad)
    A(v.cids, nc) A(v.cp, nc) A(v.cn, nc) A(v.crow, (size_t)CROW_FIELDS * v.NC * bpad)
    { SdfShape* tab = nullptr; A(tab, MAX_SDF_SHAPES) v.sdf_shapes = tab; }
    if (c->level_variant == 4) { A(v.wacc, 2 * nb) }
    if (c->level_variant == 2) {                               // chunks of a sweep <= levels + rows / 32 <= NC + NC / 32
        v.cchunk_cap = v.NC + v.NC / 32 + 2;
        A(v.cchunk, (size_t)v.cchunk_cap * v.B)
    }
    if (v.color_mode) {
        A(v.cmask, nb) A(v.cwin, nb) A(v.nclev_max, 1)
        // scenes no single CTA can hold are detected pair-parallel over the whole GPU (detect_pairs.cu), extension pairs or not
        if (!detect_warp_fits(max_bodies) || max_bodies > 3000) {
            const int cap = 64 * v.NB;
            A(v.pc_pairs, (size_t)cap * v.B) A(v.pc_count, v.B) A(v.pc_hits, v.B) A(v.pc_out, (size_t)v.NC * v.B * 6)
            v.pc_cap = cap;
        }
    }
    if (v.level_mode) {
        A(v.cpos, nc) A(v.cslot, nc) A(v.clev_end, nc + v.B + 65 * (size_t)v.B) A(v.nclev, v.B) A(v.lv_last, nb)
        A(v.jpos, nj) A(v.jlev_end, nj + v.B) A(v.njlev, v.B)
    }
#undef A
    CU(cudaMallocHost((void**)&c->h_flag, sizeof(int)));
    *c->h_flag = 0;
    CU(cudaStreamSynchronize(c->stream));
    *out = c;
    return 0;
}

void slrbs_destroy(slrbs_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void* p : c->allocs) cudaFree(p);
    if (c->stage) cudaFree(c->stage);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->graph) cudaGraphExecDestroy(c->graph);
    if (!c->halo_borrowed) { cudaFree(c->halo_src); cudaFree(c->halo_dst); cudaFree(c->halo_send); cudaFree(c->halo_recv); }
    if (c->pile.send) cudaFree(c->pile.send);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int slrbs_set_params(slrbs_ctx* c, float mu, int solver_id, int solver_iters, int collisions_enabled)
{
    if (!c || solver_id < 0 || solver_id > 3 || solver_iters < 0) return fail(SLRBS_E_ARG, "slrbs_set_params: bad argument");
    c->mu = mu; c->solver_id = solver_id; c->solver_iters = solver_iters; c->collisions = collisions_enabled ? 1 : 0;
    return 0;
}

// EXTENSION (no reference counterpart): meshes as signed-distance grids
int slrbs_sdf_build(int device, int nv, const float* verts, int nt, const int32_t* tris, int nx, int ny, int nz,
                    const float* origin3, float cell, float* grid_out)
{
    if (nv <= 0 || nt <= 0 || !verts || !tris || nx < 2 || ny < 2 || nz < 2 || !origin3 || !(cell > 0.f) || !grid_out)
        return fail(SLRBS_E_ARG, "slrbs_sdf_build: bad argument");
    for (int k = 0; k < 3 * nt; ++k) if (tris[k] < 0 || tris[k] >= nv) return fail(SLRBS_E_ARG, "slrbs_sdf_build: triangle index out of range");
    CU(cudaSetDevice(device));
    const size_t n = (size_t)nx * ny * nz;
    float *dv = nullptr, *dg = nullptr; int* dt = nullptr;
    CU(cudaMalloc((void**)&dv, sizeof(float) * 3 * nv));
    CU(cudaMalloc((void**)&dt, sizeof(int) * 3 * nt));
    CU(cudaMalloc((void**)&dg, sizeof(float) * n));
    int rc = 0;
    cudaError_t e = cudaMemcpy(dv, verts, sizeof(float) * 3 * nv, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dt, tris, sizeof(int) * 3 * nt, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { launch_build_sdf(dv, nt, dt, nx, ny, nz, origin3[0], origin3[1], origin3[2], cell, dg, 0); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(grid_out, dg, sizeof(float) * n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(SLRBS_E_CUDA, std::string("slrbs_sdf_build: ") + cudaGetErrorString(e));
    cudaFree(dv); cudaFree(dt); cudaFree(dg);
    return rc;
}

int slrbs_sdf_add_shape(slrbs_ctx* c, int nx, int ny, int nz, const float* origin3, float cell, const float* grid, int nv, const float* verts)
{
    if (!c || nx < 2 || ny < 2 || nz < 2 || !origin3 || !(cell > 0.f) || !grid || nv <= 0 || !verts)
        return fail(SLRBS_E_ARG, "slrbs_sdf_add_shape: bad argument");
    if (c->n_shapes >= MAX_SDF_SHAPES) return fail(SLRBS_E_CAPACITY, "slrbs_sdf_add_shape: shape table is full");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)nx * ny * nz;
    float *dg = nullptr, *dv = nullptr;
    int r;
    if ((r = dalloc(c, &dg, n)) || (r = dalloc(c, &dv, (size_t)3 * nv))) return r;
    CU(cudaMemcpyAsync(dg, grid, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dv, verts, sizeof(float) * 3 * nv, cudaMemcpyHostToDevice, c->stream));
    SdfShape& S = c->shapes[c->n_shapes];
    S.grid = dg; S.verts = dv; S.nx = nx; S.ny = ny; S.nz = nz; S.nv = nv;
    S.ox = origin3[0]; S.oy = origin3[1]; S.oz = origin3[2]; S.cell = cell;
    S.radius = 0.f;
    for (int k = 0; k < nv; ++k) {                     // largest vertex norm, the restatement's association a0 + (a1 + a2)
        const float x = verts[3 * k], y = verts[3 * k + 1], z = verts[3 * k + 2];
        const float rr = std::sqrt(x * x + (y * y + z * z));
        if (rr > S.radius) S.radius = rr;
    }
    CU(cudaMemcpyAsync(const_cast<SdfShape*>(c->v.sdf_shapes) + c->n_shapes, &S, sizeof(SdfShape), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));             // the host buffers may go away after the call
    c->v.n_sdf = c->n_shapes + 1;
    if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; }     // the captured launches carry the old View
    return c->n_shapes++;
}

int slrbs_set_extensions(slrbs_ctx* c, int flags)
{
    if (!c || flags < 0 || flags > 1) return fail(SLRBS_E_ARG, "slrbs_set_extensions: bad argument");
    if (c->v.ext_pairs != (flags & 1) && c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; }   // the View is baked into the graph
    c->v.ext_pairs = flags & 1;
    // scenes of >= 128 bodies with the extension pairs use the pair-parallel detection (detect_pairs.cu): candidate
    // candidate lists of 64 pairs per body and scene, one result record per contact slot; SLRBS_DETECT_PAIRS=0 keeps the row-parallel kernels
    const char* e = std::getenv("SLRBS_DETECT_PAIRS");
    if (c->v.ext_pairs && c->v.NB >= 128 && !c->v.pc_pairs && !(e && std::atoi(e) == 0)) {
        CU(cudaSetDevice(c->device));
        const int cap = 64 * c->v.NB;
        int r;
        if ((r = dalloc(c, &c->v.pc_pairs, (size_t)cap * c->v.B)) || (r = dalloc(c, &c->v.pc_count, (size_t)c->v.B)) ||
            (r = dalloc(c, &c->v.pc_hits, (size_t)c->v.B)) || (r = dalloc(c, &c->v.pc_out, (size_t)c->v.NC * c->v.B * 6)))
            return r;
        c->v.pc_cap = cap;
    }
    return 0;
}

int slrbs_upload_bodies(slrbs_ctx* c, int scene0, int ns, int n, const int32_t* counts,
                        const float* x3, const float* q4, const float* v3, const float* w3,
                        const float* mass, const float* ibody9, const float* ibodyinv9,
                        const int32_t* fixed, const int32_t* geom_type, const float* geom6)
{
    if (!range_ok(c, scene0, ns) || n <= 0 || n > c->v.NB) return fail(SLRBS_E_ARG, "slrbs_upload_bodies: bad range");
    if (!x3 || !q4 || !v3 || !w3 || !mass || !ibody9 || !ibodyinv9 || !fixed || !geom_type || !geom6)
        return fail(SLRBS_E_ARG, "slrbs_upload_bodies: all body arrays are required");
    const size_t m = (size_t)ns * n;
    if (counts) for (int k = 0; k < ns; ++k)
        if (counts[k] < 0 || counts[k] > n) return fail(SLRBS_E_ARG, "slrbs_upload_bodies: counts[] must lie in 0..n_bodies");
    for (size_t k = 0; k < m; ++k)
        if (geom_type[k] < GEOM_SPHERE || geom_type[k] > GEOM_SDF) return fail(SLRBS_E_ARG, "slrbs_upload_bodies: geom_type must be 0..4");
    CU(cudaSetDevice(c->device));
    int r = stage_reserve(c, pad256(m * 4) * (3 + 4 + 3 + 3 + 1 + 9 + 9 + 1 + 1 + 6) + pad256((size_t)ns * 4) + 4096);
    if (r) return r;
    BodyStage st{};
    const int* dcounts = nullptr;
    if ((r = stage_in(c, x3, 3 * m, &st.x)) || (r = stage_in(c, q4, 4 * m, &st.q)) || (r = stage_in(c, v3, 3 * m, &st.v)) ||
        (r = stage_in(c, w3, 3 * m, &st.w)) || (r = stage_in(c, mass, m, &st.mass)) || (r = stage_in(c, ibody9, 9 * m, &st.ibody)) ||
        (r = stage_in(c, ibodyinv9, 9 * m, &st.ibodyinv)) || (r = stage_in(c, fixed, m, &st.fixed)) ||
        (r = stage_in(c, geom_type, m, &st.geom_type)) || (r = stage_in(c, geom6, 6 * m, &st.geom)) ||
        (r = stage_in(c, counts, (size_t)ns, &dcounts)))
        return r;
    launch_upload_bodies(c->v, scene0, ns, n, dcounts, st, c->stream);
    if ((r = check_launch(c))) return r;
    if (c->h_nbodies.empty()) c->h_nbodies.assign((size_t)c->v.B, 0);
    for (int k = 0; k < ns; ++k) c->h_nbodies[scene0 + k] = counts ? counts[k] : n;
    c->joints_dirty = true;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int slrbs_upload_joints(slrbs_ctx* c, int scene0, int ns, int n, const int32_t* counts,
                        const int32_t* type, const int32_t* body0, const int32_t* body1,
                        const float* r0, const float* q0, const float* r1, const float* q1)
{
    if (!range_ok(c, scene0, ns) || n < 0 || n > c->v.NJ) return fail(SLRBS_E_ARG, "slrbs_upload_joints: bad range");
    CU(cudaSetDevice(c->device));
    if (n == 0) {
        // removing the joints of a scene range: the level tables (njlev, jlev_end, jpos) are what k_pgs_lv walks, so they
        // must be rebuilt, and the removed joints' impulses must not survive into a later upload
        CU(cudaMemsetAsync(c->v.njoints + scene0, 0, sizeof(int) * ns, c->stream));
        if (c->v.NJ > 0) { launch_clear_joint_lambda(c->v, scene0, ns, c->stream); int r0 = check_launch(c); if (r0) return r0; }
        c->joints_dirty = true;
        CU(cudaStreamSynchronize(c->stream));
        return 0;
    }
    if (!type || !body0 || !body1 || !r0 || !r1) return fail(SLRBS_E_ARG, "slrbs_upload_joints: type/body/r arrays are required");
    const size_t m = (size_t)ns * n;
    if (counts) for (int k = 0; k < ns; ++k)
        if (counts[k] < 0 || counts[k] > n) return fail(SLRBS_E_ARG, "slrbs_upload_joints: counts[] must lie in 0..n_joints");
    for (size_t i = 0; i < m; ++i) {
        const int sc = scene0 + (int)(i / (size_t)n);
        if (counts && (int)(i % (size_t)n) >= counts[sc - scene0]) continue;                             // padding entries are never read
        if (type[i] != 1 && type[i] != 2) return fail(SLRBS_E_ARG, "slrbs_upload_joints: type must be 1 (spherical) or 2 (hinge)");
        const int lim = (!c->h_nbodies.empty() && c->h_nbodies[sc] > 0) ? c->h_nbodies[sc] : c->v.NB;   // the scene's own body count when known
        if (body0[i] < 0 || body0[i] >= lim || body1[i] < 0 || body1[i] >= lim)
            return fail(SLRBS_E_ARG, "slrbs_upload_joints: body index out of range for its scene");
    }
    int r = stage_reserve(c, pad256(m * 4) * (3 + 3 + 4 + 3 + 4) + pad256((size_t)ns * 4) + 4096);
    if (r) return r;
    JointStage st{};
    const int* dcounts = nullptr;
    if ((r = stage_in(c, type, m, &st.type)) || (r = stage_in(c, body0, m, &st.body0)) || (r = stage_in(c, body1, m, &st.body1)) ||
        (r = stage_in(c, r0, 3 * m, &st.r0)) || (r = stage_in(c, q0, 4 * m, &st.q0)) || (r = stage_in(c, r1, 3 * m, &st.r1)) ||
        (r = stage_in(c, q1, 4 * m, &st.q1)) || (r = stage_in(c, counts, (size_t)ns, &dcounts)))
        return r;
    launch_upload_joints(c->v, scene0, ns, n, dcounts, st, c->stream);
    if ((r = check_launch(c))) return r;
    c->joints_dirty = true;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

static int upload_state_impl(slrbs_ctx* c, int scene0, int ns, int n, const float* x3, const float* q4, const float* v3, const float* w3, bool wait);

int slrbs_upload_state(slrbs_ctx* c, int scene0, int ns, int n, const float* x3, const float* q4, const float* v3, const float* w3)
{
    return upload_state_impl(c, scene0, ns, n, x3, q4, v3, w3, true);
}

int slrbs_upload_state_async(slrbs_ctx* c, int scene0, int ns, int n, const float* x3, const float* q4, const float* v3, const float* w3)
{
    return upload_state_impl(c, scene0, ns, n, x3, q4, v3, w3, false);
}

static int upload_state_impl(slrbs_ctx* c, int scene0, int ns, int n, const float* x3, const float* q4, const float* v3, const float* w3, bool wait)
{
    if (!range_ok(c, scene0, ns) || n <= 0 || n > c->v.NB) return fail(SLRBS_E_ARG, "slrbs_upload_state: bad range");
    CU(cudaSetDevice(c->device));
    const size_t m = (size_t)ns * n;
    int r = stage_reserve(c, pad256(m * 4) * 13 + 4096);
    if (r) return r;
    BodyStage st{};
    if ((r = stage_in(c, x3, 3 * m, &st.x)) || (r = stage_in(c, q4, 4 * m, &st.q)) || (r = stage_in(c, v3, 3 * m, &st.v)) ||
        (r = stage_in(c, w3, 3 * m, &st.w)))
        return r;
    // counts unchanged: pass the live counts back to themselves
    launch_upload_bodies(c->v, scene0, ns, n, c->v.nbodies + scene0, st, c->stream);
    if ((r = check_launch(c))) return r;
    if (wait) CU(cudaStreamSynchronize(c->stream));      // the caller may reuse its (possibly pinned) buffers on return
    return 0;
}

static int download_state_impl(slrbs_ctx* c, int scene0, int ns, int n, float* x3, float* q4, float* v3, float* w3, bool wait);

int slrbs_download_state(slrbs_ctx* c, int scene0, int ns, int n, float* x3, float* q4, float* v3, float* w3)
{
    return download_state_impl(c, scene0, ns, n, x3, q4, v3, w3, true);
}

int slrbs_download_state_async(slrbs_ctx* c, int scene0, int ns, int n, float* x3, float* q4, float* v3, float* w3)
{
    return download_state_impl(c, scene0, ns, n, x3, q4, v3, w3, false);
}

static int download_state_impl(slrbs_ctx* c, int scene0, int ns, int n, float* x3, float* q4, float* v3, float* w3, bool wait)
{
    if (!range_ok(c, scene0, ns) || n <= 0 || n > c->v.NB) return fail(SLRBS_E_ARG, "slrbs_download_state: bad range");
    CU(cudaSetDevice(c->device));
    const size_t m = (size_t)ns * n;
    int r = stage_reserve(c, pad256(m * 4) * 13 + 4096);
    if (r) return r;
    BodyStage st{};
    float *dx = x3 ? stage_take<float>(c, 3 * m) : nullptr, *dq = q4 ? stage_take<float>(c, 4 * m) : nullptr;
    float *dv = v3 ? stage_take<float>(c, 3 * m) : nullptr, *dw = w3 ? stage_take<float>(c, 3 * m) : nullptr;
    st.x = dx; st.q = dq; st.v = dv; st.w = dw;
    launch_download_state(c->v, scene0, ns, n, st, c->stream);
    if ((r = check_launch(c))) return r;
    if (x3) CU(cudaMemcpyAsync(x3, dx, 3 * m * 4, cudaMemcpyDeviceToHost, c->stream));
    if (q4) CU(cudaMemcpyAsync(q4, dq, 4 * m * 4, cudaMemcpyDeviceToHost, c->stream));
    if (v3) CU(cudaMemcpyAsync(v3, dv, 3 * m * 4, cudaMemcpyDeviceToHost, c->stream));
    if (w3) CU(cudaMemcpyAsync(w3, dw, 3 * m * 4, cudaMemcpyDeviceToHost, c->stream));
    if (wait) CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int slrbs_set_external_forces(slrbs_ctx* c, int scene0, int ns, int n, const float* f3, const float* tau3, int mode)
{
    if (mode != 0 && mode != 1) return fail(SLRBS_E_ARG, "slrbs_set_external_forces: mode must be 0 or 1");
    if (!range_ok(c, scene0, ns) || n <= 0 || n > c->v.NB) return fail(SLRBS_E_ARG, "slrbs_set_external_forces: bad range");
    CU(cudaSetDevice(c->device));
    const size_t m = (size_t)ns * n;
    int r = stage_reserve(c, pad256(m * 4) * 6 + 4096);
    if (r) return r;
    const float *df = nullptr, *dt = nullptr;
    if ((r = stage_in(c, f3, 3 * m, &df)) || (r = stage_in(c, tau3, 3 * m, &dt))) return r;
    // the mode is kept per scene (View::ext_mode): a call on a sub-range leaves the other scenes as they were
    launch_set_ext(c->v, scene0, ns, n, df, dt, (f3 || tau3) ? 1 + mode : 0, c->stream);
    if ((r = check_launch(c))) return r;
    if (f3 || tau3) c->v.has_ext = 1;
    else if (scene0 == 0 && ns == c->v.B) c->v.has_ext = 0;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int slrbs_stage_prepare(slrbs_ctx* c)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    launch_prepare(c->v, c->stream);
    return check_launch(c);
}
int slrbs_stage_detect(slrbs_ctx* c)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!c->collisions || c->v.NC == 0) return 0;               // RigidBodySystem.cpp:77
    launch_detect(c->v, c->broadphase, c->stream);
    if (c->v.color_mode) { launch_contact_colors(c->v, c->stream); return check_launch(c, 2 + (c->v.pc_pairs ? 2 : 0)); }
    if (c->v.level_mode) { launch_contact_levels(c->v, c->stream); return check_launch(c, 2); }
    return check_launch(c);
}
int slrbs_stage_jacobians(slrbs_ctx* c, float dt)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    int n = 0;
    if (c->v.level_mode && c->joints_dirty) { launch_joint_levels(c->v, c->stream); c->joints_dirty = false; ++n; }
    if (c->collisions && c->v.NC > 0) { launch_contact_rows(c->v, dt, c->stream); ++n; }
    if (c->v.NJ > 0) { launch_joint_rows(c->v, dt, c->stream); ++n; }
    return check_launch(c, n);
}
int slrbs_stage_solve(slrbs_ctx* c, float dt)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    return solve_launch(c, dt);
}
int slrbs_stage_integrate(slrbs_ctx* c, float dt)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    launch_integrate(c->v, dt, c->stream);
    return check_launch(c);
}

static int step_once(slrbs_ctx* c, float dt)
{
    int r;
    if ((r = slrbs_stage_prepare(c)) || (r = slrbs_stage_detect(c)) || (r = slrbs_stage_jacobians(c, dt)) ||
        (r = slrbs_stage_solve(c, dt)) || (r = slrbs_stage_integrate(c, dt)))
        return r;
    return 0;
}

int slrbs_step(slrbs_ctx* c, float dt, int n_steps)
{
    if (!c || n_steps < 0) return fail(SLRBS_E_ARG, "slrbs_step: bad argument");
    CU(cudaSetDevice(c->device));
    int r;
    if (!c->use_graph || c->profiling || n_steps == 0) {
        for (int i = 0; i < n_steps; ++i) if ((r = step_once(c, dt))) return r;
        return 0;
    }
    // Replay one captured step n times: a small batch is launch-bound (5-7 kernels of a few microseconds each).
    // Anything that is not a plain launch on the stream happens before the capture.
    if (c->v.level_mode && c->joints_dirty) { launch_joint_levels(c->v, c->stream); c->joints_dirty = false; if ((r = check_launch(c))) return r; }
    if ((c->solver_id == 1 || c->solver_id == 2) && c->v.NJ > 0 && !c->v.kry) { if ((r = dalloc(c, &c->v.kry, (size_t)7 * 5 * c->v.NJ * c->v.B))) return r; }
    if (c->solver_id == 3 && !c->bpp_scratch) { const int big = bpp_big_capacity(c->v); if (big && (r = dalloc(c, &c->bpp_scratch, bpp_big_scratch_floats(big) * (size_t)c->v.B))) return r; }
    const bool same = c->graph && c->gsig.dt == dt && c->gsig.mu == c->mu && c->gsig.iters == c->solver_iters &&
                      c->gsig.solver == c->solver_id && c->gsig.collisions == c->collisions && c->gsig.has_ext == c->v.has_ext;
    if (!same) {
        if (c->graph) { CU(cudaGraphExecDestroy(c->graph)); c->graph = nullptr; }
        const int64_t before = c->total_launches;
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        r = step_once(c, dt);
        const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (r) { if (g) cudaGraphDestroy(g); return r; }
        if (e != cudaSuccess) return fail(SLRBS_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
        CU(cudaGraphInstantiate(&c->graph, g, 0));
        CU(cudaGraphDestroy(g));
        c->gsig.launches = (int)(c->total_launches - before);
        c->total_launches = before;
        c->gsig.dt = dt; c->gsig.mu = c->mu; c->gsig.iters = c->solver_iters; c->gsig.solver = c->solver_id;
        c->gsig.collisions = c->collisions; c->gsig.has_ext = c->v.has_ext;
    }
    for (int i = 0; i < n_steps; ++i) CU(cudaGraphLaunch(c->graph, c->stream));
    c->total_launches += (int64_t)n_steps * c->gsig.launches;
    return 0;
}

int slrbs_sync(slrbs_ctx* c)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->h_flag, c->v.overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const int flags = *c->h_flag;
    if (!flags) return 0;
    // every cause is reported once per occurrence: contacts are re-detected each step, so a transient overflow must not
    // poison the context, and the caller can tell the causes apart (only the first is cured by a larger max_contacts)
    std::string pairs_msg;
    if (flags & 4) {
        std::vector<int> cand(c->v.B), hits(c->v.B);
        CU(cudaMemcpy(cand.data(), c->v.pc_count, sizeof(int) * c->v.B, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(hits.data(), c->v.pc_hits, sizeof(int) * c->v.B, cudaMemcpyDeviceToHost));
        int mc = 0, mh = 0;
        for (int k = 0; k < c->v.B; ++k) { mc = std::max(mc, cand[k]); mh = std::max(mh, hits[k]); }
        pairs_msg = "a scene produced more candidate or touching pairs than the pair-parallel detection holds (last step: " +
                    std::to_string(mc) + " candidates of " + std::to_string(c->v.pc_cap) + ", " + std::to_string(mh) + " touching pairs of " +
                    std::to_string(std::min(c->v.NC, 16384)) + "); contacts were dropped";
    }
    CU(cudaMemsetAsync(c->v.overflow, 0, sizeof(int), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (flags & 1) return fail(SLRBS_E_CAPACITY, "a scene produced more contacts than max_contacts; extra contacts were dropped");
    if (flags & 4) return fail(SLRBS_E_CAPACITY_PAIRS, pairs_msg);
    if (flags & 8) return fail(SLRBS_E_CAPACITY_ROWS, "graph colouring: a body has contacts of more than 64 colours; those contacts share the last colour (solved with a race)");
    if (flags & 2) return fail(SLRBS_E_CAPACITY_ROWS, "a scene has more constraint rows than the pivoting solver holds (solver_id 3); its impulses were left at zero");
    return 0;
}

int slrbs_download_contact_counts(slrbs_ctx* c, int scene0, int ns, int32_t* counts)
{
    if (!range_ok(c, scene0, ns) || !counts) return fail(SLRBS_E_ARG, "slrbs_download_contact_counts: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(counts, c->v.ncontacts + scene0, sizeof(int) * ns, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

static int scene_count(slrbs_ctx* c, const int* dev_counts, int scene, int* out)
{
    CU(cudaMemcpyAsync(out, dev_counts + scene, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int slrbs_download_contacts(slrbs_ctx* c, int scene, int max_out, int32_t* body0, int32_t* body1,
                            float* p3, float* n3, float* t3, float* b3, float* phi, float* lambda3)
{
    if (!range_ok(c, scene, 1) || max_out < 0) return fail(SLRBS_E_ARG, "slrbs_download_contacts: bad argument");
    CU(cudaSetDevice(c->device));
    int total = 0, r;
    if ((r = scene_count(c, c->v.ncontacts, scene, &total))) return r;
    const int n = total < max_out ? total : max_out;
    if (n == 0) return total;
    if ((r = stage_reserve(c, pad256((size_t)n * 4) * (2 + 3 * 4 + 1 + 3) + 4096))) return r;
    int *d0 = stage_take<int>(c, n), *d1 = stage_take<int>(c, n);
    float *dp = stage_take<float>(c, 3 * n), *dn = stage_take<float>(c, 3 * n), *dt = stage_take<float>(c, 3 * n);
    float *db = stage_take<float>(c, 3 * n), *dphi = stage_take<float>(c, n), *dl = stage_take<float>(c, 3 * n);
    launch_export_contacts(c->v, scene, n, d0, d1, dp, dn, dt, db, dphi, dl, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, c->stream);
    if ((r = check_launch(c))) return r;
#define D2H(h, d, cnt) if (h) CU(cudaMemcpyAsync(h, d, (size_t)(cnt) * 4, cudaMemcpyDeviceToHost, c->stream));
    D2H(body0, d0, n) D2H(body1, d1, n) D2H(p3, dp, 3 * n) D2H(n3, dn, 3 * n) D2H(t3, dt, 3 * n) D2H(b3, db, 3 * n)
    D2H(phi, dphi, n) D2H(lambda3, dl, 3 * n)
    CU(cudaStreamSynchronize(c->stream));
    return total;
}

int slrbs_download_contact_rows(slrbs_ctx* c, int scene, int max_out, float* J0, float* J1, float* M0, float* M1, float* A9, float* rhs3)
{
    if (!range_ok(c, scene, 1) || max_out < 0) return fail(SLRBS_E_ARG, "slrbs_download_contact_rows: bad argument");
    CU(cudaSetDevice(c->device));
    int total = 0, r;
    if ((r = scene_count(c, c->v.ncontacts, scene, &total))) return r;
    const int n = total < max_out ? total : max_out;
    if (n == 0) return total;
    if ((r = stage_reserve(c, pad256((size_t)n * 4) * (18 * 4 + 9 + 3) + 4096))) return r;
    float *d0 = stage_take<float>(c, 18 * n), *d1 = stage_take<float>(c, 18 * n), *m0 = stage_take<float>(c, 18 * n);
    float *m1 = stage_take<float>(c, 18 * n), *dA = stage_take<float>(c, 9 * n), *db = stage_take<float>(c, 3 * n);
    launch_export_contacts(c->v, scene, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d0, d1, m0, m1, dA, db, c->stream);
    if ((r = check_launch(c))) return r;
    D2H(J0, d0, 18 * n) D2H(J1, d1, 18 * n) D2H(M0, m0, 18 * n) D2H(M1, m1, 18 * n) D2H(A9, dA, 9 * n) D2H(rhs3, db, 3 * n)
    CU(cudaStreamSynchronize(c->stream));
    return total;
}

int slrbs_download_joint_rows(slrbs_ctx* c, int scene, int max_out, float* J0, float* J1, float* M0, float* M1, float* rhs5)
{
    if (!range_ok(c, scene, 1) || max_out < 0) return fail(SLRBS_E_ARG, "slrbs_download_joint_rows: bad argument");
    CU(cudaSetDevice(c->device));
    int total = 0, r;
    if ((r = scene_count(c, c->v.njoints, scene, &total))) return r;
    const int n = total < max_out ? total : max_out;
    if (n == 0) return total;
    if ((r = stage_reserve(c, pad256((size_t)n * 4) * (30 * 4 + 5) + 4096))) return r;
    float *d0 = stage_take<float>(c, 30 * n), *d1 = stage_take<float>(c, 30 * n), *m0 = stage_take<float>(c, 30 * n);
    float *m1 = stage_take<float>(c, 30 * n), *db = stage_take<float>(c, 5 * n);
    launch_export_joints(c->v, scene, n, d0, d1, m0, m1, db, c->stream);
    if ((r = check_launch(c))) return r;
    D2H(J0, d0, 30 * n) D2H(J1, d1, 30 * n) D2H(M0, m0, 30 * n) D2H(M1, m1, 30 * n) D2H(rhs5, db, 5 * n)
    CU(cudaStreamSynchronize(c->stream));
    return total;
}

int slrbs_download_joint_lambda(slrbs_ctx* c, int scene0, int ns, int n, float* lambda5)
{
    if (!range_ok(c, scene0, ns) || n <= 0 || n > c->v.NJ || !lambda5) return fail(SLRBS_E_ARG, "slrbs_download_joint_lambda: bad argument");
    CU(cudaSetDevice(c->device));
    const size_t m = (size_t)ns * n;
    int r = stage_reserve(c, pad256(m * 20) + 4096);
    if (r) return r;
    float* d = stage_take<float>(c, 5 * m);
    launch_joint_lambda(c->v, scene0, ns, n, d, 0, c->stream);
    if ((r = check_launch(c))) return r;
    CU(cudaMemcpyAsync(lambda5, d, 5 * m * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int slrbs_upload_joint_lambda(slrbs_ctx* c, int scene0, int ns, int n, const float* lambda5)
{
    if (!range_ok(c, scene0, ns) || n <= 0 || n > c->v.NJ || !lambda5) return fail(SLRBS_E_ARG, "slrbs_upload_joint_lambda: bad argument");
    CU(cudaSetDevice(c->device));
    const size_t m = (size_t)ns * n;
    int r = stage_reserve(c, pad256(m * 20) + 4096);
    if (r) return r;
    const float* d = nullptr;
    if ((r = stage_in(c, lambda5, 5 * m, &d))) return r;
    launch_joint_lambda(c->v, scene0, ns, n, (float*)d, 1, c->stream);
    if ((r = check_launch(c))) return r;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int slrbs_download_contact_schedule(slrbs_ctx* c, int scene, int max_out, int32_t* position, int max_levels, int32_t* level_end, int32_t* n_levels)
{
    if (!range_ok(c, scene, 1) || max_out < 0 || max_levels < 0) return fail(SLRBS_E_ARG, "slrbs_download_contact_schedule: bad argument");
    if (!c->v.level_mode) return fail(SLRBS_E_UNSUPPORTED, "slrbs_download_contact_schedule: the thread-per-scene solver keeps the contacts in slot order");
    CU(cudaSetDevice(c->device));
    int total = 0, r, nl = 0;
    if ((r = scene_count(c, c->v.ncontacts, scene, &total))) return r;
    if (total > 0 && (r = scene_count(c, c->v.nclev, scene, &nl))) return r;
    if (n_levels) *n_levels = nl;
    const int n = total < max_out ? total : max_out;
    // cpos is [B][NC] in level / colour mode (atc), clev_end [level][B]: a strided column of one scene
    if (n > 0 && position)
        CU(cudaMemcpyAsync(position, c->v.cpos + (size_t)scene * c->v.NC, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    const int m = nl < max_levels ? nl : max_levels;
    if (m > 0 && level_end)
        CU(cudaMemcpy2DAsync(level_end, sizeof(int), c->v.clev_end + (size_t)c->v.B + scene, sizeof(int) * (size_t)c->v.B, sizeof(int), (size_t)m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return total;
}

int slrbs_upload_contact_lambda(slrbs_ctx* c, int scene, int n, const float* lambda3)
{
    if (!range_ok(c, scene, 1) || n < 0 || (n > 0 && !lambda3)) return fail(SLRBS_E_ARG, "slrbs_upload_contact_lambda: bad argument");
    CU(cudaSetDevice(c->device));
    int total = 0, r;
    if ((r = scene_count(c, c->v.ncontacts, scene, &total))) return r;
    if (n > total) return fail(SLRBS_E_ARG, "slrbs_upload_contact_lambda: more impulses than the scene has contacts");
    if (n == 0) return 0;
    if ((r = stage_reserve(c, pad256((size_t)n * 12) + 4096))) return r;
    const float* d = nullptr;
    if ((r = stage_in(c, lambda3, (size_t)3 * n, &d))) return r;
    launch_import_contact_lambda(c->v, scene, n, d, c->stream);
    if ((r = check_launch(c))) return r;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int slrbs_download_body_dyn(slrbs_ctx* c, int scene, int n, float* f3, float* tau3, float* fc3, float* tauc3, float* I9, float* Iinv9)
{
    if (!range_ok(c, scene, 1) || n <= 0 || n > c->v.NB) return fail(SLRBS_E_ARG, "slrbs_download_body_dyn: bad argument");
    CU(cudaSetDevice(c->device));
    int r = stage_reserve(c, pad256((size_t)n * 4) * (12 + 18) + 4096);
    if (r) return r;
    float *df = stage_take<float>(c, 3 * n), *dt = stage_take<float>(c, 3 * n), *dfc = stage_take<float>(c, 3 * n);
    float *dtc = stage_take<float>(c, 3 * n), *dI = stage_take<float>(c, 9 * n), *dIi = stage_take<float>(c, 9 * n);
    launch_export_body_dyn(c->v, scene, n, df, dt, dfc, dtc, dI, dIi, c->stream);
    if ((r = check_launch(c))) return r;
    D2H(f3, df, 3 * n) D2H(tau3, dt, 3 * n) D2H(fc3, dfc, 3 * n) D2H(tauc3, dtc, 3 * n) D2H(I9, dI, 9 * n) D2H(Iinv9, dIi, 9 * n)
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
#undef D2H

int slrbs_save_state(slrbs_ctx* c)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    launch_snapshot(c->v, 0, c->v.B, 0, c->stream);
    return check_launch(c);
}

int slrbs_restore_state(slrbs_ctx* c, int scene0, int ns)
{
    if (!range_ok(c, scene0, ns)) return fail(SLRBS_E_ARG, "slrbs_restore_state: bad range");
    CU(cudaSetDevice(c->device));
    launch_snapshot(c->v, scene0, ns, 1, c->stream);
    return check_launch(c);
}

static int set_index_list(slrbs_ctx* c, int** dev, int* count, int n, const int32_t* host, const char* what)
{
    const int64_t limit = (int64_t)c->v.B * c->v.NB;
    for (int i = 0; i < n; ++i)
        if (host[i] < 0 || host[i] >= limit) return fail(SLRBS_E_ARG, std::string(what) + ": flat body index out of range");
    CU(cudaStreamSynchronize(c->stream));
    if (c->halo_borrowed) { c->halo_src = c->halo_dst = c->halo_send = c->halo_recv = nullptr; c->halo_borrowed = false; }
    if (*dev) { CU(cudaFree(*dev)); *dev = nullptr; }
    *count = 0;
    if (n > 0) {
        CU(cudaMalloc((void**)dev, sizeof(int) * (size_t)n));
        CU(cudaMemcpy(*dev, host, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
        *count = n;
    }
    return 0;
}

int slrbs_halo_set_local(slrbs_ctx* c, int n, const int32_t* src_flat, const int32_t* dst_flat)
{
    if (!c || n < 0 || (n > 0 && (!src_flat || !dst_flat))) return fail(SLRBS_E_ARG, "slrbs_halo_set_local: bad argument");
    CU(cudaSetDevice(c->device));
    int r, cnt = 0;
    if ((r = set_index_list(c, &c->halo_src, &cnt, n, src_flat, "slrbs_halo_set_local"))) return r;
    if ((r = set_index_list(c, &c->halo_dst, &c->n_halo_local, n, dst_flat, "slrbs_halo_set_local"))) return r;
    return 0;
}

int slrbs_halo_set_remote(slrbs_ctx* c, int n_send, const int32_t* send_flat, int n_recv, const int32_t* recv_flat)
{
    if (!c || n_send < 0 || n_recv < 0 || (n_send > 0 && !send_flat) || (n_recv > 0 && !recv_flat))
        return fail(SLRBS_E_ARG, "slrbs_halo_set_remote: bad argument");
    CU(cudaSetDevice(c->device));
    int r;
    if ((r = set_index_list(c, &c->halo_send, &c->n_halo_send, n_send, send_flat, "slrbs_halo_set_remote"))) return r;
    if ((r = set_index_list(c, &c->halo_recv, &c->n_halo_recv, n_recv, recv_flat, "slrbs_halo_set_remote"))) return r;
    return 0;
}

int slrbs_halo_sync_local(slrbs_ctx* c)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    launch_halo_copy(c->v, c->n_halo_local, c->halo_src, c->halo_dst, c->stream);
    return check_launch(c);
}

namespace slrbs {
__global__ void k_debug_div(const float* a, const float* b, float* out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = div_rn_nocall(a[i], b[i]);
}
}  // namespace slrbs

int slrbs_debug_div_rn(const float* a, const float* b, float* out, int n)
{
    if (!a || !b || !out || n < 0) return fail(SLRBS_E_ARG, "slrbs_debug_div_rn: null pointer or negative count");
    if (n == 0) return 0;
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    const size_t bytes = (size_t)n * sizeof(float);
    cudaError_t e = cudaMalloc(&da, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&db, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dout, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        slrbs::k_debug_div<<<(unsigned)((n + 255) / 256), 256>>>(da, db, dout, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dout);
    if (e != cudaSuccess) return fail(SLRBS_E_CUDA, std::string("slrbs_debug_div_rn: ") + cudaGetErrorString(e));
    return 0;
}

void* slrbs_stream(slrbs_ctx* c) { return c ? (void*)c->stream : nullptr; }
int slrbs_device(slrbs_ctx* c) { return c ? c->device : -1; }

int slrbs_halo_pack(slrbs_ctx* c, void* device_buffer)
{
    if (!c || (c->n_halo_send > 0 && !device_buffer)) return fail(SLRBS_E_ARG, "slrbs_halo_pack: bad argument");
    CU(cudaSetDevice(c->device));
    launch_halo_pack(c->v, c->n_halo_send, c->halo_send, (float*)device_buffer, c->stream);
    return check_launch(c);
}

int slrbs_halo_unpack(slrbs_ctx* c, const void* device_buffer)
{
    if (!c || (c->n_halo_recv > 0 && !device_buffer)) return fail(SLRBS_E_ARG, "slrbs_halo_unpack: bad argument");
    CU(cudaSetDevice(c->device));
    launch_halo_unpack(c->v, c->n_halo_recv, c->halo_recv, (const float*)device_buffer, c->stream);
    return check_launch(c);
}

// ---- one large pile, re-partitioned on the device (pile.cu) ----------------------------------------------------------
int slrbs_pile_init(slrbs_ctx* c, int n_global, const float* radius, const float* mass, const float* origin3, float block,
                    float halo_w, const int32_t* ncell3, int cx0, int cx1, int n_fixed, const float* fixed10)
{
    if (!c || n_global <= 0 || !radius || !mass || !origin3 || !ncell3 || block <= 0.f || halo_w < 0.f || halo_w > block ||
        n_fixed < 0 || (n_fixed > 0 && !fixed10) || cx0 < 0 || cx1 <= cx0 || cx1 > ncell3[0] || n_global > (1 << 29))
        return fail(SLRBS_E_ARG, "slrbs_pile_init: bad argument");
    if ((int64_t)(cx1 - cx0) * ncell3[1] * ncell3[2] != c->v.B)
        return fail(SLRBS_E_ARG, "slrbs_pile_init: the context must hold one scene per owned block ((cx1-cx0)*ncy*ncz)");
    if (c->v.NB <= n_fixed) return fail(SLRBS_E_ARG, "slrbs_pile_init: max_bodies must exceed the number of fixed bodies");
    if (c->pile.on) return fail(SLRBS_E_ARG, "slrbs_pile_init: already initialised for this context");
    CU(cudaSetDevice(c->device));
    auto& P = c->pile;
    P.g.ox = origin3[0]; P.g.oy = origin3[1]; P.g.oz = origin3[2]; P.g.block = block; P.g.halo_w = halo_w;
    P.g.ncx = ncell3[0]; P.g.ncy = ncell3[1]; P.g.ncz = ncell3[2]; P.g.cx0 = cx0; P.g.cx1 = cx1;
    P.g.nfixed = n_fixed; P.g.nbdyn = c->v.NB - n_fixed; P.g.n_global = n_global;
    if (pile_sort_smem(P.g.nbdyn) > 128 * 1024) return fail(SLRBS_E_ARG, "slrbs_pile_init: more than 32768 dynamic bodies per block");
    const size_t slots = (size_t)c->v.B * P.g.nbdyn;
    int r;
    if ((r = dalloc(c, &P.radius, (size_t)n_global)) || (r = dalloc(c, &P.mass, (size_t)n_global)) ||
        (r = dalloc(c, &P.fixed10, (size_t)10 * (n_fixed ? n_fixed : 1))) || (r = dalloc(c, &P.count, (size_t)c->v.B)) ||
        (r = dalloc(c, &P.keys, slots)) || (r = dalloc(c, &P.loc, (size_t)n_global)) || (r = dalloc(c, &P.info, 8)) ||
        (r = dalloc(c, &P.hsrc, slots)) || (r = dalloc(c, &P.hdst, slots)) || (r = dalloc(c, &P.ldst, slots)) ||
        (r = dalloc(c, &P.lgid, slots)) || (r = dalloc(c, &P.rdst, slots)) || (r = dalloc(c, &P.rgid, slots)) ||
        (r = dalloc(c, &P.recv, slots)))
        return r;
    CU(cudaMemcpyAsync(P.radius, radius, sizeof(float) * (size_t)n_global, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(P.mass, mass, sizeof(float) * (size_t)n_global, cudaMemcpyHostToDevice, c->stream));
    if (n_fixed) CU(cudaMemcpyAsync(P.fixed10, fixed10, sizeof(float) * 10 * (size_t)n_fixed, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    P.on = true;
    return 0;
}

int slrbs_pile_export(slrbs_ctx* c, void* d_table, void* d_vmax2)
{
    if (!c || !c->pile.on || !d_table || !d_vmax2) return fail(SLRBS_E_ARG, "slrbs_pile_export: bad argument");
    CU(cudaSetDevice(c->device));
    launch_pile_export(c->v, c->pile.g, c->pile.count, c->pile.keys, (float*)d_table, (unsigned*)d_vmax2, c->stream);
    return check_launch(c);
}

int slrbs_pile_rebuild(slrbs_ctx* c, const void* d_table, int32_t* info8)
{
    if (!c || !c->pile.on || !d_table || !info8) return fail(SLRBS_E_ARG, "slrbs_pile_rebuild: bad argument");
    CU(cudaSetDevice(c->device));
    auto& P = c->pile;
    launch_pile_rebuild(c->v, P.g, (const float*)d_table, P.radius, P.mass, P.fixed10, P.count, P.keys, P.loc, P.info,
                        P.hsrc, P.hdst, P.ldst, P.lgid, P.rdst, P.rgid, c->stream);
    int r;
    if ((r = check_launch(c, 4))) return r;
    CU(cudaMemcpyAsync(info8, P.info, sizeof(int) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (!c->halo_borrowed) {
        cudaFree(c->halo_src); cudaFree(c->halo_dst); cudaFree(c->halo_send); cudaFree(c->halo_recv);
        c->halo_send = nullptr; c->n_halo_send = 0;
    }
    c->halo_borrowed = true;
    c->halo_src = P.hsrc; c->halo_dst = P.hdst; c->n_halo_local = info8[0] ? 0 : info8[2];
    P.n_left = info8[3]; P.n_right = info8[4];
    if (P.n_left) CU(cudaMemcpyAsync(P.recv, P.ldst, sizeof(int) * (size_t)P.n_left, cudaMemcpyDeviceToDevice, c->stream));
    if (P.n_right) CU(cudaMemcpyAsync(P.recv + P.n_left, P.rdst, sizeof(int) * (size_t)P.n_right, cudaMemcpyDeviceToDevice, c->stream));
    c->halo_recv = P.recv; c->n_halo_recv = info8[0] ? 0 : P.n_left + P.n_right;
    c->halo_send = P.send; c->n_halo_send = 0;                  // until slrbs_pile_set_send
    c->joints_dirty = true;
    return info8[0] ? fail(SLRBS_E_CAPACITY, "slrbs_pile_rebuild: a block holds more bodies than max_bodies") : 0;
}

int slrbs_pile_requests(slrbs_ctx* c, void* d_gids_out)
{
    if (!c || !c->pile.on || (!d_gids_out && c->pile.n_left + c->pile.n_right > 0)) return fail(SLRBS_E_ARG, "slrbs_pile_requests: bad argument");
    CU(cudaSetDevice(c->device));
    auto& P = c->pile;
    int* o = (int*)d_gids_out;
    if (P.n_left) CU(cudaMemcpyAsync(o, P.lgid, sizeof(int) * (size_t)P.n_left, cudaMemcpyDeviceToDevice, c->stream));
    if (P.n_right) CU(cudaMemcpyAsync(o + P.n_left, P.rgid, sizeof(int) * (size_t)P.n_right, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int slrbs_pile_set_send(slrbs_ctx* c, const void* d_gids, int n)
{
    if (!c || !c->pile.on || n < 0 || (n > 0 && !d_gids)) return fail(SLRBS_E_ARG, "slrbs_pile_set_send: bad argument");
    CU(cudaSetDevice(c->device));
    auto& P = c->pile;
    if (n > P.send_cap) {
        CU(cudaStreamSynchronize(c->stream));
        if (P.send) CU(cudaFree(P.send));
        P.send = nullptr; P.send_cap = 0;
        const int cap = n + n / 4 + 256;
        CU(cudaMalloc((void**)&P.send, sizeof(int) * (size_t)cap));
        P.send_cap = cap;
    }
    launch_pile_lookup((const int*)d_gids, n, P.g.n_global, P.loc, P.send, P.info, c->stream);
    int r;
    if ((r = check_launch(c))) return r;
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, P.info + 7, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (flag) return fail(SLRBS_E_ARG, "slrbs_pile_set_send: a requested body is not owned by this rank");
    c->halo_send = P.send; c->n_halo_send = n;
    return 0;
}

int slrbs_pile_scene_keys(slrbs_ctx* c, int scene, int max_out, int32_t* keys_out)
{
    if (!c || !c->pile.on || scene < 0 || scene >= c->v.B || max_out < 0 || (max_out > 0 && !keys_out))
        return fail(SLRBS_E_ARG, "slrbs_pile_scene_keys: bad argument");
    CU(cudaSetDevice(c->device));
    int n = 0;
    CU(cudaMemcpyAsync(&n, c->pile.count + scene, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    n = n < c->pile.g.nbdyn ? n : c->pile.g.nbdyn;
    const int m = n < max_out ? n : max_out;
    if (m > 0) CU(cudaMemcpy(keys_out, c->pile.keys + (size_t)scene * c->pile.g.nbdyn, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost));
    return n;
}

int slrbs_profile_enable(slrbs_ctx* c, int on)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!on) { int r = fold_events(c); if (r) return r; }
    c->profiling = on ? 1 : 0;
    c->v.profiling = c->profiling;
    return 0;
}

int slrbs_timer_start(slrbs_ctx* c)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!c->t0) { CU(cudaEventCreate(&c->t0)); CU(cudaEventCreate(&c->t1)); }
    CU(cudaEventRecord(c->t0, c->stream));
    return 0;
}

int slrbs_timer_stop(slrbs_ctx* c, float* ms)
{
    if (!c || !ms || !c->t0) return fail(SLRBS_E_ARG, "slrbs_timer_stop: timer not started");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->t1, c->stream));
    CU(cudaEventSynchronize(c->t1));
    CU(cudaEventElapsedTime(ms, c->t0, c->t1));
    return 0;
}

int slrbs_profile_read(slrbs_ctx* c, int reset, int64_t* solve_launches, double* solve_ms,
                       int64_t* contact_visits, int64_t* joint_visits, int64_t* total_launches)
{
    if (!c) return fail(SLRBS_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    int r = fold_events(c);
    if (r) return r;
    unsigned long long h[2] = { 0, 0 };
    CU(cudaMemcpyAsync(h, c->v.visit_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (solve_launches) *solve_launches = c->solve_launches;
    if (solve_ms) *solve_ms = c->solve_ms;
    if (contact_visits) *contact_visits = (int64_t)h[0];
    if (joint_visits) *joint_visits = (int64_t)h[1];
    if (total_launches) *total_launches = c->total_launches;
    if (reset) {
        c->solve_launches = 0; c->solve_ms = 0.0; c->total_launches = 0;
        CU(cudaMemsetAsync(c->v.visit_counters, 0, sizeof h, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

}  // extern "C"
