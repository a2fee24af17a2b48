// kernels.cuh -- launcher declarations shared by kernels.cu and c_abi.cu
#pragma once
#include "layout.cuh"

namespace slrbs {

// device-side staging views of the scene-major host arrays (all pointers are device pointers)
struct BodyStage {
    const float *x, *q, *v, *w, *mass, *ibody, *ibodyinv, *geom;
    const int *fixed, *geom_type;
};
struct JointStage {
    const int *type, *body0, *body1;
    const float *r0, *q0, *r1, *q1;
};

// fixed grid of cubic blocks of one large pile (pile.cu): this rank owns the blocks with block-x in [cx0, cx1)
struct PileGrid {
    float ox, oy, oz, block, halo_w;
    int ncx, ncy, ncz, cx0, cx1;
    int nbdyn;        // dynamic-body capacity of a scene (max_bodies - nfixed)
    int nfixed;       // container boxes appended to every scene
    int n_global;     // bodies of the whole pile
};

void launch_prepare(const View& v, cudaStream_t s);
// detect.cu; broadphase: 0 all pairs, 1 hashed grid, -1 auto (grid for scenes of >= 128 bodies)
void launch_detect(const View& v, int broadphase, cudaStream_t s);
bool detect_warp_fits(int NB);
void init_detect_kernels();
// detect_pairs.cu: pair-parallel detection for extension scenes of >= 128 bodies (used when v.pc_pairs is set)
void init_pair_kernels();
void launch_detect_pairs(const View& v, cudaStream_t s);
const char* detect_variant(const View& v, int broadphase);
void init_level_kernels();
void launch_contact_rows(const View& v, float h, cudaStream_t s);
void launch_joint_rows(const View& v, float h, cudaStream_t s);
void launch_pgs(const View& v, float dt, int iters, float mu, cudaStream_t s);
void launch_krylov(const View& v, int solver_id, int iters, cudaStream_t s);
// bpp.cu: block principal pivoting (solverId 3, extension); one CTA per scene, at most BPP_MAX_ROWS scalar rows per scene
// (shared memory); scenes beyond that run from a global scratch area, up to BPP_BIG_MAX_ROWS
enum { BPP_MAX_ROWS = 288, BPP_BIG_MAX_ROWS = 4096 };
void init_bpp_kernel();
int  bpp_row_capacity(const View& v);
int  bpp_big_capacity(const View& v);
size_t bpp_big_scratch_floats(int cap);
int  launch_bpp(const View& v, int iters, float mu, float* big_scratch, cudaStream_t s);   // returns the number of launches
// sdf.cu (extension): signed-distance grid of a closed triangle mesh, all pointers on the device
void launch_build_sdf(const float* verts, int nt, const int* tris, int nx, int ny, int nz, float ox, float oy, float oz, float cell,
                      float* out, cudaStream_t s);
// level-scheduled solver (pgs_level.cu)
void launch_joint_levels(const View& v, cudaStream_t s);
void launch_contact_levels(const View& v, cudaStream_t s);
size_t contact_levels_smem(int NB, int NC);
void launch_pgs_level(const View& v, float dt, int iters, float mu, int group, int variant, cudaStream_t s);   // variant: 0 k_pgs_lv, 1 k_pgs_sp, 2 k_pgs_pipe, 4 k_pgs_gacc, 5 k_pgs_stage
int  level_group_size(const View& v, int requested);
// pgs_split.cu: the level schedule with two lanes per contact row (one per body side); rows = 32 or 64 rows per level step
void init_split_kernels();
void launch_pgs_split(const View& v, float dt, int iters, float mu, int rows, cudaStream_t s);
bool level_split_wanted(const View& v, int group);
// pgs_pipe.cu: one warp per contact-only scene, next chunk's rows prefetched into a second register buffer
void init_pipe_kernels();
void launch_pgs_pipe(const View& v, float dt, int iters, float mu, cudaStream_t s);
bool level_pipe_wanted(const View& v, int group);
bool pipe_rows_compact();
// pgs_stage.cu: one warp per contact-only scene, next chunk's compact row records copied to shared memory by cp.async
void init_stage_kernels();
void launch_pgs_stage(const View& v, float dt, int iters, float mu, cudaStream_t s);
bool level_stage_wanted(const View& v, int group);
// pgs_gacc.cu: one warp per contact-only scene, accumulators in global memory (L1 / L2), 18 warps per SM
void init_gacc_kernels();
void launch_pgs_gacc(const View& v, float dt, int iters, float mu, cudaStream_t s);
bool level_gacc_wanted(const View& v, int group);
bool level_mode_fits(int max_bodies);
// pgs_color.cu: graph-coloured solver for large scenes (accumulators in HBM, one grid barrier per colour)
void launch_contact_colors(const View& v, cudaStream_t s);
int  launch_pgs_color(const View& v, float dt, int iters, float mu, cudaStream_t s);   // cudaError_t of the cooperative launch
void launch_integrate(const View& v, float dt, cudaStream_t s);

void launch_upload_bodies(const View& v, int scene0, int ns, int n, const int* counts, const BodyStage& st, cudaStream_t s);
void launch_download_state(const View& v, int scene0, int ns, int n, const BodyStage& st, cudaStream_t s);
void launch_upload_joints(const View& v, int scene0, int ns, int n, const int* counts, const JointStage& st, cudaStream_t s);
void launch_clear_joint_lambda(const View& v, int scene0, int ns, cudaStream_t s);
void launch_joint_lambda(const View& v, int scene0, int ns, int n, float* lam5, int upload, cudaStream_t s);
void launch_set_ext(const View& v, int scene0, int ns, int n, const float* f3, const float* t3, int mode, cudaStream_t s);
void launch_snapshot(const View& v, int scene0, int ns, int restore, cudaStream_t s);
void launch_halo_copy(const View& v, int n, const int* src, const int* dst, cudaStream_t s);
void launch_halo_pack(const View& v, int n, const int* idx, float* buf, cudaStream_t s);
void launch_halo_unpack(const View& v, int n, const int* idx, const float* buf, cudaStream_t s);
void init_pile_kernels();
size_t pile_sort_smem(int nbdyn);
void launch_pile_export(const View& v, const PileGrid& g, const int* count, const int* keys, float* table, unsigned* vmax2, cudaStream_t s);
void launch_pile_rebuild(const View& v, const PileGrid& g, const float* table, const float* radius, const float* mass, const float* fixed10,
                         int* count, int* keys, int* loc, int* info, int* hsrc, int* hdst, int* ldst, int* lgid, int* rdst, int* rgid,
                         cudaStream_t s);
void launch_pile_lookup(const int* gids, int n, int n_global, const int* loc, int* out, int* info, cudaStream_t s);
void launch_export_contacts(const View& v, int scene, int n, int* body0, int* body1, float* p3, float* n3, float* t3,
                            float* b3, float* phi, float* lam3, float* J0, float* J1, float* M0, float* M1,
                            float* A9, float* rhs3, cudaStream_t s);
void launch_import_contact_lambda(const View& v, int scene, int n, const float* lam3, cudaStream_t s);
void launch_export_joints(const View& v, int scene, int n, float* J0, float* J1, float* M0, float* M1, float* rhs5, cudaStream_t s);
void launch_export_body_dyn(const View& v, int scene, int n, float* f3, float* tau3, float* fc3, float* tauc3,
                            float* I9, float* Iinv9, cudaStream_t s);

}  // namespace slrbs
