// layout.cuh -- device data layout of a batch of scenes in HBM.
//
// Every per-item array is "slot-major, scene-minor": element (slot i, scene s) lives at
// i * B + s, so the 32 lanes of a warp that work on 32 consecutive scenes read one
// contiguous 128 B / 512 B segment per field (coalesced float / float4 streams), whatever
// the item index is.  With B = 1 this degenerates to plain SoA.
//
// Constraint ROW records (the stream the PGS sweep re-reads every iteration) are blocked by
// warp instead: scenes are grouped 32 at a time, and the rows of a group are laid out
// [group][slot][field][lane], i.e. one row of one group is a contiguous 16-field x 32-lane x 16 B
// = 8 KB tile and consecutive rows follow each other. A warp sweeping its 32 scenes therefore
// streams one private, contiguous HBM region front to back (DRAM-page and TLB friendly; one
// tile is what a stage of the shared-memory ring holds).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slrbs {

enum { GEOM_SPHERE = 0, GEOM_BOX = 1, GEOM_PLANE = 2, GEOM_CYLINDER = 3, GEOM_SDF = 4 /* extension: mesh as signed-distance grid */ };
enum { FLAG_TYPE_MASK = 0xF, FLAG_FIXED = 0x10 };
enum { CROW_FIELDS = 16, JROW_FIELDS = 17, CROW_LAMBDA = 15 };
enum { MAX_BODIES_PER_SCENE = 32767 };   // body ids are packed 2 x 16 bit in the contact row record
enum { BODY_FIXED_BIT = (int)0x80000000 };

// Contact row record (CROW_FIELDS float4 per contact, written by k_contact_rows, streamed by k_pgs):
//  0 n   | im0      3 a0n | rhs0     6 a1n | A00      9 m0n | A01     12 m1n | A12
//  1 t   | im1      4 a0t | rhs1     7 a1t | A11     10 m0t | A02     13 m1t | A20
//  2 b   | ids      5 a0b | rhs2     8 a1b | A22     11 m0b | A10     14 m1b | A21
//  15 lambda | -   (the only field the sweep writes)
// d in {n,t,b}: J0 row = [d, a0d], J1 row = [-d, a1d], J0Minv row = [im0*d, m0d],
// J1Minv row = [-im1*d, m1d]; ids = body0 | body1 << 16, each 15-bit index with bit 15 set
// when the body is fixed.
//
// Joint row record (JROW_FIELDS float4 per joint; see k_joint_rows):
//  0 rr0 | im0     2 nxu | body0     4..8  J0Minv angular rows 0..4 | rhs[r]
//  1 rr1 | im1     3 nxv | body1     9..13 J1Minv angular rows 0..4 | 1/D[r]
//  14 (L10 L20 L21 L30)  15 (L31 L32 L40 L41)  16 (L42 L43 dim slot)
// J0 = [I hat(-rr0); 0 nxu; 0 nxv], J1 = [-I hat(rr1); 0 -nxu; 0 -nxv] (Hinge.cpp:50-57).

// extension (sdf.cuh): a mesh as signed-distance grid + sample vertices, body frame; device pointers
struct SdfShape {
    const float* grid; const float* verts;
    int nx, ny, nz, nv;
    float ox, oy, oz, cell, radius;
};
enum { MAX_SDF_SHAPES = 16 };

struct View {
    int B, NB, NJ, NC;
    int level_mode;                       // 0: rows as warp-blocked tiles in slot order (k_pgs)
                                          // 1: rows scene-major in dependency-level order (k_pgs_lv)
    int *nbodies, *njoints, *ncontacts;   // [B]
    int *overflow;                        // [1] set when a scene exceeded NC
    unsigned long long *visit_counters;   // [2] contact visits, joint visits (profiling)
    int profiling;
    int ext_pairs;                        // extension: sphere-plane, box-plane, box-box contacts (no reference counterpart)
    const SdfShape* sdf_shapes;           // extension: [MAX_SDF_SHAPES] shape table for GEOM_SDF bodies (geomA.x = shape id)
    int n_sdf;                            // shapes in the table (mesh pairs want the warp-per-scene detection kernels)
    // pair-parallel detection of extension scenes with >= 128 bodies (detect_pairs.cu); null otherwise
    int2*   pc_pairs;                     // [B][pc_cap] candidate pairs (i, j), unordered
    int*    pc_count;                     // [B] candidates appended this step (zeroed by k_prepare)
    int*    pc_hits;                      // [B] pairs that touch (zeroed by k_prepare)
    float4* pc_out;                       // [B][NC][6] narrow-phase result of each touching pair, unordered
    int     pc_cap;
    // bodies, [NB][B]
    float4 *pos;        // x y z | mass
    float4 *quat;       // x y z w
    float4 *vel;        // xdot | -
    float4 *omg;        // omega | -
    int    *flags;      // geometry type | FLAG_FIXED
    float4 *geomA;      // sphere r,-,-,- | box dim | cylinder h,r | plane p
    float4 *geomB;      // plane n
    float  *ibody;      // [9][NB][B]
    float  *ibodyinv;   // [9][NB][B]
    float4 *cproxy;     // x y z | sphere radius (0 for other shapes): narrow-phase inner-loop stream
    float4 *bext;       // boxes: half extents of the world-axis-aligned bounding box | - (narrow-phase cull)
    float4 *force;      // f | -
    float4 *torque;     // tau | -
    float4 *extf, *extt;
    int     has_ext;    // 0 until slrbs_set_external_forces is first called (k_prepare then skips ext_mode)
    int    *ext_mode;   // [B] 0 none, 1 additive (pre-step callback contribution), 2 absolute
    float  *Iw;         // [9][NB][B] world inertia
    float  *Iinvw;      // [9][NB][B]
    float4 *wl, *wa;    // PGS body accumulators sum(J^T lambda); after solve: fc, tauc
    float4 *wacc;       // [B][NB][2] scene-major accumulators of k_pgs_gacc (allocated for contexts that use it)
    float4 *brec;       // [B][NB][8] per-body record for the row builders (one 128 B line per body, written by k_prepare):
                        // x|mass, xdot|fixed, omega|-, f|-, tau|-, Iinv rows 0..2
    float4 *s_pos, *s_quat, *s_vel, *s_omg;   // RigidBodySystemState snapshot
    // joints, [NJ][B]
    int4   *jdef;       // type, body0, body1, dim
    float4 *jr0, *jq0, *jr1, *jq1;
    float4 *jlam;       // lambda 0..3
    float  *jlam4;      // lambda 4
    float4 *jrow;       // mode 0: [B/32][NJ][JROW_FIELDS][32]; mode 1: [B][JROW_FIELDS][NJ]
    int    *jpos;       // [NJ][B] position of joint slot in the level-ordered row stream (level mode)
    int    *jlev_end;   // [NJ+1][B] end offset of joint level l (1-based; entry 0 is 0)
    int    *njlev;      // [B] number of joint levels
    float  *kry;        // [7][5][NJ][B] Krylov work vectors x, r, p, b, Ap, Ar, Ax (allocated on first CG/CR solve)
    // contacts, [NC][B]
    // contacts: cids, cp, cn, cpos are indexed with atc() ([NC][B], or [B][NC] in level / colour mode)
    int2   *cids;       // body0, body1 (role ordered: sphere/box, cylinder/plane)
    float4 *cp;         // p | pene
    float4 *cn;         // n | -
    float4 *crow;       // mode 0: [B/32][NC][CROW_FIELDS][32]; mode 1: [B][CROW_FIELDS][NC]
    int    *cpos;       // [NC][B] position of contact slot in the level-ordered row stream (level mode)
    int    *cslot;      // [B][NC] inverse of cpos: contact slot at a position of the level-ordered stream (level mode)
    int    *clev_end;   // [NC+1][B] end offset of contact level l (1-based; entry 0 is 0)
    int    *nclev;      // [B] number of contact levels
    int    *lv_last;    // [NB][B] scratch: last level that touched a body
    // graph-coloured solver for scenes too large for one SM (pgs_color.cu): colours take the place of the contact levels
    int     color_mode;
    // contact row records in the 13-field compact form (208 B instead of 256 B: rr0, rr1 stored, a0d = rr0 x d and
    // a1d = -(rr1 x d) recomputed by the consumer, bit-identically); contexts served by k_pgs_pipe, SLRBS_ROWS_COMPACT
    int     crow_compact;
    unsigned long long *cmask, *cwin;   // [B][NB] colours used by a body's contacts | lowest bid among its uncoloured contacts
    int    *nclev_max;  // [1] largest colour count over the scenes (reset by k_prepare)
    // chunk table of k_pgs_pipe (pgs_pipe.cu): [B][cchunk_cap] entries `base | count << 16`, written and read by the scene's warp
    unsigned *cchunk;
    int     cchunk_cap;
};

__host__ __device__ __forceinline__ size_t at(const View& v, int slot, int scene) { return (size_t)slot * v.B + scene; }
// contact-slot arrays (cids, cp, cn, cpos): scene-minor like everything else for the thread-per-scene kernels, but SCENE-MAJOR
// ([scene][slot]) in level / colour mode, where a warp or a CTA works on ONE scene and walks consecutive slots
__host__ __device__ __forceinline__ size_t atc(const View& v, int slot, int scene) { return v.level_mode ? (size_t)scene * v.NC + slot : (size_t)slot * v.B + scene; }
__host__ __device__ __forceinline__ size_t at_brec(const View& v, int slot, int scene) { return ((size_t)scene * v.NB + slot) * 8; }
__host__ __device__ __forceinline__ size_t at_kry(const View& v, int vec, int k, int slot, int scene) { return (((size_t)vec * 5 + k) * v.NJ + slot) * v.B + scene; }
__host__ __device__ __forceinline__ size_t at_b9(const View& v, int k, int slot, int scene) { return ((size_t)k * v.NB + slot) * v.B + scene; }
// row record addressing: `pos` is the contact / joint slot in mode 0 and the position in the
// level-ordered stream (cpos / jpos of the slot) in mode 1
__host__ __device__ __forceinline__ size_t at_crow(const View& v, int f, int pos, int scene)
{
    if (v.level_mode) return ((size_t)scene * CROW_FIELDS + f) * v.NC + pos;
    return ((((size_t)(scene >> 5) * v.NC + pos) * CROW_FIELDS + f) << 5) + (scene & 31);
}
__host__ __device__ __forceinline__ size_t at_jrow(const View& v, int f, int pos, int scene)
{
    if (v.level_mode) return ((size_t)scene * JROW_FIELDS + f) * v.NJ + pos;
    return ((((size_t)(scene >> 5) * v.NJ + pos) * JROW_FIELDS + f) << 5) + (scene & 31);
}
__host__ __device__ __forceinline__ int contact_pos(const View& v, int slot, int scene) { return v.level_mode ? v.cpos[(size_t)scene * v.NC + slot] : slot; }
__host__ __device__ __forceinline__ int joint_pos(const View& v, int slot, int scene) { return v.level_mode ? v.jpos[at(v, slot, scene)] : slot; }

}  // namespace slrbs
