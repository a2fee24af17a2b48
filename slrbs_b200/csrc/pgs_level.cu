// pgs_level.cu -- level-scheduled Box-PGS: the reference's sequential Gauss-Seidel order, run in parallel.
//
// SolverBoxPGS.cpp:217-248 visits the row blocks strictly one after another. A row only reads and
// writes the accumulators of its own (non-fixed) bodies, so the result of the sequential sweep is
// fully determined by, for every body, the ORDER of the rows that touch it. k_levels gives every row
// the level  L(row) = 1 + max(L of the previous row on each of its non-fixed bodies): rows of one
// level share no dynamic body, every body still sees its rows in the original order, and processing
// level after level therefore reproduces the sequential sweep exactly (bit for bit, the row maths
// being the same functions of pgs_math.cuh that the thread-per-scene kernel uses). The marble box's
// ~450 rows fall into ~50 levels, a 1000-sphere pile's ~2800 rows into ~80.
//
// k_joint_levels  thread / scene  joint levels + stable counting sort -> jpos (row position in the level-
//                                 ordered stream) and jlev_end (level boundaries); joints are static, so this
//                                 runs only after an upload. Contact levels are assigned by the narrow-phase
//                                 warp as it emits the contacts (k_detect_warp in kernels.cu).
// k_pgs_lv   G lanes / scene  body accumulators of the scene in shared memory, row records streamed
//                             from HBM in level order (scene-major [field][pos] so the lanes of a level
//                             read consecutive records), group barrier between levels
#include "kernels.cuh"
#ifndef SLRBS_LV_LO
#define SLRBS_LV_LO 12      // warps per SM the register allocation of the "lo" variants must allow
#endif
#include <cstdlib>
#include "pgs_math.cuh"
#include "pgs_level.cuh"

namespace slrbs {

// ------------------------------------------------------------------------------------
// joint levels (joints and contacts form two separate level sequences: every sweep runs all joint
// levels, then all contact levels, which is what "all joints, then all contacts" means for a body)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_joint_levels(View v)
{
    const int scene = blockIdx.x * blockDim.x + threadIdx.x;
    if (scene >= v.B) return;
    const int nb = v.nbodies[scene], nj = v.njoints[scene];
    for (int i = 0; i < nb; ++i) v.lv_last[at(v, i, scene)] = 0;
    for (int l = 0; l <= nj; ++l) v.jlev_end[at(v, l, scene)] = 0;
    int nl = 0;
    for (int s = 0; s < nj; ++s) {
        const int4 def = v.jdef[at(v, s, scene)];
        const bool dyn0 = !(v.flags[at(v, def.y, scene)] & FLAG_FIXED), dyn1 = !(v.flags[at(v, def.z, scene)] & FLAG_FIXED);
        int L = 0;
        if (dyn0) L = max(L, v.lv_last[at(v, def.y, scene)]);
        if (dyn1) L = max(L, v.lv_last[at(v, def.z, scene)]);
        ++L;
        if (dyn0) v.lv_last[at(v, def.y, scene)] = L;
        if (dyn1) v.lv_last[at(v, def.z, scene)] = L;
        v.jpos[at(v, s, scene)] = L;
        v.jlev_end[at(v, L, scene)] += 1;
        nl = max(nl, L);
    }
    int start = 0;
    for (int l = 1; l <= nl; ++l) { const int c = v.jlev_end[at(v, l, scene)]; v.jlev_end[at(v, l, scene)] = start; start += c; }
    for (int s = 0; s < nj; ++s) {
        const int L = v.jpos[at(v, s, scene)];
        v.jpos[at(v, s, scene)] = v.jlev_end[at(v, L, scene)]++;       // afterwards jlev_end[L] is the end of level L
    }
    v.njlev[scene] = nl;
}

// ------------------------------------------------------------------------------------
// contact levels: one warp per scene. Lane 0 walks the contact list in order (the level of a contact
// depends on the levels of all earlier contacts of its bodies, a strictly sequential chain kept short by
// doing it out of shared-memory tables); the lanes stage the ids in chunks of 32 and do the counting sort.
// ------------------------------------------------------------------------------------
// Shared memory per scene decides how many scenes an SM walks at once, and the walk is a latency chain: `last` is 16 bit
// (a level number < 65535; 0xFFFF = fixed body; level mode needs max_contacts < 65535, slrbs_create), and the per-level counters live in shared memory only for the first LEVEL_HIST_CAP
// levels -- a thousand-sphere pile has ~130 -- with the scene's clev_end column in global memory as the overflow area for
// deeper chains. (NB, NC) = (1005, 4096): 4.5 KB per scene instead of 12.6 KB, so the 18 scenes per SM of a 2664-scene batch
// are resident together (before: 16 per SM, i.e. a second, nearly empty wave).
enum { LEVEL_HIST_CAP = 1024, LV_FIXED = 0xFFFF };

struct LevelHist {
    unsigned short* s; int* g; size_t B;                          // g[l * B]: the scene's clev_end column, used as scratch for l >= CAP
    __device__ __forceinline__ int get(int l) const { return l < LEVEL_HIST_CAP ? (int)s[l] : g[(size_t)l * B]; }
    __device__ __forceinline__ void set(int l, int x) const { if (l < LEVEL_HIST_CAP) s[l] = (unsigned short)x; else g[(size_t)l * B] = x; }
};

__global__ void __launch_bounds__(32) k_contact_levels(View v)
{
    extern __shared__ int lsm[];
    const int lane = threadIdx.x, scene = blockIdx.x;
    unsigned short* last = reinterpret_cast<unsigned short*>(lsm);   // [NB] last level that touched a body; LV_FIXED marks a fixed body
    unsigned short* hist_s = reinterpret_cast<unsigned short*>(last + (((size_t)v.NB + 3) & ~(size_t)3));   // [LEVEL_HIST_CAP]
    int2* idbuf = reinterpret_cast<int2*>(hist_s + LEVEL_HIST_CAP);   // [32], 8-byte aligned
    int* outbuf = reinterpret_cast<int*>(idbuf + 32);             // [32]
    const LevelHist hist{ hist_s, v.clev_end + scene, (size_t)v.B };
    const int nb = v.nbodies[scene], nc = v.ncontacts[scene];
    for (int b0 = 0; b0 < nb; b0 += 256) {                      // 8 independent flag loads in flight per lane
        int fl[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int b = b0 + u * 32 + lane; fl[u] = b < nb ? v.flags[at(v, b, scene)] : 0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int b = b0 + u * 32 + lane; if (b < nb) last[b] = (fl[u] & FLAG_FIXED) ? (unsigned short)LV_FIXED : (unsigned short)0; }
    }
    // a chain can be at most nc levels deep; counters of levels the walk never reaches are never read
    for (int l = lane; l <= min(nc, LEVEL_HIST_CAP - 1); l += 32) hist_s[l] = 0;
    for (int l = LEVEL_HIST_CAP + lane; l <= nc; l += 32) hist.g[(size_t)l * hist.B] = 0;
    __syncwarp();
    // The level of a contact depends on the levels of all earlier contacts of its two bodies: a sequential chain, walked
    // by lane 0 out of shared memory (ids staged by the warp in chunks of 32; two dependent shared-memory accesses per
    // contact). The stable rank inside a level (counting sort) is off that chain: per chunk all lanes match the lanes
    // of equal level and count the lower ones.
    int nlev = 0;
    int2 nxt = make_int2(0, 0);
    if (lane < nc) nxt = v.cids[atc(v, lane, scene)];
    for (int c0 = 0; c0 < nc; c0 += 32) {
        const int m = min(32, nc - c0);
        idbuf[lane] = nxt;
        if (c0 + 32 + lane < nc) nxt = v.cids[atc(v, c0 + 32 + lane, scene)];   // next chunk's ids, in flight during the chain
        __syncwarp();
        if (lane == 0) {
            if (m == 32) {                                       // full chunk: ids in registers, only `last` is on the chain
                int2 r[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) r[k] = idbuf[k];
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const int l0 = last[r[k].x], l1 = last[r[k].y];
                    const int L = max(l0 == LV_FIXED ? 0 : l0, l1 == LV_FIXED ? 0 : l1) + 1;   // fixed bodies do not order anything
                    if (l0 != LV_FIXED) last[r[k].x] = (unsigned short)L;
                    if (l1 != LV_FIXED) last[r[k].y] = (unsigned short)L;
                    outbuf[k] = L;
                }
            } else {
                for (int k = 0; k < m; ++k) {
                    const int2 ids = idbuf[k];
                    const int l0 = last[ids.x], l1 = last[ids.y];
                    const int L = max(l0 == LV_FIXED ? 0 : l0, l1 == LV_FIXED ? 0 : l1) + 1;
                    if (l0 != LV_FIXED) last[ids.x] = (unsigned short)L;
                    if (l1 != LV_FIXED) last[ids.y] = (unsigned short)L;
                    outbuf[k] = L;
                }
            }
        }
        __syncwarp();
        const unsigned act = __ballot_sync(0xffffffffu, lane < m);
        if (lane < m) {
            const int myL = outbuf[lane];
            const unsigned same = __match_any_sync(act, myL);
            const int base = hist.get(myL);
            __syncwarp(act);
            if (lane == __ffs(same) - 1) hist.set(myL, base + __popc(same));
            v.cpos[atc(v, c0 + lane, scene)] = (int)(((unsigned)myL << 16) | (unsigned)(base + __popc(same & ((1u << lane) - 1u))));
            nlev = max(nlev, myL);
        }
        __syncwarp();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nlev = max(nlev, __shfl_xor_sync(0xffffffffu, nlev, d));
    // exclusive prefix of the per-level counts -> level starts (shared counters, in place) and level ends (global). A level
    // past the shared counters had its count in clev_end itself, which now receives its end; its start is read back below
    // as the end of the level before it.
    int carry = 0;
    for (int l0 = 1; l0 <= nlev; l0 += 32) {
        const int l = l0 + lane;
        const int c = l <= nlev ? hist.get(l) : 0;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (l <= nlev) {
            if (l < LEVEL_HIST_CAP) hist_s[l] = (unsigned short)(carry + incl - c);   // level start
            v.clev_end[at(v, l, scene)] = carry + incl;                               // level end (for l >= CAP the start is the previous level's end)
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    // (level, rank) -> position, 8 independent loads in flight per lane (one load per trip was 58 % of this kernel's time)
    for (int c0 = 0; c0 < nc; c0 += 256) {
        int pk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int c = c0 + u * 32 + lane; pk[u] = c < nc ? v.cpos[atc(v, c, scene)] : 0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * 32 + lane;
            if (c < nc) {
                const int L = (int)((unsigned)pk[u] >> 16);
                const int start = L < LEVEL_HIST_CAP ? (int)hist_s[L] : v.clev_end[at(v, L - 1, scene)];
                const int pos = start + (pk[u] & 0xffff);
                v.cpos[atc(v, c, scene)] = pos;
                v.cslot[(size_t)scene * v.NC + pos] = c;
            }
        }
    }
    if (lane == 0) v.nclev[scene] = nlev;
}

// ------------------------------------------------------------------------------------
// solve
// ------------------------------------------------------------------------------------
// CTA size: one warp holding 32 / G scenes (small CTAs pack shared memory tightly: a 1000-body scene needs 32 KB of
// accumulators, so 7 scenes fit an SM), or G = 64 / 128 lanes = 2 / 4 warps on ONE scene whose levels are wider
// than a warp (block barrier between levels instead of a warp barrier).
template <int G> struct LvCfg { enum { THREADS = G < 32 ? 32 : G, SPC = THREADS / G }; };

#define ld_w Acc<W24>::ld
#define st_w Acc<W24>::st

template <int G>
__device__ __forceinline__ void group_sync(unsigned mask)
{
    if (G > 32) __syncthreads();
    else if (G == 32) __syncwarp();
    else __syncwarp(mask);
}

// MINB = warps per SM the register allocation must allow: 16 (128 registers, a few spills) wins when the batch
// can fill 16 warps per SM, 12 (162 registers, no spills) when it cannot and single-warp latency is what matters.
template <int G, int MINB, bool W24>
__global__ void __launch_bounds__(LvCfg<G>::THREADS, MINB) k_pgs_lv(View v, float dt, int iters, float mu)
{
    extern __shared__ float4 smem_w4[];
    float* smem_w = reinterpret_cast<float*>(smem_w4);
    constexpr int SPC = LvCfg<G>::SPC;                           // scenes per CTA
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    const int scene = blockIdx.x * SPC + grp;
    const unsigned mask = (G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
    if (scene >= v.B) return;                                    // whole groups (for G > 32: whole CTAs) leave together
    float* w = smem_w + (size_t)grp * v.NB * Acc<W24>::FLOATS;
    const int nb = v.nbodies[scene], nj = v.njoints[scene], nc = v.ncontacts[scene];
    const int njl = nj > 0 ? v.njlev[scene] : 0, ncl = nc > 0 ? v.nclev[scene] : 0;   // no rows, no levels (stale tables are never walked)

    for (int i = gl; i < Acc<W24>::FLOATS * nb; i += G) w[i] = 0.f;
    group_sync<G>(mask);
    // joints keep last step's lambda (Hinge.cpp:33-63 never zeroes it): seed w with it, level by level
    for (int lev = 1, p0 = 0; lev <= njl; ++lev) {
        const int p1 = v.jlev_end[at(v, lev, scene)];
        for (int p = p0 + gl; p < p1; p += G) {
            const JointRow j = load_joint_head(v, p, scene);
            const int slot = __float_as_int(__ldg(&v.jrow[at_jrow(v, 16, p, scene)]).w);
            const size_t ji = at(v, slot, scene);
            const float4 lam03 = v.jlam[ji];
            const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
            V3 wl, wa;
            if (j.id0 >= 0) { ld_w(w, j.id0, wl, wa); joint_seed(j, l, 0, wl, wa); st_w(w, j.id0, wl, wa); }
            if (j.id1 >= 0) { const int b = j.id1 & 0x7fffffff; ld_w(w, b, wl, wa); joint_seed(j, l, 1, wl, wa); st_w(w, b, wl, wa); }
        }
        p0 = p1;
        group_sync<G>(mask);
    }

    for (int it = 0; it < iters; ++it) {                          // SolverBoxPGS.cpp:217-248
        for (int lev = 1, p0 = 0; lev <= njl; ++lev) {
            const int p1 = v.jlev_end[at(v, lev, scene)];
            for (int p = p0 + gl; p < p1; p += G) {
                JointRec r;
                load_joint_rec(v, p, scene, r);
                const size_t ji = at(v, joint_slot(r), scene);
                const float4 lam03 = v.jlam[ji];
                float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
                const bool dyn0 = r.h.id0 >= 0, dyn1 = r.h.id1 >= 0;
                const int b0 = r.h.id0 & 0x7fffffff, b1 = r.h.id1 & 0x7fffffff;
                V3 w0l = mk3(0.f, 0.f, 0.f), w0a = w0l, w1l = w0l, w1a = w0l;
                if (dyn0) ld_w(w, b0, w0l, w0a);
                if (dyn1) ld_w(w, b1, w1l, w1a);
                joint_solve(r, l, dyn0, dyn1, w0l, w0a, w1l, w1a);
                if (dyn0) st_w(w, b0, w0l, w0a);
                if (dyn1) st_w(w, b1, w1l, w1a);
                v.jlam[ji] = make_float4(l[0], l[1], l[2], l[3]);
                v.jlam4[ji] = l[4];
            }
            p0 = p1;
            group_sync<G>(mask);
        }
        for (int lev = 1, p0 = 0; lev <= ncl; ++lev) {
            const int p1 = v.clev_end[at(v, lev, scene)];
            if (G >= 32 && v.B >= 256) {   // (measured: pays only for wide scenes in batches whose rows stream from HBM)
                // pull the rows this lane will solve in the NEXT level (or the first level of the next sweep) towards L2
                // while the current level is solved: the row stream is far larger than L2 and would otherwise pay an
                // HBM round trip at the start of every level
                int q0 = p1, q1 = p1;
                if (lev < ncl) q1 = v.clev_end[at(v, lev + 1, scene)];
                else if (it + 1 < iters) { q0 = 0; q1 = v.clev_end[at(v, 1, scene)]; }
                // the rows of a level are contiguous per field ([scene][field][pos]): one prefetch per 128 B line
                if (q1 > q0) prefetch_rows_l2(v.crow + (size_t)scene * CROW_FIELDS * (size_t)v.NC, (unsigned)v.NC, q0, q1, CROW_LAMBDA, gl, G);
            }
            for (int p = p0 + gl; p < p1; p += G) {
                float4 F[CROW_FIELDS];
#pragma unroll
                for (int f = 0; f < CROW_LAMBDA; ++f) F[f] = __ldg(&v.crow[at_crow(v, f, p, scene)]);
                const size_t li = at_crow(v, CROW_LAMBDA, p, scene);
                const float4 lam = v.crow[li];
                const ContactIds c = contact_ids(F);
                V3 w0l = mk3(0.f, 0.f, 0.f), w0a = w0l, w1l = w0l, w1a = w0l;
                if (c.dyn0) ld_w(w, c.id0, w0l, w0a);
                if (c.dyn1) ld_w(w, c.id1, w1l, w1a);
                const float4 out = contact_solve(F, lam, mu, c.dyn0, c.dyn1, w0l, w0a, w1l, w1a);
                if (c.dyn0) st_w(w, c.id0, w0l, w0a);
                if (c.dyn1) st_w(w, c.id1, w1l, w1a);
                v.crow[li] = out;
            }
            p0 = p1;
            group_sync<G>(mask);
        }
    }

    // constraint forces (RigidBodySystem.cpp:185-220) in the same per-body order: w becomes fc / tauc
    for (int i = gl; i < Acc<W24>::FLOATS * nb; i += G) w[i] = 0.f;
    group_sync<G>(mask);
    for (int lev = 1, p0 = 0; lev <= njl; ++lev) {
        const int p1 = v.jlev_end[at(v, lev, scene)];
        for (int p = p0 + gl; p < p1; p += G) {
            const JointRow j = load_joint_head(v, p, scene);
            const float4 tail = __ldg(&v.jrow[at_jrow(v, 16, p, scene)]);
            const size_t ji = at(v, __float_as_int(tail.w), scene);
            const float4 lam03 = v.jlam[ji];
            const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
            float f0[6], f1[6];
            joint_force(j, __float_as_int(tail.z), l, dt, f0, f1);
            V3 wl, wa;
            if (j.id0 >= 0) { ld_w(w, j.id0, wl, wa); st_w(w, j.id0, wl + mk3(f0[0], f0[1], f0[2]), wa + mk3(f0[3], f0[4], f0[5])); }
            if (j.id1 >= 0) { const int b = j.id1 & 0x7fffffff; ld_w(w, b, wl, wa); st_w(w, b, wl + mk3(f1[0], f1[1], f1[2]), wa + mk3(f1[3], f1[4], f1[5])); }
        }
        p0 = p1;
        group_sync<G>(mask);
    }
    // the reference also adds joint forces to FIXED bodies (:189-192, harmless: they do not integrate).
    // Levels do not separate rows that share a fixed body, so one lane adds those in joint order.
    if (gl == 0) {
        for (int s = 0; s < nj; ++s) {
            const int4 def = v.jdef[at(v, s, scene)];
            const bool fix0 = v.flags[at(v, def.y, scene)] & FLAG_FIXED, fix1 = v.flags[at(v, def.z, scene)] & FLAG_FIXED;
            if (!fix0 && !fix1) continue;
            const int p = v.jpos[at(v, s, scene)];
            const JointRow j = load_joint_head(v, p, scene);
            const size_t ji = at(v, s, scene);
            const float4 lam03 = v.jlam[ji];
            const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
            float f0[6], f1[6];
            joint_force(j, def.w, l, dt, f0, f1);
            V3 wl, wa;
            if (fix0) { ld_w(w, def.y, wl, wa); st_w(w, def.y, wl + mk3(f0[0], f0[1], f0[2]), wa + mk3(f0[3], f0[4], f0[5])); }
            if (fix1) { ld_w(w, def.z, wl, wa); st_w(w, def.z, wl + mk3(f1[0], f1[1], f1[2]), wa + mk3(f1[3], f1[4], f1[5])); }
        }
    }
    group_sync<G>(mask);
    for (int lev = 1, p0 = 0; lev <= ncl; ++lev) {
        const int p1 = v.clev_end[at(v, lev, scene)];
        for (int p = p0 + gl; p < p1; p += G) {
            float4 F[9];
#pragma unroll
            for (int f = 0; f < 9; ++f) F[f] = __ldg(&v.crow[at_crow(v, f, p, scene)]);
            const float4 lam = v.crow[at_crow(v, CROW_LAMBDA, p, scene)];
            const ContactIds c = contact_ids(F);
            V3 fl, fa, wl, wa;
            if (c.dyn0) { contact_force(F, lam, dt, 0, fl, fa); ld_w(w, c.id0, wl, wa); st_w(w, c.id0, wl + fl, wa + fa); }
            if (c.dyn1) { contact_force(F, lam, dt, 1, fl, fa); ld_w(w, c.id1, wl, wa); st_w(w, c.id1, wl + fl, wa + fa); }
        }
        p0 = p1;
        group_sync<G>(mask);
    }
    for (int i = gl; i < nb; i += G) {
        V3 fl, fa;
        ld_w(w, i, fl, fa);
        v.wl[at(v, i, scene)] = mk4(fl, 0.f);
        v.wa[at(v, i, scene)] = mk4(fa, 0.f);
    }
    if (v.profiling && gl == 0) {
        atomicAdd(&v.visit_counters[0], (unsigned long long)nc * iters);
        atomicAdd(&v.visit_counters[1], (unsigned long long)nj * iters);
    }
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
size_t contact_levels_smem(int NB, int)
{
    return (((size_t)NB + 3) & ~(size_t)3) * 2 + (size_t)LEVEL_HIST_CAP * 2 + 32 * 8 + 32 * 4;
}

void launch_contact_levels(const View& v, cudaStream_t s)
{
    k_contact_levels<<<(unsigned)v.B, 32, contact_levels_smem(v.NB, v.NC), s>>>(v);
}

void launch_joint_levels(const View& v, cudaStream_t s)
{
    k_joint_levels<<<(unsigned)((v.B + 127) / 128), 128, 0, s>>>(v);
}

// padded accumulators unless they would keep the SM below 12 warps
static bool lv_pack24(int threads, int spc, int NB)
{
    return ((size_t)spc * NB * 32 + 1024) * (size_t)(12 * 32 / threads) > 227u * 1024u;
}
size_t level_smem_bytes(int threads, int spc, int NB) { return (size_t)spc * NB * (lv_pack24(threads, spc, NB) ? 24 : 32); }

template <int G>
static void launch_lv(const View& v, float dt, int iters, float mu, cudaStream_t s)
{
    constexpr int threads = LvCfg<G>::THREADS, spc = LvCfg<G>::SPC;
    const bool w24 = lv_pack24(threads, spc, v.NB);
    const size_t smem = level_smem_bytes(threads, spc, v.NB);
    const unsigned blocks = (unsigned)((v.B + spc - 1) / spc);
    constexpr int lo = SLRBS_LV_LO * 32 / threads, hi = 16 * 32 / threads;       // 12 / 16 warps per SM
    const bool many = (size_t)blocks * threads >= 148u * 8u * 32u;
    if (w24) {
        // shared memory already keeps the SM below 12 warps here: the variant with the larger register budget (no spills)
        static int force_hi = -1;
        if (force_hi < 0) { const char* e = getenv("SLRBS_PGS_W24_HI"); force_hi = e ? atoi(e) : 0; }
        if (many && force_hi) k_pgs_lv<G, hi, true><<<blocks, threads, smem, s>>>(v, dt, iters, mu);
        else k_pgs_lv<G, lo, true><<<blocks, threads, smem, s>>>(v, dt, iters, mu);
    } else {
        static int minb = -1;                                   // SLRBS_PGS_MINB: 1 = always the 170-register variant, 2 = always the 128-register one
        if (minb < 0) { const char* e = getenv("SLRBS_PGS_MINB"); minb = e ? atoi(e) : 0; }
        // measured (profiles/README.md): scenes of ~100-500 bodies (G = 16) run faster under the 128-register cap at every
        // batch size (16 instead of 12 warps per SM), tiny scenes (G = 4, 8) and single wide scenes faster without it
        const bool auto_hi = G == 16 ? true : (G == 32 ? many : false);
        const bool use_hi = minb == 1 ? false : (minb == 2 ? true : auto_hi);
        if (use_hi) k_pgs_lv<G, hi, false><<<blocks, threads, smem, s>>>(v, dt, iters, mu);
        else k_pgs_lv<G, lo, false><<<blocks, threads, smem, s>>>(v, dt, iters, mu);
    }
}

template <int G>
static void init_lv()
{
    constexpr int threads = LvCfg<G>::THREADS;
    cudaFuncSetAttribute(k_pgs_lv<G, SLRBS_LV_LO * 32 / threads, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_pgs_lv<G, 16 * 32 / threads, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_pgs_lv<G, SLRBS_LV_LO * 32 / threads, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_pgs_lv<G, 16 * 32 / threads, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}
// per-device kernel attributes (called from slrbs_create with the device current)
void init_level_kernels()
{
    init_lv<4>(); init_lv<8>(); init_lv<16>(); init_lv<32>(); init_lv<64>(); init_lv<128>();
    init_split_kernels();
    init_pipe_kernels();
    init_gacc_kernels();
    init_stage_kernels();
    cudaFuncSetAttribute(k_contact_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

int level_group_size(const View& v, int requested)
{
    int g = requested;
    if (g != 4 && g != 8 && g != 16 && g != 32 && g != 64 && g != 128) g = v.NB < 24 ? 4 : (v.NB <= 96 ? 8 : (v.NB <= 512 ? 16 : ((v.B <= 296 && v.NJ > 0) ? 64 : 32)));   // few wide scenes with joints: 2 warps each (contact-only scenes: k_pgs_pipe, one warp, is faster at every batch size -- 0.70 / 0.85 against 0.73 / 1.13 ms for 16 / 296 thousand-body scenes)
    // the accumulators of the 32 / g scenes of a warp must fit in shared memory
    while (g < 32 && (size_t)(32 / g) * v.NB * 24 > 200 * 1024) g <<= 1;
    return g;
}

bool level_mode_fits(int max_bodies) { return (size_t)max_bodies * 24 <= 200 * 1024; }

// The side-split kernel (pgs_split.cu: two lanes per contact row) is opt-in (SLRBS_PGS_SPLIT=1) for the wide-scene groups
// G = 32 / 64. Measured on marble_box_1k x 2664 it is slower than one lane per row (6.3 vs 5.1 ms): the level step is bound
// by the latency of its critical path, not by issue slots, and a second warp per scene shortens the chain by less than
// its barrier adds (profiles/README.md).
bool level_split_wanted(const View& v, int group)
{
    const char* e = getenv("SLRBS_PGS_SPLIT");                   // read when a context is created
    const int on = e ? atoi(e) : 0;
    return on && (group == 32 || group == 64) && (size_t)v.NB * 24 <= 200 * 1024;
}

void launch_pgs_level(const View& v, float dt, int iters, float mu, int group, int variant, cudaStream_t s)
{
    if (variant == 1) { launch_pgs_split(v, dt, iters, mu, group, s); return; }
    if (variant == 2) { launch_pgs_pipe(v, dt, iters, mu, s); return; }
    if (variant == 4) { launch_pgs_gacc(v, dt, iters, mu, s); return; }
    if (variant == 5) { launch_pgs_stage(v, dt, iters, mu, s); return; }
    switch (group) {
    case 4:  launch_lv<4>(v, dt, iters, mu, s); break;
    case 8:  launch_lv<8>(v, dt, iters, mu, s); break;
    case 16: launch_lv<16>(v, dt, iters, mu, s); break;
    case 64: launch_lv<64>(v, dt, iters, mu, s); break;
    case 128: launch_lv<128>(v, dt, iters, mu, s); break;
    default: launch_lv<32>(v, dt, iters, mu, s); break;
    }
}

}  // namespace slrbs
