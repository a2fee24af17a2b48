// pgs_color.cu -- graph-coloured Box-PGS for scenes too large for one SM.
//
// The scene-resident kernels (pgs_level.cu, pgs_pipe.cu) keep a scene's body accumulators in one SM's shared memory and
// reproduce the reference's sequential row order (SolverBoxPGS.cpp:217-248). A scene of more than ~8500 bodies does not
// fit, and its dependency levels would be hundreds of grid-wide steps deep. Here the ORDER is given up, as north_star
// allows for the large-scene configurations ("graph-coloured PGS validated on constraint residual rather than
// trajectory"): contacts are coloured so that two contacts of one colour share no dynamic body, every sweep visits the
// colours in turn, and all rows of a colour are solved in parallel by the whole GPU. The per-row maths is unchanged --
// the same contact_solve / joint_solve of pgs_math.cuh (SolverBoxPGS.cpp:94-113, 115-118, 49-82) -- only the
// Gauss-Seidel sequence differs, so impulses are validated on the box-LCP residual and on penetration
// (tests/test_gpu_color.py), not bit for bit.
//
// k_contact_colors   CTA per scene. Jones-Plassmann rounds: every uncoloured contact bids a hashed priority on its
//                    dynamic bodies (64-bit atomicMin); a contact that holds both of its bodies takes the lowest colour
//                    neither body has used yet (a 64-bit colour mask per body). A stable counting sort by colour then
//                    fills the SAME tables the level schedule uses (cpos, cslot, clev_end, nclev), so the row builder
//                    (k_contact_rows_lv) writes the row records colour by colour: the lanes of a colour stream
//                    consecutive records. Deterministic: priorities depend on the contact slot only, and the slots
//                    are in the reference's (i, j) order.
// k_pgs_color        persistent cooperative kernel, all SMs. Body accumulators w = sum J^T lambda live in HBM / L2
//                    (wl, wa; read with ld.global.cg since other SMs wrote them one colour earlier), one grid-wide barrier
//                    per colour. Joints keep their dependency levels (exact order; scenes this large have few).
#include "kernels.cuh"
#include "pgs_math.cuh"

#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace slrbs {

enum { COLOR_THREADS = 1024, MAX_COLORS = 64 };

__device__ __forceinline__ unsigned mix32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(COLOR_THREADS) k_contact_colors(View v)
{
    __shared__ int s_remaining, s_ncol, s_hist[MAX_COLORS], s_run[MAX_COLORS];
    __shared__ unsigned short s_wcnt[COLOR_THREADS / 32][MAX_COLORS];
    const int scene = blockIdx.x, tid = threadIdx.x, T = COLOR_THREADS, lane = tid & 31, warp = tid >> 5;
    const int nb = v.nbodies[scene], nc = v.ncontacts[scene];
    unsigned long long* mask = v.cmask + (size_t)scene * v.NB;    // colours taken by the coloured contacts of a body
    unsigned long long* win = v.cwin + (size_t)scene * v.NB;      // lowest bid of the uncoloured contacts of a body
    int* color = v.cpos;                                          // (atc) the colour while colouring, then the position
    for (int b = tid; b < nb; b += T) { mask[b] = 0ull; win[b] = ~0ull; }
    for (int c = tid; c < nc; c += T) color[atc(v, c, scene)] = -1;
    if (tid == 0) { s_remaining = nc; s_ncol = 0; }
    if (tid < MAX_COLORS) { s_hist[tid] = 0; s_run[tid] = 0; }
    __syncthreads();
    while (s_remaining > 0) {
        // bids
        for (int c = tid; c < nc; c += T) {
            if (color[atc(v, c, scene)] >= 0) continue;
            const int2 ids = v.cids[atc(v, c, scene)];
            const unsigned long long bid = ((unsigned long long)mix32((unsigned)c) << 32) | (unsigned)c;
            if (!(v.flags[at(v, ids.x, scene)] & FLAG_FIXED)) atomicMin(&win[ids.x], bid);
            if (!(v.flags[at(v, ids.y, scene)] & FLAG_FIXED)) atomicMin(&win[ids.y], bid);
        }
        __syncthreads();
        // contacts that hold both bodies take a colour; they share no body with each other, so the masks they read are final
        int done = 0;
        for (int c = tid; c < nc; c += T) {
            if (color[atc(v, c, scene)] >= 0) continue;
            const int2 ids = v.cids[atc(v, c, scene)];
            const unsigned long long bid = ((unsigned long long)mix32((unsigned)c) << 32) | (unsigned)c;
            const bool dyn0 = !(v.flags[at(v, ids.x, scene)] & FLAG_FIXED), dyn1 = !(v.flags[at(v, ids.y, scene)] & FLAG_FIXED);
            if ((dyn0 && __ldcg(&win[ids.x]) != bid) || (dyn1 && __ldcg(&win[ids.y]) != bid)) continue;
            const unsigned long long used = (dyn0 ? __ldcg(&mask[ids.x]) : 0ull) | (dyn1 ? __ldcg(&mask[ids.y]) : 0ull);
            int k = __ffsll((long long)~used) - 1;
            if (k < 0) { k = MAX_COLORS - 1; atomicOr(v.overflow, 8); }      // a body with 64 contacts of distinct colours
            color[atc(v, c, scene)] = k;
            if (dyn0) { atomicOr(&mask[ids.x], 1ull << k); atomicExch(&win[ids.x], ~0ull); }
            if (dyn1) { atomicOr(&mask[ids.y], 1ull << k); atomicExch(&win[ids.y], ~0ull); }
            atomicAdd(&s_hist[k], 1);
            ++done;
        }
        if (done) atomicSub(&s_remaining, done);
        __syncthreads();
    }
    // colour starts
    if (tid == 0) {
        int start = 0, ncol = 0;
        for (int k = 0; k < MAX_COLORS; ++k) {
            const int cnt = s_hist[k];
            s_hist[k] = start;
            start += cnt;
            if (cnt) ncol = k + 1;
            v.clev_end[at(v, k + 1, scene)] = start;               // colour k is "level" k + 1 of the row stream
        }
        s_ncol = ncol;
        v.nclev[scene] = ncol;
        if (v.nclev_max) atomicMax(v.nclev_max, ncol);
    }
    __syncthreads();
    // stable counting sort: position = start of the colour + number of earlier contacts of that colour
    for (int c0 = 0; c0 < nc; c0 += T) {
        const int c = c0 + tid;
        const int k = c < nc ? color[atc(v, c, scene)] : -1;
        for (int q = lane; q < MAX_COLORS; q += 32) s_wcnt[warp][q] = 0;
        __syncwarp();
        const unsigned act = __ballot_sync(0xffffffffu, k >= 0);
        int rank = 0;
        if (k >= 0) {
            const unsigned same = __match_any_sync(act, k);
            rank = __popc(same & ((1u << lane) - 1u));
            if (lane == __ffs(same) - 1) s_wcnt[warp][k] = (unsigned short)__popc(same);
        }
        __syncthreads();
        if (k >= 0) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += s_wcnt[w][k];
            const int pos = s_hist[k] + s_run[k] + before + rank;
            color[atc(v, c, scene)] = pos;
            v.cslot[(size_t)scene * v.NC + pos] = c;
        }
        __syncthreads();
        if (tid < MAX_COLORS) { int tot = 0; for (int w = 0; w < COLOR_THREADS / 32; ++w) tot += s_wcnt[w][tid]; s_run[tid] += tot; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------
// solve
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void ldw(const View& v, int body, int scene, V3& l, V3& a)
{
    const size_t i = at(v, body, scene);
    l = mk3(__ldcg(&v.wl[i])); a = mk3(__ldcg(&v.wa[i]));
}
__device__ __forceinline__ void stw(const View& v, int body, int scene, const V3& l, const V3& a)
{
    const size_t i = at(v, body, scene);
    __stcg(&v.wl[i], mk4(l, 0.f)); __stcg(&v.wa[i], mk4(a, 0.f));
}

__global__ void __launch_bounds__(256, 2) k_pgs_color(View v, float dt, int iters, float mu)
{
    cg::grid_group grid = cg::this_grid();
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsize = (size_t)gridDim.x * blockDim.x;
    const size_t nbt = (size_t)v.NB * v.B;
    const int maxcol = *v.nclev_max;
    int maxjl = 0;
    if (v.NJ > 0) for (int s = 0; s < v.B; ++s) maxjl = max(maxjl, v.njoints[s] > 0 ? v.njlev[s] : 0);

    // rows of "level" lev of every scene, dealt to the whole grid: f(scene, position)
    auto for_rows = [&](const int* lev_end, const int* nlev, const int* count, int lev, auto&& f) {
        for (int scene = 0; scene < v.B; ++scene) {
            if (count[scene] <= 0 || lev > nlev[scene]) continue;
            const int p0 = lev > 1 ? lev_end[at(v, lev - 1, scene)] : 0, p1 = lev_end[at(v, lev, scene)];
            for (size_t p = p0 + gtid; p < (size_t)p1; p += gsize) f(scene, (int)p);
        }
    };

    for (size_t i = gtid; i < nbt; i += gsize) { v.wl[i] = make_float4(0.f, 0.f, 0.f, 0.f); v.wa[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
    grid.sync();
    // joints keep last step's lambda (Hinge.cpp:33-63 never zeroes it): seed w with it, level by level
    for (int lev = 1; lev <= maxjl; ++lev) {
        for_rows(v.jlev_end, v.njlev, v.njoints, lev, [&](int scene, int p) {
            const JointRow j = load_joint_head(v, p, scene);
            const int slot = __float_as_int(__ldg(&v.jrow[at_jrow(v, 16, p, scene)]).w);
            const size_t ji = at(v, slot, scene);
            const float4 lam03 = v.jlam[ji];
            const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
            V3 wl, wa;
            if (j.id0 >= 0) { ldw(v, j.id0, scene, wl, wa); joint_seed(j, l, 0, wl, wa); stw(v, j.id0, scene, wl, wa); }
            if (j.id1 >= 0) { const int b = j.id1 & 0x7fffffff; ldw(v, b, scene, wl, wa); joint_seed(j, l, 1, wl, wa); stw(v, b, scene, wl, wa); }
        });
        grid.sync();
    }

    for (int it = 0; it < iters; ++it) {                          // SolverBoxPGS.cpp:217-248, colours instead of row order
        for (int lev = 1; lev <= maxjl; ++lev) {
            for_rows(v.jlev_end, v.njlev, v.njoints, lev, [&](int scene, int p) {
                JointRec r;
                load_joint_rec(v, p, scene, r);
                const size_t ji = at(v, joint_slot(r), scene);
                const float4 lam03 = __ldcg(&v.jlam[ji]);
                float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, __ldcg(&v.jlam4[ji]) };
                const bool dyn0 = r.h.id0 >= 0, dyn1 = r.h.id1 >= 0;
                const int b0 = r.h.id0 & 0x7fffffff, b1 = r.h.id1 & 0x7fffffff;
                V3 w0l = mk3(0.f, 0.f, 0.f), w0a = w0l, w1l = w0l, w1a = w0l;
                if (dyn0) ldw(v, b0, scene, w0l, w0a);
                if (dyn1) ldw(v, b1, scene, w1l, w1a);
                joint_solve(r, l, dyn0, dyn1, w0l, w0a, w1l, w1a);
                if (dyn0) stw(v, b0, scene, w0l, w0a);
                if (dyn1) stw(v, b1, scene, w1l, w1a);
                v.jlam[ji] = make_float4(l[0], l[1], l[2], l[3]);
                v.jlam4[ji] = l[4];
            });
            grid.sync();
        }
        for (int col = 1; col <= maxcol; ++col) {
            for_rows(v.clev_end, v.nclev, v.ncontacts, col, [&](int scene, int p) {
                float4 F[CROW_FIELDS];
#pragma unroll
                for (int f = 0; f < CROW_LAMBDA; ++f) F[f] = __ldg(&v.crow[at_crow(v, f, p, scene)]);
                const size_t li = at_crow(v, CROW_LAMBDA, p, scene);
                const float4 lam = __ldcg(&v.crow[li]);
                const ContactIds c = contact_ids(F);
                V3 w0l = mk3(0.f, 0.f, 0.f), w0a = w0l, w1l = w0l, w1a = w0l;
                if (c.dyn0) ldw(v, c.id0, scene, w0l, w0a);
                if (c.dyn1) ldw(v, c.id1, scene, w1l, w1a);
                const float4 out = contact_solve(F, lam, mu, c.dyn0, c.dyn1, w0l, w0a, w1l, w1a);
                if (c.dyn0) stw(v, c.id0, scene, w0l, w0a);
                if (c.dyn1) stw(v, c.id1, scene, w1l, w1a);
                v.crow[li] = out;
            });
            grid.sync();
        }
    }

    // constraint forces (RigidBodySystem.cpp:185-220): w becomes fc / tauc, summed per body in colour order
    for (size_t i = gtid; i < nbt; i += gsize) { v.wl[i] = make_float4(0.f, 0.f, 0.f, 0.f); v.wa[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
    grid.sync();
    for (int lev = 1; lev <= maxjl; ++lev) {
        for_rows(v.jlev_end, v.njlev, v.njoints, lev, [&](int scene, int p) {
            const JointRow j = load_joint_head(v, p, scene);
            const float4 tail = __ldg(&v.jrow[at_jrow(v, 16, p, scene)]);
            const size_t ji = at(v, __float_as_int(tail.w), scene);
            const float4 lam03 = __ldcg(&v.jlam[ji]);
            const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, __ldcg(&v.jlam4[ji]) };
            float f0[6], f1[6];
            joint_force(j, __float_as_int(tail.z), l, dt, f0, f1);
            V3 wl, wa;
            if (j.id0 >= 0) { ldw(v, j.id0, scene, wl, wa); stw(v, j.id0, scene, wl + mk3(f0[0], f0[1], f0[2]), wa + mk3(f0[3], f0[4], f0[5])); }
            if (j.id1 >= 0) { const int b = j.id1 & 0x7fffffff; ldw(v, b, scene, wl, wa); stw(v, b, scene, wl + mk3(f1[0], f1[1], f1[2]), wa + mk3(f1[3], f1[4], f1[5])); }
        });
        grid.sync();
    }
    // the reference also adds joint forces to FIXED bodies (:189-192, harmless: they do not integrate); levels do not
    // separate rows that share a fixed body, so one thread per scene adds those in joint order
    if (v.NJ > 0) {
        for (size_t scene = gtid; scene < (size_t)v.B; scene += gsize) {
            const int sc = (int)scene, nj = v.njoints[sc];
            for (int s = 0; s < nj; ++s) {
                const int4 def = v.jdef[at(v, s, sc)];
                const bool fix0 = v.flags[at(v, def.y, sc)] & FLAG_FIXED, fix1 = v.flags[at(v, def.z, sc)] & FLAG_FIXED;
                if (!fix0 && !fix1) continue;
                const JointRow j = load_joint_head(v, v.jpos[at(v, s, sc)], sc);
                const size_t ji = at(v, s, sc);
                const float4 lam03 = __ldcg(&v.jlam[ji]);
                const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, __ldcg(&v.jlam4[ji]) };
                float f0[6], f1[6];
                joint_force(j, def.w, l, dt, f0, f1);
                V3 wl, wa;
                if (fix0) { ldw(v, def.y, sc, wl, wa); stw(v, def.y, sc, wl + mk3(f0[0], f0[1], f0[2]), wa + mk3(f0[3], f0[4], f0[5])); }
                if (fix1) { ldw(v, def.z, sc, wl, wa); stw(v, def.z, sc, wl + mk3(f1[0], f1[1], f1[2]), wa + mk3(f1[3], f1[4], f1[5])); }
            }
        }
        grid.sync();
    }
    for (int col = 1; col <= maxcol; ++col) {
        for_rows(v.clev_end, v.nclev, v.ncontacts, col, [&](int scene, int p) {
            float4 F[9];
#pragma unroll
            for (int f = 0; f < 9; ++f) F[f] = __ldg(&v.crow[at_crow(v, f, p, scene)]);
            const float4 lam = __ldcg(&v.crow[at_crow(v, CROW_LAMBDA, p, scene)]);
            const ContactIds c = contact_ids(F);
            V3 fl, fa, wl, wa;
            if (c.dyn0) { contact_force(F, lam, dt, 0, fl, fa); ldw(v, c.id0, scene, wl, wa); stw(v, c.id0, scene, wl + fl, wa + fa); }
            if (c.dyn1) { contact_force(F, lam, dt, 1, fl, fa); ldw(v, c.id1, scene, wl, wa); stw(v, c.id1, scene, wl + fl, wa + fa); }
        });
        grid.sync();
    }
    if (v.profiling && gtid == 0) {
        unsigned long long cv = 0, jv = 0;
        for (int s = 0; s < v.B; ++s) { cv += (unsigned long long)v.ncontacts[s] * iters; jv += (unsigned long long)v.njoints[s] * iters; }
        atomicAdd(&v.visit_counters[0], cv);
        atomicAdd(&v.visit_counters[1], jv);
    }
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
void launch_contact_colors(const View& v, cudaStream_t s)
{
    k_contact_colors<<<(unsigned)v.B, COLOR_THREADS, 0, s>>>(v);
}

// returns cudaSuccess or the launch error (cooperative launch: every CTA must be resident at once)
int launch_pgs_color(const View& v, float dt, int iters, float mu, cudaStream_t s)
{
    static int max_ctas = 0;
    if (!max_ctas) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pgs_color, 256, 0);
        max_ctas = sms * (per_sm > 0 ? per_sm : 1);
    }
    // no more CTAs than the widest thing to deal out: the bodies (accumulator reset) or the contact capacity
    const size_t want = ((size_t)v.B * (size_t)(v.NB > v.NC ? v.NB : v.NC) + 255) / 256;
    unsigned blocks = (unsigned)(want < (size_t)max_ctas ? want : (size_t)max_ctas);
    if (blocks < 1) blocks = 1;
    View vv = v;
    void* args[] = { &vv, &dt, &iters, &mu };
    return (int)cudaLaunchCooperativeKernel((const void*)k_pgs_color, dim3(blocks), dim3(256), args, 0, s);
}

}  // namespace slrbs
