// bpp.cu -- K6d  block principal pivoting (solverId 3) over the joints and contacts of a scene.
//
// EXTENSION: the reference has no such solver; the algorithm is defined by
// oracle/slrbs_oracle.c solve_bpp, which this kernel follows operation for operation (same row order,
// same summation order, no fused multiply-adds: the translation unit is compiled with -fmad=false), so
// on the same rows both produce the same impulses.
//
// Mapping: one CTA per scene. The scalar rows (joints in slot order, then contacts in slot order: normal,
// two friction rows) are decoded from the row records the assembly kernels wrote into shared memory
// ([field][row], 24 floats + two body ids + rhs per row); entries of A = J Minv J^T + eps I are formed on
// demand from them (two rows couple only through a shared non-fixed body, so most entries are four integer
// compares). Each pivot step: ordered compaction of the free set (one warp, ballots), gather of A_FF into a
// packed lower triangle in shared memory, left-looking Cholesky with the forward substitution riding along as an
// extra row (one thread per row, one barrier per column), back substitution, residual velocities of the bound
// rows (one thread per row, sequential over columns), set changes by the Judice-Pires rule. Scenes of at most 64
// rows run on a single warp (32-thread CTA, warp barriers) with the dense A kept in shared memory; a batch is
// served by one launch per size class, each skipping the other classes' scenes. Shared memory bounds the scene at
// BPP_MAX_ROWS scalar rows; a larger scene raises the context's capacity flag (bit 1) and keeps zero impulses.
#include "kernels.cuh"
#include "pgs_math.cuh"

namespace slrbs {

enum { BPP_FREE = 0, BPP_LOWER = 1, BPP_UPPER = 2 };
enum { BPP_BILATERAL = 0, BPP_NORMAL = 1, BPP_FRICTION = 2 };
enum { BPP_ROW_FLOATS = 24, BPP_ROW_WORDS = 36, BPP_WARP_ROWS = 64 };     // J0 J1 M0 M1 | b0 b1 kind parent rhs lam y hi col state F move
#define BPP_TOL 1e-6f

// shared-memory words of a scene of at most n rows. One-warp variant: the dense A ([n][n | 1]) is kept, and the
// Cholesky triangle (n + 1 rows: the right-hand side is its last row) reuses the row records' space once A is formed.
__host__ __device__ inline size_t bpp_tri_words(int n) { return (size_t)n * (n + 1) / 2; }
__host__ __device__ inline size_t bpp_rec_words(int n, bool one_warp)
{
    const size_t rec = (size_t)BPP_ROW_FLOATS * n;
    return one_warp && bpp_tri_words(n + 1) > rec ? bpp_tri_words(n + 1) : rec;
}
__host__ __device__ inline size_t bpp_smem_words(int n, bool one_warp)
{
    const size_t vec = (size_t)(BPP_ROW_WORDS - BPP_ROW_FLOATS) * n + 8;
    return bpp_rec_words(n, one_warp) + vec + (one_warp ? (size_t)n * (n | 1) : bpp_tri_words(n + 1));
}

struct BppRows {
    int n;                     // row capacity = stride of the [field][row] arrays
    float* Rf;                 // [24][n]: J0[6] J1[6] M0[6] M1[6]
    int *b0, *b1, *kind, *parent;
    float* rhs;
};

__device__ __forceinline__ float bpp_dot6(const BppRows& R, int fa, int i, int fb, int j)
{
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += R.Rf[(fa + k) * R.n + i] * R.Rf[(fb + k) * R.n + j];
    return acc;
}

// A(i, j), the oracle's bpp_entry
__device__ __forceinline__ float bpp_entry(const BppRows& R, int i, int j)
{
    const int a0 = R.b0[i], a1 = R.b1[i], c0 = R.b0[j], c1 = R.b1[j];
    float acc = 0.f;
    if (a0 >= 0) {
        if (a0 == c0) acc += bpp_dot6(R, 12, i, 0, j);
        if (a0 == c1) acc += bpp_dot6(R, 12, i, 6, j);
    }
    if (a1 >= 0) {
        if (a1 == c0) acc += bpp_dot6(R, 18, i, 0, j);
        if (a1 == c1) acc += bpp_dot6(R, 18, i, 6, j);
    }
    if (i == j) acc += R.kind[i] == BPP_BILATERAL ? 1e-5f : 1e-3f;
    return acc;
}

__device__ __forceinline__ void bpp_put_row(const BppRows& R, int k, const float* J0, const float* J1, const float* M0, const float* M1)
{
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        R.Rf[c * R.n + k] = J0[c]; R.Rf[(6 + c) * R.n + k] = J1[c];
        R.Rf[(12 + c) * R.n + k] = M0[c]; R.Rf[(18 + c) * R.n + k] = M1[c];
    }
}

__device__ __forceinline__ size_t tri(int i, int j) { return (size_t)i * (i + 1) / 2 + j; }

// ONE_WARP: a scene of at most 64 rows is solved by a single warp (32-thread CTA), every barrier a __syncwarp
template <bool ONE_WARP>
__device__ __forceinline__ void bpp_sync() { if (ONE_WARP) __syncwarp(); else __syncthreads(); }

// BIG: scenes beyond BPP_MAX_ROWS rows (the reference's own marble box has 1107): the row records and the packed Cholesky
// triangle of the free set live in a per-scene global scratch area (L2 resident: 2.5 MB for 1107 rows) instead of shared
// memory, which then only holds the per-row vectors; same code, same operation order, one 512-thread CTA per scene.
template <bool ONE_WARP, bool BIG = false>
__global__ void __launch_bounds__(BIG ? 512 : 256) k_bpp(View v, int maxit, float mu, int cap, int min_rows, int flag_overflow, float* scratch = nullptr, size_t scratch_stride = 0)
{
    extern __shared__ float sm[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
    const int scene = blockIdx.x;
    const int nj = v.njoints[scene], nc = v.ncontacts[scene];
    BppRows R;
    R.n = cap;
    float* const gs = BIG ? scratch + (size_t)blockIdx.x * scratch_stride : nullptr;
    R.Rf = BIG ? gs : sm;
    R.b0 = (int*)(sm + (BIG ? 0 : bpp_rec_words(cap, ONE_WARP))); R.b1 = R.b0 + cap; R.kind = R.b1 + cap; R.parent = R.kind + cap;
    R.rhs = (float*)(R.parent + cap);
    float* lam = R.rhs + cap; float* y = lam + cap; float* hi = y + cap; float* col = hi + cap;
    int* state = (int*)(col + cap); int* F = state + cap; int* move = F + cap;
    int* ctl = move + cap;                        // [0] ninf [1] last [2] shift bits [3] nf
    float* M = BIG ? gs + (size_t)BPP_ROW_FLOATS * cap : (ONE_WARP ? sm : (float*)(ctl + 8));
    const float* Ad = (const float*)(ctl + 8);    // one-warp variant: dense A, row stride ld
    const int ld = cap | 1;
    auto entry = [&](int i, int j) { return ONE_WARP ? Ad[i * ld + j] : bpp_entry(R, i, j); };

    // a batch is served by one launch per size class (launch_bpp); this CTA skips a scene that belongs to another
    // class. All exits below are uniform per CTA.
    if (3 * nj + 3 * nc > cap) {                   // every joint has at least three rows
        if (flag_overflow && tid == 0) atomicOr(v.overflow, 2);
        return;
    }
    // joint row offsets: dims into move[], serial prefix by one thread (a scene has a handful of joints)
    for (int s = tid; s < nj; s += T) move[s] = __float_as_int(__ldg(&v.jrow[at_jrow(v, 16, joint_pos(v, s, scene), scene)]).z);
    bpp_sync<ONE_WARP>();
    if (tid == 0) {
        int acc = 0;
        for (int s = 0; s < nj; ++s) { const int d = move[s]; F[s] = acc; acc += d; }
        ctl[3] = acc;
    }
    bpp_sync<ONE_WARP>();
    const int njr = ctl[3];
    const int n = njr + 3 * nc;
    if (n > cap) {
        if (flag_overflow && tid == 0) atomicOr(v.overflow, 2);
        return;
    }
    if (n == 0 || n < min_rows || maxit <= 0) return;
    bpp_sync<ONE_WARP>();

    for (int s = tid; s < nj; s += T) {
        JointRec r;
        load_joint_rec(v, joint_pos(v, s, scene), scene, r);
        const int dim = joint_dim(r), k0 = F[s];
        float A0[5][6], A1[5][6];
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int c = 0; c < 6; ++c) { A0[a][c] = 0.f; A1[a][c] = 0.f; }
        A0[0][0] = A0[1][1] = A0[2][2] = 1.f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) A1[a][c] = (a == c) ? -1.f : -0.f;
        set_hat(A0, -r.h.rr0); set_hat(A1, r.h.rr1);
        set_row(A0[3], mk3(0.f, 0.f, 0.f), r.h.nxu); set_row(A0[4], mk3(0.f, 0.f, 0.f), r.h.nxv);
        set_row(A1[3], mk3(0.f, 0.f, 0.f), -r.h.nxu); set_row(A1[4], mk3(0.f, 0.f, 0.f), -r.h.nxv);
#pragma unroll
        for (int a = 0; a < 5; ++a) if (a < dim) {
            const float m0[6] = { r.h.im0 * A0[a][0], r.h.im0 * A0[a][1], r.h.im0 * A0[a][2], r.M0[a].x, r.M0[a].y, r.M0[a].z };
            const float m1[6] = { r.h.im1 * A1[a][0], r.h.im1 * A1[a][1], r.h.im1 * A1[a][2], r.M1[a].x, r.M1[a].y, r.M1[a].z };
            const int k = k0 + a;
            bpp_put_row(R, k, A0[a], A1[a], m0, m1);
            R.b0[k] = r.h.id0 >= 0 ? r.h.id0 : -1;
            R.b1[k] = r.h.id1 >= 0 ? r.h.id1 : -1;
            R.kind[k] = BPP_BILATERAL; R.parent[k] = k0;
            R.rhs[k] = r.M0[a].w;
        }
    }
    for (int s = tid; s < nc; s += T) {
        float4 Fr[CROW_FIELDS];
        const int p = contact_pos(v, s, scene);
#pragma unroll
        load_crow_fields(v, p, scene, Fr);
        const ContactIds c = contact_ids(Fr);
        const float im0 = Fr[0].w, im1 = Fr[1].w;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float4 d = Fr[a], a0 = Fr[3 + a], a1 = Fr[6 + a], m0 = Fr[9 + a], m1 = Fr[12 + a];
            const float J0[6] = { d.x, d.y, d.z, a0.x, a0.y, a0.z };
            const float J1[6] = { -d.x, -d.y, -d.z, a1.x, a1.y, a1.z };
            const float M0[6] = { im0 * d.x, im0 * d.y, im0 * d.z, m0.x, m0.y, m0.z };
            const float M1[6] = { im1 * -d.x, im1 * -d.y, im1 * -d.z, m1.x, m1.y, m1.z };
            const int k = njr + 3 * s + a;
            bpp_put_row(R, k, J0, J1, M0, M1);
            R.b0[k] = c.dyn0 ? c.id0 : -1;
            R.b1[k] = c.dyn1 ? c.id1 : -1;
            R.kind[k] = a == 0 ? BPP_NORMAL : BPP_FRICTION; R.parent[k] = njr + 3 * s;
            R.rhs[k] = Fr[3 + a].w;
        }
    }
    bpp_sync<ONE_WARP>();
    for (int i = tid; i < n; i += T) { lam[i] = 0.f; state[i] = R.kind[i] == BPP_BILATERAL ? BPP_FREE : BPP_LOWER; }
    if (ONE_WARP) {
        float* A = (float*)(ctl + 8);
        for (int e = tid; e < n * n; e += T) { const int i = e / n, j = e - i * n; A[i * ld + j] = bpp_entry(R, i, j); }
    }
    bpp_sync<ONE_WARP>();

    int best = n + 1, budget = 3, refresh = 1, steps = 0;
    for (int it = 0; it < maxit; ++it) {
        ++steps;
        if (refresh) {
            for (int i = tid; i < n; i += T) {
                float h = 0.f;
                if (R.kind[i] == BPP_FRICTION) {
                    const float ln = lam[R.parent[i]];
                    h = mu * (ln > 0.f ? ln : 0.f);
                    if (h <= 0.f) state[i] = BPP_LOWER;
                }
                hi[i] = h;
            }
            refresh = 0; best = n + 1; budget = 3;
            bpp_sync<ONE_WARP>();
        }
        // free list in ascending order (warp 0), bound values
        if (warp == 0) {
            int base = 0;
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                const bool fr = i < n && state[i] == BPP_FREE;
                const unsigned m = __ballot_sync(0xffffffffu, fr);
                if (fr) F[base + __popc(m & ((1u << lane) - 1u))] = i;
                base += __popc(m);
            }
            if (lane == 0) { ctl[0] = 0; ctl[1] = -1; ctl[2] = 0; ctl[3] = base; }
        }
        for (int i = tid; i < n; i += T)
            if (state[i] != BPP_FREE) lam[i] = R.kind[i] == BPP_FRICTION ? (state[i] == BPP_LOWER ? -hi[i] : hi[i]) : 0.f;
        bpp_sync<ONE_WARP>();
        const int nf = ctl[3];
        // A_FF (packed lower triangle) and the right-hand side b_F - A_FB lam_B
        for (int a = warp; a < nf; a += nw) {
            const int ia = F[a];
            for (int c = lane; c <= a; c += 32) M[tri(a, c)] = entry(ia, F[c]);
        }
        for (int a = tid; a < nf; a += T) {
            const int ia = F[a];
            float acc = R.rhs[ia];
            for (int j = 0; j < n; ++j) if (state[j] != BPP_FREE && lam[j] != 0.f) acc -= entry(ia, j) * lam[j];
            M[tri(nf, a)] = acc;                   // the right-hand side rides along as row nf of the triangle
        }
        // Cholesky and L z = y in one left-looking pass, one thread per row: element (i, c) takes its updates in
        // the order k = 0 .. c-1 exactly as the right-looking loop of the restatement applies them, and every
        // thread recomputes the pivot of its column from the same numbers (no second barrier per column).
        // Row c's diagonal slot keeps A(c, c); the pivots live in col[].
        for (int c = 0; c < nf; ++c) {
            bpp_sync<ONE_WARP>();
            const size_t rc = tri(c, 0);
            for (int i = c + tid; i <= nf; i += T) {
                const size_t ri = tri(i, 0);
                float acc = M[ri + c], dc = M[rc + c];
                for (int j = 0; j < c; ++j) { const float lc = M[rc + j]; acc -= M[ri + j] * lc; dc -= lc * lc; }
                dc = sqrtf(dc > 1e-12f ? dc : 1e-12f);
                if (i == c) col[c] = dc; else M[ri + c] = acc / dc;
            }
        }
        bpp_sync<ONE_WARP>();
        // L^T x = z: z from row nf into y, x back into row nf
        for (int a = tid; a < nf; a += T) y[a] = M[tri(nf, a)];
        for (int k = nf - 1; k >= 0; --k) {
            bpp_sync<ONE_WARP>();
            const float xk = y[k] / col[k];
            for (int i = tid; i < k; i += T) y[i] -= M[tri(k, i)] * xk;
            if (tid == 0) M[tri(nf, k)] = xk;
        }
        bpp_sync<ONE_WARP>();
        for (int a = tid; a < nf; a += T) lam[F[a]] = M[tri(nf, a)];
        bpp_sync<ONE_WARP>();
        // residual velocities of the bound rows, rows that must change set, drift of the friction bounds
        for (int i = tid; i < n; i += T) {
            int mv = -1;
            const int kind = R.kind[i];
            if (kind == BPP_FRICTION) {
                const float ln = lam[R.parent[i]];
                const float nh = mu * (ln > 0.f ? ln : 0.f);
                const float ds = fabsf(nh - hi[i]) / (1.0f + nh);
                atomicMax(&ctl[2], __float_as_int(ds));
            }
            if (kind != BPP_BILATERAL && !(kind == BPP_FRICTION && hi[i] <= 0.f)) {
                const int st = state[i];
                if (st == BPP_FREE) {
                    if (kind == BPP_NORMAL) { if (lam[i] < -BPP_TOL) mv = BPP_LOWER; }
                    else if (lam[i] < -hi[i] - BPP_TOL) mv = BPP_LOWER;
                    else if (lam[i] > hi[i] + BPP_TOL) mv = BPP_UPPER;
                } else {
                    float acc = -R.rhs[i];
                    for (int j = 0; j < n; ++j) if (lam[j] != 0.f) acc += entry(i, j) * lam[j];
                    if (st == BPP_LOWER && acc < -BPP_TOL) mv = BPP_FREE;
                    if (st == BPP_UPPER && acc > BPP_TOL) mv = BPP_FREE;
                }
            }
            move[i] = mv;
            if (mv >= 0) { atomicAdd(&ctl[0], 1); atomicMax(&ctl[1], i); }
        }
        bpp_sync<ONE_WARP>();
        const int ninf = ctl[0], last = ctl[1];
        const float shift = __int_as_float(ctl[2]);
        if (ninf == 0) {
            if (shift <= 1e-5f) break;
            refresh = 1;
            bpp_sync<ONE_WARP>();                      // ctl is reset by warp 0 at the top of the next step
            continue;
        }
        bool block = true;
        if (ninf < best) { best = ninf; budget = 3; }
        else if (budget > 0) --budget;
        else block = false;
        if (block) { for (int i = tid; i < n; i += T) if (move[i] >= 0) state[i] = move[i]; }
        else if (tid == 0) state[last] = move[last];
        bpp_sync<ONE_WARP>();
    }
    bpp_sync<ONE_WARP>();
    if (v.profiling && tid == 0) atomicAdd(&v.visit_counters[0], (unsigned long long)steps);     // pivot steps
    // project onto the bounds (normal rows first: the friction bounds use the projected normals), store
    for (int i = tid; i < n; i += T) if (R.kind[i] == BPP_NORMAL && lam[i] < 0.f) lam[i] = 0.f;
    bpp_sync<ONE_WARP>();
    for (int i = tid; i < n; i += T) if (R.kind[i] == BPP_FRICTION) {
        const float ln = lam[R.parent[i]];
        const float h = mu * (ln > 0.f ? ln : 0.f);
        if (lam[i] < -h) lam[i] = -h; else if (lam[i] > h) lam[i] = h;
    }
    bpp_sync<ONE_WARP>();
    for (int s = tid; s < nc; s += T) {
        const int k = njr + 3 * s;
        v.crow[at_crow(v, CROW_LAMBDA, contact_pos(v, s, scene), scene)] = make_float4(lam[k], lam[k + 1], lam[k + 2], 0.f);
    }
    if (tid == 0) {
        int k = 0;
        for (int s = 0; s < nj; ++s) {
            const int dim = __float_as_int(__ldg(&v.jrow[at_jrow(v, 16, joint_pos(v, s, scene), scene)]).z);
            const size_t ji = at(v, s, scene);
            v.jlam[ji] = make_float4(lam[k], lam[k + 1], lam[k + 2], dim == 5 ? lam[k + 3] : 0.f);
            v.jlam4[ji] = dim == 5 ? lam[k + 4] : 0.f;
            k += dim;
        }
    }
}

int bpp_row_capacity(const View& v)
{
    const long want = 3L * v.NC + 5L * v.NJ;
    return (int)(want < BPP_MAX_ROWS ? (want > 0 ? want : 1) : BPP_MAX_ROWS);
}

// rows the global-scratch variant holds for this context (0: not needed or too much scratch), floats of scratch per scene
int bpp_big_capacity(const View& v)
{
    const long want = 3L * v.NC + 5L * v.NJ;
    if (want <= BPP_MAX_ROWS) return 0;
    const long cap = want < BPP_BIG_MAX_ROWS ? want : BPP_BIG_MAX_ROWS;
    return (double)bpp_big_scratch_floats((int)cap) * 4.0 * v.B <= 8e9 ? (int)cap : 0;
}
size_t bpp_big_scratch_floats(int cap) { return (size_t)BPP_ROW_FLOATS * cap + bpp_tri_words(cap + 1); }
static size_t bpp_big_smem_words(int cap) { return (size_t)(BPP_ROW_WORDS - BPP_ROW_FLOATS) * cap + 8; }

void init_bpp_kernel()
{
    cudaFuncSetAttribute(k_bpp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bpp_smem_words(BPP_MAX_ROWS, false) * sizeof(float)));
    cudaFuncSetAttribute(k_bpp<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_bpp<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_bpp<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bpp_big_smem_words(BPP_BIG_MAX_ROWS) * sizeof(float)));
}

// One launch per size class; every launch covers all scenes and a scene is solved by the smallest class that holds
// its rows (a CTA of another class exits at once), so shared memory per CTA -- and with it the number of scenes
// resident on an SM -- follows the scene's size rather than the context's capacity.
int launch_bpp(const View& v, int iters, float mu, float* big_scratch, cudaStream_t s)
{
    static const int classes[] = { 32, 48, BPP_WARP_ROWS, 96, 144, 208, BPP_MAX_ROWS };
    const int cap = bpp_row_capacity(v);
    const int big = big_scratch ? bpp_big_capacity(v) : 0;
    int prev = 0, launches = 0;
    for (int cls : classes) {
        const int c = cls < cap ? cls : cap;
        const int min_rows = prev ? prev + 1 : 0, last = c == cap && !big;        // the last class flags scenes nobody holds
        if (c <= BPP_WARP_ROWS)
            k_bpp<true><<<(unsigned)v.B, 32, bpp_smem_words(c, true) * sizeof(float), s>>>(v, 5 * iters, mu, c, min_rows, last);
        else
            k_bpp<false><<<(unsigned)v.B, c <= 96 ? 64 : (c <= 144 ? 128 : 256), bpp_smem_words(c, false) * sizeof(float), s>>>(
                v, 5 * iters, mu, c, min_rows, last);
        ++launches;
        prev = c;
        if (c == cap) break;
    }
    if (big) {
        k_bpp<false, true><<<(unsigned)v.B, 512, bpp_big_smem_words(big) * sizeof(float), s>>>(v, 5 * iters, mu, big, BPP_MAX_ROWS + 1, 1, big_scratch,
                                                                                           bpp_big_scratch_floats(big));
        ++launches;
    }
    return launches;
}

}  // namespace slrbs
