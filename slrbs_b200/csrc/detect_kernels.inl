// detect_kernels.inl -- the detection kernels, included twice by detect.cu: once with DETECT_EXT = false (the
// reference's pair dispatch, namespace dref) and once with DETECT_EXT = true (plus the extension pairs, namespace dext), so
// that the default kernels carry none of the extension code (registers, stack, instruction cache).

// Tests the pair (i, j) on every lane that has a partner (j >= 0) and appends the contacts in lane order.
// `count` is the warp-uniform running total.
__device__ __forceinline__ void emit_chunk(const View& v, const DetectSmem& sm, int scene, int lane, int i, const float4& ci,
                                           int fi, int j, int& count)
{
    PairOut o;
    test_pair_warp<DETECT_EXT>(o, v, scene, i, j, ci, fi, sm.sp[j >= 0 ? j : i], sm.sf[j >= 0 ? j : i], lane);
    const unsigned any = __ballot_sync(0xffffffffu, o.cnt > 0);
    if (any == 0) return;
    const unsigned multi = __ballot_sync(0xffffffffu, o.cnt > 1);
    const unsigned lt = (1u << lane) - 1u;
    int before, total;
    if (multi == 0) {                                            // common case: at most one contact per pair
        before = __popc(any & lt);
        total = __popc(any);
    } else {                                                     // up to 4 per pair: shuffle prefix sum
        int incl = o.cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        before = incl - o.cnt;
        total = __shfl_sync(0xffffffffu, incl, 31);
    }
    write_contacts(v, scene, count + before, o);
    count += total;
}

// Partners j > i of sphere i from the grid: the spheres of the 27 neighbouring cells plus the non-sphere bodies,
// unordered in sm.cand. Two neighbouring cells may hash to the same bucket, so a partner can appear twice; the
// duplicates are dropped once the touching pairs are known (grid_collect). Returns the number of candidates, or
// -1 when there are more than CAND_CAP (the caller then walks all j > i). Warp-collective.
__device__ __forceinline__ int gather_candidates(const DetectSmem& sm, int lane, int i, const float4& ci, float inv_h, int H,
                                                 int nsph, int nspecial)
{
    const int Hmask = H - 1;
    if (lane == 0) *sm.counter = 0;
    __syncwarp();
    if (lane < 27) {
        const int ix = (int)floorf(ci.x * inv_h), iy = (int)floorf(ci.y * inv_h), iz = (int)floorf(ci.z * inv_h);
        const int h = cell_hash(ix + lane % 3 - 1, iy + (lane / 3) % 3 - 1, iz + lane / 9 - 1, Hmask);
        const int beg = sm.cell_start[h + 1];
        const int end = (h + 2 <= H) ? sm.cell_start[h + 2] : nsph;
        for (int k = beg; k < end; ++k) {
            const int j = sm.cell_items[k];
            if (j > i) { const int slot = atomicAdd(sm.counter, 1); if (slot < CAND_CAP) sm.cand[slot] = j; }
        }
    }
    for (int k = lane; k < nspecial; k += 32) {
        const int j = sm.special[k];
        if (j > i) { const int slot = atomicAdd(sm.counter, 1); if (slot < CAND_CAP) sm.cand[slot] = j; }
    }
    __syncwarp();
    const int total = *sm.counter;
    return total > CAND_CAP ? -1 : total;
}

// One grid row, step 1: tests the unordered candidates and buffers the pairs that touch in sm.hits (duplicates of a
// partner dropped). Returns the row's contact count (nh = number of buffered pairs), or -1 if more than HIT_CAP
// pairs touch (the caller then redoes the row with the ordered all-pairs walk). Warp-collective.
__device__ __forceinline__ int grid_collect(const View& v, const DetectSmem& sm, int scene, int lane, int i, const float4& ci, int fi,
                                            int ncand, int& nh)
{
    const unsigned lt = (1u << lane) - 1u;
    nh = 0;
    for (int k0 = 0; k0 < ncand; k0 += 32) {
        const int k = k0 + lane;
        const int j = k < ncand ? sm.cand[k] : -1;
        PairOut o;
        o.cnt = 0;
        if (j >= 0) test_pair<DETECT_EXT>(o, v, scene, i, j, ci, fi, sm.sp[j], sm.sf[j]);
        const unsigned any = __ballot_sync(0xffffffffu, o.cnt > 0);
        if (any == 0) continue;
        const int slot = nh + __popc(any & lt);
        if (o.cnt > 0 && slot < HIT_CAP) {
            int* hrec = sm.hits + slot * HIT_WORDS;
            float* f = reinterpret_cast<float*>(hrec);
            hrec[0] = j; hrec[1] = o.b0; hrec[2] = o.b1; hrec[3] = o.cnt;
            f[4] = o.n.x; f[5] = o.n.y; f[6] = o.n.z;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q < o.cnt) { f[7 + 3 * q] = o.p[q].x; f[8 + 3 * q] = o.p[q].y; f[9 + 3 * q] = o.p[q].z; f[19 + q] = o.phi[q]; }
            }
        }
        nh += __popc(any);
    }
    __syncwarp();
    if (nh > HIT_CAP) return -1;
    if (nh == 0) return 0;
    // drop repeated partners (keep the first), count the contacts
    int c = 0;
    if (lane < nh) {
        const int myj = sm.hits[lane * HIT_WORDS];
        bool dup = false;
        for (int m = 0; m < lane; ++m) dup |= sm.hits[m * HIT_WORDS] == myj;
        c = dup ? 0 : sm.hits[lane * HIT_WORDS + 3];
    }
    __syncwarp();
    if (lane < nh) sm.hits[lane * HIT_WORDS + 3] = c;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    __syncwarp();
    return c;
}

// One grid row, step 2: orders the buffered pairs by partner index (distinct) and appends their contacts at `base`.
__device__ __forceinline__ void grid_flush(const View& v, const DetectSmem& sm, int scene, int lane, int nh, int base)
{
    if (nh == 0) return;
    if (lane < nh) {
        const int myj = sm.hits[lane * HIT_WORDS];
        int rank = 0;
        for (int m = 0; m < nh; ++m) { const int oj = sm.hits[m * HIT_WORDS]; rank += (oj < myj) || (oj == myj && m < lane); }
        sm.order[rank] = lane;
    }
    __syncwarp();
    PairOut o;
    o.cnt = 0;
    if (lane < nh) {
        const int* hrec = sm.hits + sm.order[lane] * HIT_WORDS;
        const float* f = reinterpret_cast<const float*>(hrec);
        o.b0 = hrec[1]; o.b1 = hrec[2]; o.cnt = hrec[3];
        o.n = mk3(f[4], f[5], f[6]);
#pragma unroll
        for (int q = 0; q < 4; ++q) { o.p[q] = mk3(f[7 + 3 * q], f[8 + 3 * q], f[9 + 3 * q]); o.phi[q] = f[19 + q]; }
    }
    int incl = o.cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    write_contacts(v, scene, base + incl - o.cnt, o);
    __syncwarp();
}

template <bool GRID>
__global__ void __launch_bounds__(256) k_detect_warp(View v, int warps_per_cta, int H)
{
    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int scene = blockIdx.x * warps_per_cta + warp;
    if (scene >= v.B) return;
    DetectSmem sm;
    sm.sp = smem4 + (size_t)warp * v.NB;
    int* ip = reinterpret_cast<int*>(smem4 + (size_t)warps_per_cta * v.NB) + (size_t)warp * detect_smem_ints(v.NB, GRID, H);
    sm.sf = ip; ip += v.NB;
    sm.cell_start = ip; sm.cell_items = ip + H + 1; sm.special = sm.cell_items + v.NB;
    sm.counter = sm.special + v.NB; sm.cand = sm.counter + 2; sm.order = sm.cand + CAND_CAP; sm.hits = sm.order + HIT_CAP;

    const int n = v.nbodies[scene];
    float rmax = 0.f;
    for (int b = lane; b < n; b += 32) {
        const size_t i = at(v, b, scene);
        const float4 c = v.cproxy[i];
        const int f = v.flags[i];
        sm.sp[b] = c;
        sm.sf[b] = f;
        if ((f & FLAG_TYPE_MASK) == GEOM_SPHERE) rmax = fmaxf(rmax, c.w);
    }
    __syncwarp();

    // ---- broad phase: hashed uniform grid over the spheres, side list of the other shapes ----
    int nspecial = 0;
    float inv_h = 0.f;
    const int Hmask = H - 1;
    if (GRID) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, d));
        inv_h = 1.0f / (2.0f * rmax + 2e-3f);
        for (int c = lane; c <= H; c += 32) sm.cell_start[c] = 0;
        __syncwarp();
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int b = b0 + lane;
            bool spec = false;
            if (b < n) {
                if ((sm.sf[b] & FLAG_TYPE_MASK) == GEOM_SPHERE) {
                    const float4 c = sm.sp[b];
                    const int h = cell_hash((int)floorf(c.x * inv_h), (int)floorf(c.y * inv_h), (int)floorf(c.z * inv_h), Hmask);
                    atomicAdd(&sm.cell_start[h + 1], 1);
                } else {
                    spec = true;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, spec);
            if (spec) sm.special[nspecial + __popc(m & ((1u << lane) - 1u))] = b;
            nspecial += __popc(m);
        }
        __syncwarp();
        // inclusive scan of the counts: cell_start[c] = first slot of bucket c
        int carry = 0;
        for (int c0 = 1; c0 <= H; c0 += 32) {
            const int c = c0 + lane;
            const int cnt = c <= H ? sm.cell_start[c] : 0;
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (c <= H) sm.cell_start[c] = carry + incl;         // = end of bucket c-1 = start of bucket c
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        // fill: bucket c occupies [cell_start[c], cell_start[c+1]); use cell_start[c+1]-- as a cursor from the end
        for (int b = lane; b < n; b += 32) {
            if ((sm.sf[b] & FLAG_TYPE_MASK) == GEOM_SPHERE) {
                const float4 c = sm.sp[b];
                const int h = cell_hash((int)floorf(c.x * inv_h), (int)floorf(c.y * inv_h), (int)floorf(c.z * inv_h), Hmask);
                const int slot = atomicSub(&sm.cell_start[h + 1], 1) - 1;
                sm.cell_items[slot] = b;
            }
        }
        __syncwarp();
        // after the fill cell_start[c+1] == old cell_start[c]: bucket c is now [cell_start[c+1], cell_start[c+2]) for
        // c < H-1 and the last bucket ends at the number of spheres; keep that convention below (bucket_range)
    }
    int nsph = 0;
    if (GRID) nsph = n - nspecial;

    int count = 0;
    for (int i = 0; i < n; ++i) {
        const float4 ci = sm.sp[i];
        const int fi = sm.sf[i];
        bool dense_row = true;
        if (GRID && (fi & FLAG_TYPE_MASK) == GEOM_SPHERE) {
            const int ncand = gather_candidates(sm, lane, i, ci, inv_h, H, nsph, nspecial);
            if (ncand >= 0) {
                int nh;
                const int c = grid_collect(v, sm, scene, lane, i, ci, fi, ncand, nh);
                if (c >= 0) { dense_row = false; grid_flush(v, sm, scene, lane, nh, count); count += c; }
            }
        }
        if (dense_row) {
            for (int j0 = i + 1; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                emit_chunk(v, sm, scene, lane, i, ci, fi, j < n ? j : -1, count);
            }
        }
    }
    if (lane == 0) v.ncontacts[scene] = count < v.NC ? count : v.NC;
}

// ------------------------------------------------------------------------------------
// k_detect_cta<GRID>  one CTA per scene, for scenes of hundreds or thousands of bodies. Rows i are dealt
// to the warps in rounds of consecutive rows and tested in parallel; a block barrier per round turns the
// rows' contact counts into output offsets, so the list stays in (i, j) order with a single pass over the pairs.
// ------------------------------------------------------------------------------------
// all-pairs walk of row i (pairs (i, j), j > i ascending): counts the contacts, and writes them at `base` when WRITE
template <bool WRITE>
__device__ __forceinline__ int dense_row(const View& v, const DetectSmem& sm, int scene, int lane, int i, int n, int base)
{
    const float4 ci = sm.sp[i];
    const int fi = sm.sf[i];
    int off = 0;
    for (int j0 = i + 1; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        PairOut o;
        test_pair_warp<DETECT_EXT>(o, v, scene, i, j < n ? j : -1, ci, fi, sm.sp[j < n ? j : i], sm.sf[j < n ? j : i], lane);
        const unsigned any = __ballot_sync(0xffffffffu, o.cnt > 0);
        if (any == 0) continue;
        int incl = o.cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (WRITE) write_contacts(v, scene, base + off + incl - o.cnt, o);
        off += __shfl_sync(0xffffffffu, incl, 31);
    }
    return off;
}

template <bool GRID>
__global__ void __launch_bounds__(DETECT_CTA_THREADS, 2) k_detect_cta(View v, int H)
{
    extern __shared__ float4 smem4[];
    __shared__ int s_rmax, s_nspecial, s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = DETECT_CTA_THREADS / 32;
    const int scene = blockIdx.x;
    DetectSmem sm;
    sm.sp = smem4;
    int* ip = reinterpret_cast<int*>(smem4 + v.NB);
    sm.sf = ip; ip += v.NB;
    int* rowcnt = ip; ip += v.NB + 2;
    sm.cell_start = ip; sm.cell_items = ip + H + 1; sm.special = sm.cell_items + v.NB;
    sm.counter = sm.special + v.NB + (size_t)warp * (2 + CAND_CAP + HIT_CAP + HIT_CAP * HIT_WORDS);
    sm.cand = sm.counter + 2; sm.order = sm.cand + CAND_CAP; sm.hits = sm.order + HIT_CAP;

    const int n = v.nbodies[scene];
    if (tid == 0) { s_rmax = 0; s_nspecial = 0; s_total = 0; }
    __syncthreads();
    float rmax = 0.f;
    for (int b = tid; b < n; b += DETECT_CTA_THREADS) {
        const size_t i = at(v, b, scene);
        const float4 c = v.cproxy[i];
        const int f = v.flags[i];
        sm.sp[b] = c; sm.sf[b] = f;
        if ((f & FLAG_TYPE_MASK) == GEOM_SPHERE) rmax = fmaxf(rmax, c.w);
    }
    if (GRID) {
        for (int c = tid; c <= H; c += DETECT_CTA_THREADS) sm.cell_start[c] = 0;
        atomicMax(&s_rmax, __float_as_int(rmax));              // radii are >= 0: integer order = float order
    }
    __syncthreads();
    float inv_h = 0.f;
    int nspecial = 0, nsph = 0;
    const int Hmask = H - 1;
    if (GRID) {
        inv_h = 1.0f / (2.0f * __int_as_float(s_rmax) + 2e-3f);
        for (int b = tid; b < n; b += DETECT_CTA_THREADS) {
            if ((sm.sf[b] & FLAG_TYPE_MASK) == GEOM_SPHERE) {
                const float4 c = sm.sp[b];
                atomicAdd(&sm.cell_start[cell_hash((int)floorf(c.x * inv_h), (int)floorf(c.y * inv_h), (int)floorf(c.z * inv_h), Hmask) + 1], 1);
            }
        }
        if (warp == 0) {                                         // non-sphere bodies, ascending
            int ns = 0;
            for (int b0 = 0; b0 < n; b0 += 32) {
                const int b = b0 + lane;
                const bool spec = b < n && (sm.sf[b] & FLAG_TYPE_MASK) != GEOM_SPHERE;
                const unsigned m = __ballot_sync(0xffffffffu, spec);
                if (spec) sm.special[ns + __popc(m & ((1u << lane) - 1u))] = b;
                ns += __popc(m);
            }
            if (lane == 0) s_nspecial = ns;
        }
        __syncthreads();
        if (warp == 0) {
            int carry = 0;
            for (int c0 = 1; c0 <= H; c0 += 32) {
                const int c = c0 + lane;
                const int cnt = c <= H ? sm.cell_start[c] : 0;
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (c <= H) sm.cell_start[c] = carry + incl;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
        for (int b = tid; b < n; b += DETECT_CTA_THREADS) {
            if ((sm.sf[b] & FLAG_TYPE_MASK) == GEOM_SPHERE) {
                const float4 c = sm.sp[b];
                const int h = cell_hash((int)floorf(c.x * inv_h), (int)floorf(c.y * inv_h), (int)floorf(c.z * inv_h), Hmask);
                sm.cell_items[atomicSub(&sm.cell_start[h + 1], 1) - 1] = b;
            }
        }
        __syncthreads();
        nspecial = s_nspecial; nsph = n - nspecial;
    }
    // Rounds of W consecutive rows, one per warp: every warp tests its row and buffers the pairs that touch, a block
    // barrier publishes the W counts, each warp derives its output offset and appends. The contact list comes out in
    // (i, j) order from a single pass over the pairs (rows whose hits do not fit the buffer count first and re-test
    // when they write).
    __shared__ int s_cnt[DETECT_CTA_THREADS / 32];
    int running = 0;
    for (int i0 = 0; i0 < n; i0 += W) {
        const int i = i0 + warp;
        int count = 0, nh = 0;
        bool buffered = true;
        if (i < n) {
            const float4 ci = sm.sp[i];
            const int fi = sm.sf[i];
            count = -1;
            if (GRID && (fi & FLAG_TYPE_MASK) == GEOM_SPHERE) {
                const int ncand = gather_candidates(sm, lane, i, ci, inv_h, H, nsph, nspecial);
                if (ncand >= 0) count = grid_collect(v, sm, scene, lane, i, ci, fi, ncand, nh);
            }
            buffered = count >= 0;
            if (!buffered) count = dense_row<false>(v, sm, scene, lane, i, n, 0);
        }
        if (lane == 0) s_cnt[warp] = count;
        __syncthreads();
        int base = running, tot = 0;
#pragma unroll
        for (int w2 = 0; w2 < DETECT_CTA_THREADS / 32; ++w2) {
            const int c = s_cnt[w2];
            if (w2 < warp) base += c;
            tot += c;
        }
        if (i < n) {
            if (buffered) grid_flush(v, sm, scene, lane, nh, base);
            else dense_row<true>(v, sm, scene, lane, i, n, base);
        }
        running += tot;
        __syncthreads();
    }
    if (tid == 0) s_total = running;
    __syncthreads();
    if (tid == 0) v.ncontacts[scene] = s_total < v.NC ? s_total : v.NC;
}

// ------------------------------------------------------------------------------------
// k_detect_rows  one CTA per scene, one THREAD per sphere row. After the grid is built in shared memory, thread i
// walks the cell of sphere i and the 13 cells after it (half stencil) and records each touching pair once, in the
// row of the lower index (short lists of 16-bit indices in shared memory, filled through per-row atomics); the row
// owners then sort their lists and add the non-sphere bodies they touch. Rows of non-sphere bodies, and sphere rows with more than ROW_HITS partners, are
// walked by a warp each in the reference's all-pairs order (dense_row). A block scan turns the counts into output
// offsets, then every thread re-evaluates its few touching pairs and writes them: the list is in the reference's
// lexicographic (i, j) order with no sorting of candidates, no per-row barriers and all 32 lanes testing pairs.
// Every sphere-row partner yields exactly one contact (sphere-sphere, sphere-box; CollisionDetect.cpp:102-150).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 2) k_detect_rows(View v, int H)
{
    extern __shared__ float4 smem4[];
    __shared__ int s_rmax, s_nspecial, s_ndense, s_warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, W = T >> 5;
    const int scene = blockIdx.x;
    DetectSmem sm;
    sm.sp = smem4;
    int* ip = reinterpret_cast<int*>(smem4 + v.NB);
    sm.sf = ip; ip += v.NB;
    int* rowoff = ip; ip += v.NB + 1;
    sm.cell_start = ip; ip += H + 2;
    sm.cell_items = ip; ip += v.NB;
    sm.special = ip; ip += v.NB;
    int* dense_list = ip; ip += v.NB;
    unsigned short* hits = reinterpret_cast<unsigned short*>(ip);
    unsigned char* nhits = reinterpret_cast<unsigned char*>(hits + (size_t)v.NB * ROW_HITS);

    const int n = v.nbodies[scene];
    if (tid == 0) { s_rmax = 0; s_nspecial = 0; s_ndense = 0; }
    for (int c = tid; c <= H + 1; c += T) sm.cell_start[c] = 0;
    __syncthreads();
    float rmax = 0.f;
    for (int b = tid; b < n; b += T) {
        const size_t i = at(v, b, scene);
        const float4 c = v.cproxy[i];
        const int f = v.flags[i];
        sm.sp[b] = c; sm.sf[b] = f;
        if ((f & FLAG_TYPE_MASK) == GEOM_SPHERE) rmax = fmaxf(rmax, c.w);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, d));
    if (lane == 0) atomicMax(&s_rmax, __float_as_int(rmax));    // radii are >= 0: integer order = float order
    __syncthreads();
    const float inv_h = 1.0f / (2.0f * __int_as_float(s_rmax) + 2e-3f);
    const int Hmask = H - 1;
    // bucket counts; non-sphere bodies listed in ascending order by warp 0
    for (int b = tid; b < n; b += T) {
        if ((sm.sf[b] & FLAG_TYPE_MASK) == GEOM_SPHERE) {
            const float4 c = sm.sp[b];
            atomicAdd(&sm.cell_start[cell_hash((int)floorf(c.x * inv_h), (int)floorf(c.y * inv_h), (int)floorf(c.z * inv_h), Hmask) + 1], 1);
        }
    }
    if (warp == 0) {
        int ns = 0;
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int b = b0 + lane;
            const bool spec = b < n && (sm.sf[b] & FLAG_TYPE_MASK) != GEOM_SPHERE;
            const unsigned m = __ballot_sync(0xffffffffu, spec);
            if (spec) { const int k = ns + __popc(m & ((1u << lane) - 1u)); sm.special[k] = b; dense_list[k] = b; }
            ns += __popc(m);
        }
        if (lane == 0) { s_nspecial = ns; s_ndense = ns; }
    }
    __syncthreads();
    if (warp == 0) {                                             // inclusive scan: cell_start[c] = end of bucket c-1
        int carry = 0;
        for (int c0 = 1; c0 <= H; c0 += 32) {
            const int c = c0 + lane;
            int incl = c <= H ? sm.cell_start[c] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (c <= H) sm.cell_start[c] = carry + incl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) sm.cell_start[H + 1] = carry;
    }
    __syncthreads();
    // fill backwards: afterwards cell_start[h+1] = begin of bucket h, cell_start[h+2] = its end (bucket H-1 ends at nsph)
    for (int b = tid; b < n; b += T) {
        if ((sm.sf[b] & FLAG_TYPE_MASK) == GEOM_SPHERE) {
            const float4 c = sm.sp[b];
            const int h = cell_hash((int)floorf(c.x * inv_h), (int)floorf(c.y * inv_h), (int)floorf(c.z * inv_h), Hmask);
            sm.cell_items[atomicSub(&sm.cell_start[h + 1], 1) - 1] = b;
        }
    }
    __syncthreads();
    const int nspecial = s_nspecial;

    // ---- pass 1a: grid walk, one thread per sphere, HALF stencil ------------------------------------------------------
    // Thread i walks its own cell and the 13 neighbouring cells that follow it in (z, y, x) order; a touching pair is
    // recorded once, by the sphere whose cell comes first (same cell: by the lower index), in the row of the LOWER index --
    // which may be another thread's row: the rows' counters are shared-memory atomics (rowoff doubles as the counter array
    // until the block scan) and the lists are sorted afterwards by their owners. A bucket is hashed, so it also holds
    // spheres of other cells: a candidate counts only if it really lives in the cell being visited (its pair with i is
    // found from the cell it does live in), which also makes every pair unique without de-duplication.
    for (int i = tid; i < n; i += T) rowoff[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += T) {
        const int fi = sm.sf[i];
        if ((fi & FLAG_TYPE_MASK) != GEOM_SPHERE) continue;
        const float4 ci = sm.sp[i];
        const int ix = (int)floorf(ci.x * inv_h), iy = (int)floorf(ci.y * inv_h), iz = (int)floorf(ci.z * inv_h);
#pragma unroll 1
        for (int cc = 13; cc < 27; ++cc) {                       // 13 = the sphere's own cell, 14 .. 26 = the cells after it
            const int cx = ix + cc % 3 - 1, cy = iy + (cc / 3) % 3 - 1, cz = iz + cc / 9 - 1;
            const int h = cell_hash(cx, cy, cz, Hmask);
            const int beg = sm.cell_start[h + 1], end = sm.cell_start[h + 2];
            for (int k = beg; k < end; ++k) {
                const int j = sm.cell_items[k];
                if (j == i || (cc == 13 && j < i)) continue;
                // the grid holds spheres only: the touch test of collisionDetectSphereSphere (CollisionDetect.cpp:111-115,
                // same expressions as sphere_sphere(), so the same decision to the bit -- from either side of the pair)
                // without building the contact
                const float4 cj = sm.sp[j];
                if ((fi & sm.sf[j]) & FLAG_FIXED) continue;
                const V3 vec = mk3(ci) - mk3(cj);
                const float sq = sqnorm(vec), lim = (ci.w + cj.w) + 1e-3f;
                if (sq > 1.0001f * (lim * lim)) continue;        // cannot pass the exact test below
                if (!(sqrtf(sq) < lim)) continue;
                if ((int)floorf(cj.x * inv_h) != cx || (int)floorf(cj.y * inv_h) != cy || (int)floorf(cj.z * inv_h) != cz) continue;
                const int r = min(i, j), p = max(i, j);
                const int slot = atomicAdd(&rowoff[r], 1);
                if (slot < ROW_HITS) hits[(size_t)r * ROW_HITS + slot] = (unsigned short)p;
            }
        }
    }
    __syncthreads();
    // ---- pass 1b: every sphere row sorts its partners and adds the non-sphere bodies -----------------------------------
    for (int i = tid; i < n; i += T) {
        const int fi = sm.sf[i];
        if ((fi & FLAG_TYPE_MASK) != GEOM_SPHERE) { nhits[i] = ROW_DENSE; continue; }
        const float4 ci = sm.sp[i];
        unsigned short* my = hits + (size_t)i * ROW_HITS;
        int nh = rowoff[i];
        bool over = nh > ROW_HITS;
        if (!over) {
            for (int a = 1; a < nh; ++a) {                       // insertion sort, ascending partner index
                const unsigned short x = my[a];
                int b = a;
                while (b > 0 && my[b - 1] > x) { my[b] = my[b - 1]; --b; }
                my[b] = x;
            }
        }
        // The non-sphere bodies, 32 at a time: the cheap cull for all of them first (every lane busy), then each lane works
        // through its own survivors -- a warp of spheres along two walls runs the box test twice with a dozen lanes in it
        // instead of once per wall with three or four
        for (int k0 = 0; k0 < nspecial && !over; k0 += 32) {
            unsigned pend = 0;
            const int kn = min(32, nspecial - k0);
            for (int k = 0; k < kn; ++k) {
                const int j = sm.special[k0 + k];
                const int fj = sm.sf[j];
                bool may = j > i && !((fi & fj) & FLAG_FIXED);
                if (may && (fj & FLAG_TYPE_MASK) == GEOM_BOX) may = sphere_box_may_touch(ci, sm.sp[j], __ldg(&v.bext[at(v, j, scene)]));
                if (may) pend |= 1u << k;
            }
            while (pend) {
                const int j = sm.special[k0 + __ffs(pend) - 1];
                pend &= pend - 1;
                PairOut o;
                test_pair<DETECT_EXT>(o, v, scene, i, j, ci, fi, sm.sp[j], sm.sf[j]);
                if (o.cnt == 0) continue;
                if (nh == ROW_HITS) { over = true; break; }
                int pos = nh;
                while (pos > 0 && my[pos - 1] > j) --pos;
                for (int m = nh; m > pos; --m) my[m] = my[m - 1];
                my[pos] = (unsigned short)j;
                ++nh;
            }
        }
        if (over) { nhits[i] = ROW_DENSE; dense_list[atomicAdd(&s_ndense, 1)] = i; }
        else { nhits[i] = (unsigned char)nh; rowoff[i] = nh; }
    }
    __syncthreads();
    // ---- dense rows (non-sphere bodies, overfull sphere rows): one warp per row, all j > i ---------------------------
    const int ndense = s_ndense;
    for (int k = warp; k < ndense; k += W) {
        const int i = dense_list[k];
        const int c = dense_row<false>(v, sm, scene, lane, i, n, 0);
        if (lane == 0) rowoff[i] = c;
    }
    __syncthreads();
    // ---- exclusive block scan of the row counts ----------------------------------------------------------------------
    {
        const int per = (n + T - 1) / T;
        const int r0 = min(tid * per, n), r1 = min(r0 + per, n);
        int sum = 0;
        for (int r = r0; r < r1; ++r) sum += rowoff[r];
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        __syncthreads();
        int base = 0;
        for (int w2 = 0; w2 < warp; ++w2) base += s_warp_tot[w2];
        int run = base + incl - sum;
        for (int r = r0; r < r1; ++r) { const int c = rowoff[r]; rowoff[r] = run; run += c; }
        if (tid == T - 1) rowoff[n] = run;
        __syncthreads();
    }
    // ---- pass 2: write ---------------------------------------------------------------------------------------------------
    for (int i = tid; i < n; i += T) {
        const int nh = nhits[i];
        if (nh == ROW_DENSE) continue;
        const float4 ci = sm.sp[i];
        const int fi = sm.sf[i];
        const int base = rowoff[i];
        // back to front: the slot of a contact is base + k whatever the order, and a row's non-sphere partners (the walls of
        // the marble box: the highest indices, i.e. the END of the ascending list) then meet in the first trips across the
        // lanes instead of in whichever trip each lane's list happens to reach them
        for (int k = nh - 1; k >= 0; --k) {
            const int j = hits[(size_t)i * ROW_HITS + k];
            PairOut o;
            test_pair<DETECT_EXT>(o, v, scene, i, j, ci, fi, sm.sp[j], sm.sf[j]);
            write_contacts(v, scene, base + k, o);
        }
    }
    for (int k = warp; k < ndense; k += W) {
        const int i = dense_list[k];
        dense_row<true>(v, sm, scene, lane, i, n, rowoff[i]);
    }
    if (tid == 0) { const int tot = rowoff[n]; v.ncontacts[scene] = tot < v.NC ? tot : v.NC; }
}

// scenes of a few bodies: one thread per scene walks the pairs
__global__ void __launch_bounds__(128) k_detect_serial(View v)
{
    const int scene = blockIdx.x * blockDim.x + threadIdx.x;
    if (scene >= v.B) return;
    const int n = v.nbodies[scene];
    int count = 0;
    for (int i = 0; i < n; ++i) {
        const size_t ii = at(v, i, scene);
        const float4 ci = v.cproxy[ii];
        const int fi = v.flags[ii];
        for (int j = i + 1; j < n; ++j) {
            const size_t jj = at(v, j, scene);
            PairOut o;
            test_pair<DETECT_EXT>(o, v, scene, i, j, ci, fi, v.cproxy[jj], v.flags[jj]);
            write_contacts(v, scene, count, o);
            count += o.cnt;
        }
    }
    v.ncontacts[scene] = count < v.NC ? count : v.NC;
}
