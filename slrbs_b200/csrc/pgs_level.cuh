// pgs_level.cuh -- pieces shared by the level-scheduled solver kernels (pgs_level.cu: one lane per row; pgs_split.cu:
// two lanes per contact row, one per body side).
#pragma once
#include "pgs_math.cuh"

namespace slrbs {

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(p), "r"(bytes) : "memory");
}
// Pull positions [q0, q1) of fields 0 .. nf-1 and of the impulse field of one scene's record stream (`rows`, [field][pos],
// uNC positions per field) towards L2, the (field, 128 B line) pairs dealt over `width` lanes, one prefetch.global.L2 each.
// (Round 2 used one cp.async.bulk.prefetch per field here; that instruction takes its address from a uniform register, so
// with a different address per lane the compiler emits a loop over the lanes -- 100 instructions per level and a third of
// k_pgs_pipe's stall samples, profiles/README.md.)
__device__ __forceinline__ void prefetch_rows_l2(const float4* rows, unsigned uNC, int q0, int q1, int nf, int lane, int width)
{
    const int nl = ((q1 - q0 + 7) >> 3) + 1;                     // lines a span can touch when q0 is not line aligned
    for (int idx = lane; idx < (nf + 1) * nl; idx += width) {
        const int ln = idx / (nf + 1), fi = idx - ln * (nf + 1);
        const int row = min(q0 + 8 * ln, q1 - 1);
        prefetch_l2(rows + ((unsigned)(fi < nf ? fi : CROW_LAMBDA) * uNC + (unsigned)row));
    }
}

// shared-memory accumulators of one scene, per body (linear xyz, angular xyz): either two padded float4 (32 B, two
// 16 B accesses: best when shared memory is plentiful) or three float2 (24 B: a 1000-body scene then needs 24 KB and 9
// scenes fit an SM instead of 7; chosen when the padded layout would cap the warps per SM)
template <bool W24> struct Acc;
template <> struct Acc<false> {
    enum { FLOATS = 8 };
    static __device__ __forceinline__ void ld(const float* w, int body, V3& l, V3& a)
    {
        const float4* p = reinterpret_cast<const float4*>(w) + 2 * body;
        l = mk3(p[0]); a = mk3(p[1]);
    }
    static __device__ __forceinline__ void st(float* w, int body, const V3& l, const V3& a)
    {
        float4* p = reinterpret_cast<float4*>(w) + 2 * body;
        p[0] = mk4(l, 0.f); p[1] = mk4(a, 0.f);
    }
};
template <> struct Acc<true> {
    enum { FLOATS = 6 };
    static __device__ __forceinline__ void ld(const float* w, int body, V3& l, V3& a)
    {
        const float2* p = reinterpret_cast<const float2*>(w) + 3 * body;
        const float2 p0 = p[0], p1 = p[1], p2 = p[2];
        l = mk3(p0.x, p0.y, p1.x); a = mk3(p1.y, p2.x, p2.y);
    }
    static __device__ __forceinline__ void st(float* w, int body, const V3& l, const V3& a)
    {
        float2* p = reinterpret_cast<float2*>(w) + 3 * body;
        p[0] = make_float2(l.x, l.y); p[1] = make_float2(l.z, a.x); p[2] = make_float2(a.y, a.z);
    }
};

// The same accumulators addressed in the shared state space (32-bit address, ld.shared / st.shared): through a generic
// pointer the compiler re-derives the shared window base (S2R SR_CgaCtaId, a ~25-cycle read, + MOV + LEA) in front of
// every gather and every scatter of the sweep.
template <bool W24> struct AccSh;
template <> struct AccSh<true> {
    static __device__ __forceinline__ void ld(unsigned base, int body, V3& l, V3& a)
    {
        const unsigned ad = base + (unsigned)body * 24u;
        float2 p0, p1, p2;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(p0.x), "=f"(p0.y) : "r"(ad));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+8];" : "=f"(p1.x), "=f"(p1.y) : "r"(ad));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+16];" : "=f"(p2.x), "=f"(p2.y) : "r"(ad));
        l = mk3(p0.x, p0.y, p1.x); a = mk3(p1.y, p2.x, p2.y);
    }
    static __device__ __forceinline__ void st(unsigned base, int body, const V3& l, const V3& a)
    {
        const unsigned ad = base + (unsigned)body * 24u;
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ad), "f"(l.x), "f"(l.y) : "memory");
        asm volatile("st.shared.v2.f32 [%0+8], {%1, %2};" ::"r"(ad), "f"(l.z), "f"(a.x) : "memory");
        asm volatile("st.shared.v2.f32 [%0+16], {%1, %2};" ::"r"(ad), "f"(a.y), "f"(a.z) : "memory");
    }
};
template <> struct AccSh<false> {
    static __device__ __forceinline__ void ld(unsigned base, int body, V3& l, V3& a)
    {
        const unsigned ad = base + (unsigned)body * 32u;
        float4 p0, p1;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p0.x), "=f"(p0.y), "=f"(p0.z), "=f"(p0.w) : "r"(ad));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(p1.x), "=f"(p1.y), "=f"(p1.z), "=f"(p1.w) : "r"(ad));
        l = mk3(p0); a = mk3(p1);
    }
    static __device__ __forceinline__ void st(unsigned base, int body, const V3& l, const V3& a)
    {
        const unsigned ad = base + (unsigned)body * 32u;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "f"(l.x), "f"(l.y), "f"(l.z), "f"(0.f) : "memory");
        asm volatile("st.shared.v4.f32 [%0+16], {%1, %2, %3, %4};" ::"r"(ad), "f"(a.x), "f"(a.y), "f"(a.z), "f"(0.f) : "memory");
    }
};

// level-ordered row streams ([scene][field][pos], layout.cuh mode 1) addressed without the layout-mode test of at_crow
__device__ __forceinline__ const float4* crow_scene(const View& v, int scene) { return v.crow + (size_t)scene * CROW_FIELDS * v.NC; }

}  // namespace slrbs
