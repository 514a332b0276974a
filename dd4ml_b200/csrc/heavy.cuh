// Helpers shared by the heavy (L-SR1 memory streaming) ops.
#pragma once
#include "stream.cuh"

namespace {

// Loader handed to Op::body_l: data straight from global memory (register staged).  The TMA
// kernel passes a SmemLoader with the same interface (tma.cuh).
struct GlobalLoader {
    static constexpr bool kStaged = false;  // every load is a memory round trip: ops issue them all up front
    template <int V, class T>
    __device__ __forceinline__ void get(int, const T* __restrict__ gptr, int64_t i, bool rw, Raw<V, T>& r) const {
        if (rw) ldr_rw<V, T>(r, gptr, i); else ldr_ro<V, T>(r, gptr, i);
    }
};

// all live S_j / Y_j vectors of one iteration, loaded before any arithmetic.  Stream ids:
// S_j -> F + j, Y_j -> F + KMAX + j (F = number of fixed streams of the op).
template <int V, class HT, int KMAX>
struct PairRegs {
    Raw<V, HT> s[KMAX > 0 ? KMAX : 1];
    Raw<V, HT> y[KMAX > 0 ? KMAX : 1];
    template <class L>
    __device__ __forceinline__ void load(const L& l, int F, const HT* __restrict__ S, const HT* __restrict__ Y,
                                         int64_t ld, const Slots<KMAX>& sl, int64_t i) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (sl.live(j)) {
                l.template get<V, HT>(F + j, S + (int64_t)sl.slot[j] * ld, i, false, s[j]);
                l.template get<V, HT>(F + KMAX + j, Y + (int64_t)sl.slot[j] * ld, i, false, y[j]);
            }
    }
    // descriptor of pair stream `r` (0 .. 2*KMAX-1) for the TMA producer
    __device__ __forceinline__ static void desc(int r, const HT* S, const HT* Y, int64_t ld, const Slots<KMAX>& sl,
                                                const void*& p, int& esz) {
        const int j = r < KMAX ? r : r - KMAX;
        esz = (int)sizeof(HT);
        p = sl.live(j) ? (const void*)((r < KMAX ? S : Y) + (int64_t)sl.slot[j] * ld) : nullptr;
    }
};

// KMAX = 10 variants (the reference's default memory length): 20+ pair streams make the consumers of the
// TMA ring latency bound on the LDS -> F2F.F64.F32 (quarter-rate XU pipe) -> DFMA chain with 40-70 fp64
// accumulators (ncu, round 1: XU 31-44 %, F2F 17 % of the stall samples).  For all-float32 operands the
// products of one thread-tile (V consecutive elements) are therefore summed in float32 FFMA chains and
// only the per-tile partial is converted and added to the fp64 accumulator: V x fewer conversions,
// 4-cycle FFMA latency in the chain.  The partial's rounding (V products, ~1e-7 relative) is below the
// reference's own float32 dot products; sums stay deterministic.
template <class A, class B> struct BothF32 { static constexpr bool value = false; };
template <> struct BothF32<float, float> { static constexpr bool value = true; };
template <int V>
__device__ __forceinline__ float dot32(const float (&a)[V], const float (&b)[V]) {
    float s = a[0] * b[0];
#pragma unroll
    for (int e = 1; e < V; ++e) s = fmaf(a[e], b[e], s);
    return s;
}

#define DD4ML_KCAP_SWITCH(kcap, CALLK) \
    if ((kcap) <= 4) { CALLK(4) } else if ((kcap) <= 6) { CALLK(6) } else { CALLK(DD4ML_MAXM) }

}  // namespace
