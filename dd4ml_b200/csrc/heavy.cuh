// Helpers shared by the heavy (L-SR1 memory streaming) ops.
#pragma once
#include "stream.cuh"

namespace {

// all live S_j / Y_j vectors of one iteration, loaded before any arithmetic
template <int V, class HT, int KMAX>
struct PairRegs {
    Raw<V, HT> s[KMAX > 0 ? KMAX : 1];
    Raw<V, HT> y[KMAX > 0 ? KMAX : 1];
    __device__ __forceinline__ void load(const HT* __restrict__ S, const HT* __restrict__ Y, int64_t ld,
                                         const Slots<KMAX>& sl, int64_t i) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < sl.k) {
                ldr_ro<V, HT>(s[j], S + (int64_t)sl.slot[j] * ld, i);
                ldr_ro<V, HT>(y[j], Y + (int64_t)sl.slot[j] * ld, i);
            }
    }
};

#define DD4ML_KCAP_SWITCH(kcap, CALLK) \
    if ((kcap) <= 4) { CALLK(4) } else { CALLK(8) }

}  // namespace
