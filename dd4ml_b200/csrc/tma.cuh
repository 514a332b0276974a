// TMA-staged variant of the streaming kernel for the ops that read many vectors per element
// (the L-SR1 pair ring plus 3-4 work vectors).
//
// One elected thread of a producer warp issues 1-D bulk async copies (cp.async.bulk, the TMA unit)
// of the next tiles of EVERY input stream into a multi-stage shared-memory ring and arms one
// mbarrier per stage with the expected byte count; the 8 consumer warps wait on the stage's
// mbarrier, run the op's arithmetic out of shared memory (LDS.128, conflict free) and store
// results straight to global memory.  Data in flight per SM = (stages - 1) x (#streams x 4 KB),
// independent of the register file, which is what the register-staged kernel cannot provide once
// an op has 25-55 fp64 accumulators.  Persistent grid: one CTA per SM.  The stage layout is static
// (stream sid at sid * 4 KB) so the consumer loop has no table look-ups or integer divisions.
#pragma once
#include <stdlib.h>

#include <type_traits>

#include "heavy.cuh"

namespace {

constexpr int TILE = BLOCK * 4;          // elements per stream per stage (ops with TMA_V = 4)
constexpr int MAX_STAGES = 6;
constexpr int SMEM_BUDGET = 200 * 1024;  // dynamic shared memory for the ring (+ ~14 KB static)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// data from the shared-memory ring (same interface as GlobalLoader, heavy.cuh).  Every stream
// of the TMA path has 4-byte elements (all-fp32 instantiations only), so stream `sid` sits at the
// compile-time offset sid * TILE * 4 of its stage: one base register + immediates, no lookups.
template <int STREAM_BYTES>
struct SmemLoader {
    static constexpr bool kStaged = true;   // data already on chip: ops may load just in time
    const unsigned char* mine;   // stage base + 4 V * (thread index within its consumer group)
    template <int V, class T>
    __device__ __forceinline__ void get(int sid, const T*, int64_t, bool, Raw<V, T>& r) const {
        static_assert((V == 4 || V == 2) && sizeof(T) == 4, "the smem ring holds fp32 streams, read 4 or 2 elements per thread");
        if constexpr (V == 4) {
            const float4 t = *reinterpret_cast<const float4*>(mine + sid * STREAM_BYTES);
            r.x[0] = t.x; r.x[1] = t.y; r.x[2] = t.z; r.x[3] = t.w;
        } else {
            const float2 t = *reinterpret_cast<const float2*>(mine + sid * STREAM_BYTES);
            r.x[0] = t.x; r.x[1] = t.y;
        }
    }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// CW consumer warps + 1 producer warp.  full[s]: armed by the producer with the byte count of the
// stage, completed by the TMA unit.  empty[s]: one arrival per consumer warp when it has read
// the stage; the producer waits for it before refilling.  No CTA-wide barrier in the loop, so
// the warps drift apart by up to `nstage` tiles and hide each other's latencies.
//
// Measured alternatives (n = 2^26, k = 3; profiles/README.md): (a) no producer warp, the last
// reader of a stage (shared-memory ticket) issues the refill: the ~10 bulk-copy issues land on the
// critical path of the slowest warp -- propose 0.86 -> 0.53 of peak, apply 1.05 -> 0.65;
// (b) 16 consumer warps x 2 elements: 17 warps cap ptxas at 96 registers and lose on every op.
// The consumer side is bound by dependent-instruction latency (cvt -> DFMA chains), not by a pipe.
//
// PW = 1: one producer warp (9 warps: ptxas may use 168 registers).  PW = 4: a whole producer
// warpgroup that hands its registers to the two consumer warpgroups (setmaxnreg 40 / 232) and
// exits after its loop -- for the ops whose accumulators do not fit 168 registers.
template <class Op, int TILE>
__device__ __forceinline__ void tma_produce(unsigned char* ring, uint64_t* full, uint64_t* empty,
                                            const void* const* s_ptr, uint32_t nlive, int64_t nv4, int nstage) {
    constexpr int STREAM_BYTES = TILE * 4;
    constexpr int STAGE_BYTES = Op::NS_MAX * STREAM_BYTES;
    const int64_t ntiles = (nv4 + TILE - 1) / TILE;
    int stg = 0;
    uint32_t ph = 1;   // a fresh mbarrier passes a wait on parity 1: the first lap does not block
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(&empty[stg], ph);
        const int64_t e0 = t * TILE;
        const uint32_t bytes = (uint32_t)((nv4 - e0 < TILE) ? (nv4 - e0) : TILE) * 4u;
        mbar_expect_tx(&full[stg], nlive * bytes);
        unsigned char* base = ring + (size_t)stg * STAGE_BYTES;
        for (int s = 0; s < Op::NS_MAX; ++s) {
            const void* p = s_ptr[s];
            if (p) bulk_g2s(base + s * STREAM_BYTES, reinterpret_cast<const unsigned char*>(p) + e0 * 4, bytes, &full[stg]);
        }
        if (++stg == nstage) { stg = 0; ph ^= 1u; }
    }
}

template <class Op, int CW, int PW>
__global__ void __launch_bounds__((CW + PW) * 32, 1) tma_kernel(Op op, int64_t n, double* ws, int nstage) {
    static_assert(PW == 1 || (PW == 4 && CW == 8), "register hand-over needs whole warpgroups");
    constexpr int CT = CW * 32;       // consumer threads
    constexpr int V = Op::TMA_V;      // elements per consumer thread per tile
    // SPLIT consumer groups take alternate tiles: a tile is then GT x V elements, i.e. the stage shrinks
    // (more stages in the 200 KB ring) while a thread still owns V consecutive elements -- what the ops
    // with 60+ accumulators want (one float32 -> float64 conversion per accumulator per V elements)
    constexpr int SPLIT = Op::TMA_SPLIT;
    constexpr int GT = CT / SPLIT;
    constexpr int TILE = GT * V;
    constexpr int STREAM_BYTES = TILE * 4;
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t full[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty[MAX_STAGES];
    __shared__ const void* s_ptr[Op::NS_MAX];
    __shared__ int s_nlive;
    constexpr int NR = Op::NRED;
    constexpr int NRA = NR > 0 ? NR : 1;
    constexpr int STAGE_BYTES = Op::NS_MAX * STREAM_BYTES;
    pdl_trigger();
    op.setup();
    const bool producer = threadIdx.x >= CT;
    const int64_t nv4 = n & ~(int64_t)3;
    const bool tma = op.active() && op.use_tma();
    if (threadIdx.x == 0) {
        int nlive = 0;
#pragma unroll
        for (int s = 0; s < Op::NS_MAX; ++s) {   // unrolled: no dynamic indexing into the op
            const void* p = nullptr;
            int esz = 0;
            op.stream_desc(s, p, esz);
            if (p && esz != 4) __trap();
            s_ptr[s] = p;
            nlive += p ? 1 : 0;
        }
        s_nlive = nlive;
        for (int s = 0; s < nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW / SPLIT); }
        fence_barrier_init();
    }
    __syncthreads();
    if constexpr (PW == 4) {
        if (producer) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
            if (tma && threadIdx.x == CT) tma_produce<Op, TILE>(ring, full, empty, s_ptr, (uint32_t)s_nlive, nv4, nstage);
            return;   // exited threads count as arrived at the barriers of the finish
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    } else {
        if (tma && threadIdx.x == CT) tma_produce<Op, TILE>(ring, full, empty, s_ptr, (uint32_t)s_nlive, nv4, nstage);
    }
    double acc[NRA];
#pragma unroll
    for (int j = 0; j < NRA; ++j) acc[j] = 0.0;
    if (op.active() && !producer) {
        if (tma) {
            const int64_t ntiles = (nv4 + TILE - 1) / TILE;
            const int grp = (int)threadIdx.x / GT, lt = (int)threadIdx.x - grp * GT;
            int stg = grp;            // the CTA's tiles in order use stages 0, 1, 2 ...; this group takes every SPLIT-th
            uint32_t ph = 0;
            const unsigned char* mine = ring + lt * (4 * V);
            for (int64_t t = blockIdx.x + (int64_t)grp * gridDim.x; t < ntiles; t += (int64_t)SPLIT * gridDim.x) {
                mbar_wait(&full[stg], ph);
                const int64_t i0 = t * TILE + (int64_t)lt * V;
                if (i0 < nv4) {
                    SmemLoader<STREAM_BYTES> ld{mine + stg * STAGE_BYTES};
                    op.template body_l<V>(ld, i0, acc);
                }
                __syncwarp();
                if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stg]);
                stg += SPLIT;
                if (stg >= nstage) { stg -= nstage; ph ^= 1u; }
            }
        } else {
            const int64_t tid = (int64_t)blockIdx.x * CT + threadIdx.x;
            const int64_t nthr = (int64_t)gridDim.x * CT;
            for (int64_t v = tid; v < (nv4 >> 2); v += nthr) op.template body<4>(v << 2, acc);
        }
        const int64_t tt = nv4 + (int64_t)blockIdx.x * CT + threadIdx.x;
        if (tt < n) op.template body<1>(tt, acc);
    }
    if constexpr (PW == 4) finish_reduce<Op, CW>(op, acc, ws);
    else finish_reduce<Op, CW + 1>(op, acc, ws);
}

// vectors at least this long go through the TMA ring (smaller ones are L2 resident and
// latency bound: the register-staged kernel has less fixed overhead there)
inline int64_t tma_min_n() {
    static int64_t v = -1;
    if (v < 0) {
        const char* e = getenv("DD4ML_B200_TMA_MIN_N");
        v = e ? atoll(e) : (int64_t)1 << 20;
    }
    return v;
}

inline int tma_mode() {   // 0: per-op default, 1: every heavy op, 2: none
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DD4ML_B200_TMA_MODE");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <class Op, int CW, int PW>
inline int launch_tma(Op op, int64_t n, double* ws, int nstage, int grid, size_t smem, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tma_kernel<Op, CW, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    tma_kernel<Op, CW, PW><<<grid, (CW + PW) * 32, smem, s>>>(op, n, ws, nstage);
    return (int)cudaGetLastError();
}

template <class Op>
inline int launch_heavy(Op op, int64_t n, bool vec, double* ws, int per_sm, int stage_bytes, bool tma_default,
                        cudaStream_t s) {
    static_assert(Op::NRED <= MAXRED, "too many fused reductions");
    (void)stage_bytes;   // the ring is laid out for all NS_MAX streams of the op
    constexpr int TILE = Op::TMA_CW * 32 / Op::TMA_SPLIT * Op::TMA_V;
    constexpr int STAGE_BYTES = Op::NS_MAX * TILE * 4;
    int nstage = SMEM_BUDGET / STAGE_BYTES;
    if (nstage > MAX_STAGES) nstage = MAX_STAGES;
    static_assert(SMEM_BUDGET / STAGE_BYTES >= Op::TMA_SPLIT, "every consumer group needs a stage of its own");
    const bool want = tma_mode() == 1 || (tma_mode() == 0 && tma_default);
    if (!want || !vec || nstage < 2 || n < tma_min_n()) return launch(op, n, vec, ws, per_sm, s);
    const int64_t ntiles = ((n & ~(int64_t)3) + TILE - 1) / TILE;
    int grid = (int)(ntiles < NUM_SMS ? (ntiles < 1 ? 1 : ntiles) : NUM_SMS);
    return launch_tma<Op, Op::TMA_CW, Op::TMA_PW>(op, n, ws, nstage, grid, (size_t)nstage * STAGE_BYTES, s);
}

// FAST = all-fp32 operands: TMA + register-staged variants for KMAX 4/6/8.  The mixed / fp64
// combinations (rare: PINN runs in float64) only get the register-staged KMAX = 8 kernel, which
// keeps the build small.
template <bool FAST, class Op>
inline int launch_sel(Op op, int64_t n, bool vec, double* ws, int per_sm, int stage_bytes, bool tma_default,
                      cudaStream_t s) {
    if constexpr (FAST) return launch_heavy(op, n, vec, ws, per_sm, stage_bytes, tma_default, s);
    else return launch(op, n, vec, ws, per_sm, s);
}

template <int K> using KC = std::integral_constant<int, K>;

// calls run(KC<4|6|8|MAXM>) for FAST combinations, run(KC<MAXM>) otherwise
#define DD4ML_RUN_K(FAST, kcap, run)                           \
    if constexpr (FAST) {                                      \
        if ((kcap) <= 4) return run(KC<4>{});                  \
        if ((kcap) <= 6) return run(KC<6>{});                  \
        if ((kcap) <= 8) return run(KC<8>{});                  \
    }                                                          \
    return run(KC<DD4ML_MAXM>{});

}  // namespace
