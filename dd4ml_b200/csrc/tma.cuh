// TMA-staged variant of the streaming kernel for the ops that read many vectors per element
// (the L-SR1 pair ring plus 3-4 work vectors).
//
// One elected thread per CTA issues 1-D bulk async copies (cp.async.bulk, the TMA unit) of the
// next tiles of EVERY input stream into a multi-stage shared-memory ring and arms one mbarrier
// per stage with the expected byte count; the 8 warps wait on the stage's mbarrier, run the
// op's arithmetic out of shared memory (LDS.128, conflict free) and store results straight
// to global memory.  Data in flight per SM = (stages - 1) x (#streams x 4 KB), independent of
// the register file, which is what the register-staged kernel cannot provide once an op has
// 25-55 fp64 accumulators.  Persistent grid: one CTA per SM.
#pragma once
#include <stdlib.h>

#include <type_traits>

#include "heavy.cuh"

namespace {

constexpr int TILE = BLOCK * 4;          // elements per stream per stage
constexpr int MAX_STAGES = 4;
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

// data from the shared-memory ring (same interface as GlobalLoader, heavy.cuh)
struct SmemLoader {
    const unsigned char* stage;
    const int* off;
    int local;
    template <int V, class T>
    __device__ __forceinline__ void get(int sid, const T*, int64_t, bool, Raw<V, T>& r) const {
        static_assert(V == 4, "the smem ring is read 4 elements per thread");
        const T* p = reinterpret_cast<const T*>(stage + off[sid]) + local;
        if constexpr (sizeof(T) == 4) {
            const float4 t = *reinterpret_cast<const float4*>(p);
            r.x[0] = t.x; r.x[1] = t.y; r.x[2] = t.z; r.x[3] = t.w;
        } else {
            const double2 a = *reinterpret_cast<const double2*>(p);
            const double2 b = *reinterpret_cast<const double2*>(p + 2);
            r.x[0] = a.x; r.x[1] = a.y; r.x[2] = b.x; r.x[3] = b.y;
        }
    }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 8 consumer warps + 1 producer warp.  full[s]: armed by the producer with the byte count of the
// stage, completed by the TMA unit.  empty[s]: one arrival per consumer warp when it has read
// the stage; the producer waits for it before refilling.  No CTA-wide barrier in the loop, so
// the warps drift apart by up to `nstage` tiles and hide each other's latencies.
constexpr int TMA_THREADS = BLOCK + 32;

template <class Op>
__global__ void __launch_bounds__(TMA_THREADS, 1) tma_kernel(Op op, int64_t n, double* ws, int nstage, int stage_bytes) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t full[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty[MAX_STAGES];
    __shared__ const void* s_ptr[Op::NS_MAX];
    __shared__ int s_esz[Op::NS_MAX];
    __shared__ int s_off[Op::NS_MAX];
    constexpr int NR = Op::NRED;
    constexpr int NRA = NR > 0 ? NR : 1;
    double acc[NRA];
#pragma unroll
    for (int j = 0; j < NRA; ++j) acc[j] = 0.0;
    op.setup();
    const bool producer = threadIdx.x >= BLOCK;
    if (threadIdx.x == 0) {
        int off = 0;
        for (int s = 0; s < Op::NS_MAX; ++s) {
            const void* p = nullptr;
            int esz = 0;
            op.stream_desc(s, p, esz);
            s_ptr[s] = p;
            s_esz[s] = p ? esz : 0;
            s_off[s] = off;
            if (p) off += TILE * esz;
        }
        for (int s = 0; s < nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWARP); }
        fence_barrier_init();
    }
    __syncthreads();
    if (op.active()) {
        const int64_t nv4 = n & ~(int64_t)3;
        if (op.use_tma()) {
            const int64_t ntiles = (nv4 + TILE - 1) / TILE;
            if (producer) {
                if (threadIdx.x == BLOCK) {
                    int it = 0;
                    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                        const int stg = it % nstage;
                        if (it >= nstage) mbar_wait(&empty[stg], (uint32_t)(((it / nstage) - 1) & 1));
                        const int64_t e0 = t * TILE;
                        const int64_t cnt = (nv4 - e0 < TILE) ? (nv4 - e0) : TILE;
                        uint32_t total = 0;
                        for (int s = 0; s < Op::NS_MAX; ++s) total += (uint32_t)(cnt * s_esz[s]);
                        mbar_expect_tx(&full[stg], total);
                        unsigned char* base = ring + (size_t)stg * stage_bytes;
                        for (int s = 0; s < Op::NS_MAX; ++s)
                            if (s_esz[s])
                                bulk_g2s(base + s_off[s],
                                         reinterpret_cast<const unsigned char*>(s_ptr[s]) + e0 * s_esz[s],
                                         (uint32_t)(cnt * s_esz[s]), &full[stg]);
                    }
                }
            } else {
                int it = 0;
                for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                    const int stg = it % nstage;
                    mbar_wait(&full[stg], (uint32_t)((it / nstage) & 1));
                    const int64_t i0 = t * TILE + (int64_t)threadIdx.x * 4;
                    if (i0 < nv4) {
                        SmemLoader ld{ring + (size_t)stg * stage_bytes, s_off, (int)threadIdx.x * 4};
                        op.template body_l<4>(ld, i0, acc);
                    }
                    __syncwarp();
                    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stg]);
                }
            }
        } else if (!producer) {
            const int64_t tid = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
            const int64_t nthr = (int64_t)gridDim.x * BLOCK;
            for (int64_t v = tid; v < (nv4 >> 2); v += nthr) op.template body<4>(v << 2, acc);
        }
        const int64_t tt = nv4 + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
        if (!producer && tt < n) op.template body<1>(tt, acc);
    }
    finish_reduce<Op, NWARP + 1>(op, acc, ws);
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

template <class Op>
inline int launch_heavy(Op op, int64_t n, bool vec, double* ws, int per_sm, int stage_bytes, bool tma_default,
                        cudaStream_t s) {
    static_assert(Op::NRED <= MAXRED, "too many fused reductions");
    int nstage = stage_bytes > 0 ? SMEM_BUDGET / stage_bytes : 0;
    if (nstage > MAX_STAGES) nstage = MAX_STAGES;
    const bool want = tma_mode() == 1 || (tma_mode() == 0 && tma_default);
    if (!want || !vec || nstage < 2 || n < tma_min_n()) return launch(op, n, vec, ws, per_sm, s);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tma_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int64_t ntiles = ((n & ~(int64_t)3) + TILE - 1) / TILE;
    int grid = (int)(ntiles < NUM_SMS ? (ntiles < 1 ? 1 : ntiles) : NUM_SMS);
    tma_kernel<Op><<<grid, TMA_THREADS, (size_t)nstage * stage_bytes, s>>>(op, n, ws, nstage, stage_bytes);
    return (int)cudaGetLastError();
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

// calls run(KC<4|6|8>) for FAST combinations, run(KC<8>) otherwise
#define DD4ML_RUN_K(FAST, kcap, run)                           \
    if constexpr (FAST) {                                      \
        if ((kcap) <= 4) return run(KC<4>{});                  \
        if ((kcap) <= 6) return run(KC<6>{});                  \
    }                                                          \
    return run(KC<8>{});

}  // namespace
