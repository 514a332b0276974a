#pragma once
// dd4ml_b200 -- hand-written sm_100a kernels of the trust-region hot path and their C ABI.
//
// Every kernel is one instance of `stream_kernel<Op>` (register staged, this file) or of
// `tma_kernel<Op>` (TMA ring, tma.cuh): a persistent-grid, 128-bit vectorised streaming pass
// over the flat n-vectors (HBM-bound; no tensor cores), with
//   * up to MAXRED fused reductions accumulated in fp64 registers,
//   * warp-shuffle -> shared-memory block reduction,
//   * a deterministic two-stage finish: per-CTA partials + "last CTA done" ticket, the
//     last CTA sums the partials (one warp per column, a fixed tree for a given grid) and
//     stores the sums in the device state block; the k-space work on them (L-SR1 update chain,
//     OBS / dogleg solve ...) runs in the one-CTA k-space kernel that is enqueued right behind
//     with programmatic dependent launch (kspace_task.h) -- no host round trip.
// The Ops below fuse what the reference does with 10-40 ATen launches per TR step.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dd4ml_b200.h"
#include "kspace.h"
#include "kspace_task.h"

namespace {

constexpr int BLOCK = 256;
constexpr int NWARP = BLOCK / 32;
constexpr int NUM_SMS = 148;
constexpr int MAXGRID = NUM_SMS * 4;
constexpr int MAXRED = 72;
constexpr int WS_PARTIALS = MAXGRID * MAXRED;
constexpr int WS_DOUBLES = WS_PARTIALS + 8;  // + ticket counter

// ------------------------------------------------------------------ element access
template <int V, class T>
__device__ __forceinline__ void ld_ro(const T* __restrict__ p, int64_t i, double (&v)[V]);
template <>
__device__ __forceinline__ void ld_ro<4, float>(const float* __restrict__ p, int64_t i, double (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void ld_ro<4, double>(const double* __restrict__ p, int64_t i, double (&v)[4]) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p + i));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p + i + 2));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <>
__device__ __forceinline__ void ld_ro<1, float>(const float* __restrict__ p, int64_t i, double (&v)[1]) { v[0] = __ldg(p + i); }
template <>
__device__ __forceinline__ void ld_ro<1, double>(const double* __restrict__ p, int64_t i, double (&v)[1]) { v[0] = __ldg(p + i); }

// plain (coherent) loads for buffers the same kernel also writes
template <int V, class T>
__device__ __forceinline__ void ld_rw(const T* p, int64_t i, double (&v)[V]);
template <>
__device__ __forceinline__ void ld_rw<4, float>(const float* p, int64_t i, double (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p + i);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void ld_rw<4, double>(const double* p, int64_t i, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p + i);
    const double2 b = *reinterpret_cast<const double2*>(p + i + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <>
__device__ __forceinline__ void ld_rw<1, float>(const float* p, int64_t i, double (&v)[1]) { v[0] = p[i]; }
template <>
__device__ __forceinline__ void ld_rw<1, double>(const double* p, int64_t i, double (&v)[1]) { v[0] = p[i]; }

template <>
__device__ __forceinline__ void ld_ro<2, float>(const float* __restrict__ p, int64_t i, double (&v)[2]) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p + i));
    v[0] = t.x; v[1] = t.y;
}
template <>
__device__ __forceinline__ void ld_ro<2, double>(const double* __restrict__ p, int64_t i, double (&v)[2]) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p + i));
    v[0] = a.x; v[1] = a.y;
}
template <>
__device__ __forceinline__ void ld_rw<2, float>(const float* p, int64_t i, double (&v)[2]) {
    const float2 t = *reinterpret_cast<const float2*>(p + i);
    v[0] = t.x; v[1] = t.y;
}
template <>
__device__ __forceinline__ void ld_rw<2, double>(const double* p, int64_t i, double (&v)[2]) {
    const double2 a = *reinterpret_cast<const double2*>(p + i);
    v[0] = a.x; v[1] = a.y;
}

template <int V, class T>
__device__ __forceinline__ void st_v(T* p, int64_t i, const double (&v)[V]);
template <>
__device__ __forceinline__ void st_v<2, float>(float* p, int64_t i, const double (&v)[2]) {
    *reinterpret_cast<float2*>(p + i) = make_float2((float)v[0], (float)v[1]);
}
template <>
__device__ __forceinline__ void st_v<2, double>(double* p, int64_t i, const double (&v)[2]) {
    *reinterpret_cast<double2*>(p + i) = make_double2(v[0], v[1]);
}
template <>
__device__ __forceinline__ void st_v<4, float>(float* p, int64_t i, const double (&v)[4]) {
    *reinterpret_cast<float4*>(p + i) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
}
template <>
__device__ __forceinline__ void st_v<4, double>(double* p, int64_t i, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p + i) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + i + 2) = make_double2(v[2], v[3]);
}
template <>
__device__ __forceinline__ void st_v<1, float>(float* p, int64_t i, const double (&v)[1]) { p[i] = (float)v[0]; }
template <>
__device__ __forceinline__ void st_v<1, double>(double* p, int64_t i, const double (&v)[1]) { p[i] = v[0]; }

// Raw (unconverted) V-element vectors: all loads of one loop iteration are issued into these
// BEFORE any arithmetic, so a thread has (#streams x 16 B) in flight per iteration and pays one
// memory round trip per iteration instead of one per stream.
template <int V, class T> struct Raw { T x[V]; };
template <int V, class T> __device__ __forceinline__ void ldr_ro(Raw<V, T>& r, const T* __restrict__ p, int64_t i);
template <> __device__ __forceinline__ void ldr_ro<4, float>(Raw<4, float>& r, const float* __restrict__ p, int64_t i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
    r.x[0] = t.x; r.x[1] = t.y; r.x[2] = t.z; r.x[3] = t.w;
}
template <> __device__ __forceinline__ void ldr_ro<4, double>(Raw<4, double>& r, const double* __restrict__ p, int64_t i) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p + i));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p + i + 2));
    r.x[0] = a.x; r.x[1] = a.y; r.x[2] = b.x; r.x[3] = b.y;
}
template <> __device__ __forceinline__ void ldr_ro<2, float>(Raw<2, float>& r, const float* __restrict__ p, int64_t i) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p + i));
    r.x[0] = t.x; r.x[1] = t.y;
}
template <> __device__ __forceinline__ void ldr_ro<2, double>(Raw<2, double>& r, const double* __restrict__ p, int64_t i) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p + i));
    r.x[0] = a.x; r.x[1] = a.y;
}
template <> __device__ __forceinline__ void ldr_ro<1, float>(Raw<1, float>& r, const float* __restrict__ p, int64_t i) { r.x[0] = __ldg(p + i); }
template <> __device__ __forceinline__ void ldr_ro<1, double>(Raw<1, double>& r, const double* __restrict__ p, int64_t i) { r.x[0] = __ldg(p + i); }
template <int V, class T> __device__ __forceinline__ void ldr_rw(Raw<V, T>& r, const T* p, int64_t i);
template <> __device__ __forceinline__ void ldr_rw<4, float>(Raw<4, float>& r, const float* p, int64_t i) {
    const float4 t = *reinterpret_cast<const float4*>(p + i);
    r.x[0] = t.x; r.x[1] = t.y; r.x[2] = t.z; r.x[3] = t.w;
}
template <> __device__ __forceinline__ void ldr_rw<4, double>(Raw<4, double>& r, const double* p, int64_t i) {
    const double2 a = *reinterpret_cast<const double2*>(p + i);
    const double2 b = *reinterpret_cast<const double2*>(p + i + 2);
    r.x[0] = a.x; r.x[1] = a.y; r.x[2] = b.x; r.x[3] = b.y;
}
template <> __device__ __forceinline__ void ldr_rw<2, float>(Raw<2, float>& r, const float* p, int64_t i) {
    const float2 t = *reinterpret_cast<const float2*>(p + i);
    r.x[0] = t.x; r.x[1] = t.y;
}
template <> __device__ __forceinline__ void ldr_rw<2, double>(Raw<2, double>& r, const double* p, int64_t i) {
    const double2 a = *reinterpret_cast<const double2*>(p + i);
    r.x[0] = a.x; r.x[1] = a.y;
}
template <> __device__ __forceinline__ void ldr_rw<1, float>(Raw<1, float>& r, const float* p, int64_t i) { r.x[0] = p[i]; }
template <> __device__ __forceinline__ void ldr_rw<1, double>(Raw<1, double>& r, const double* p, int64_t i) { r.x[0] = p[i]; }

// round a double to the storage dtype and back (what a store + reload would give)
template <class T> __device__ __forceinline__ double rnd(double x) { return (double)(T)x; }

__device__ __forceinline__ double ld_scalar(const void* p, int dt) {
    return dt == 0 ? (double)*reinterpret_cast<const float*>(p) : *reinterpret_cast<const double*>(p);
}
__device__ __forceinline__ void st_scalar(void* p, int dt, double v) {
    if (dt == 0) *reinterpret_cast<float*>(p) = (float)v; else *reinterpret_cast<double*>(p) = v;
}

// live slots of the L-SR1 ring, read once per thread from the state block
// Only ever indexed with compile-time constants (unrolled loops, at()): a dynamic index would
// push the whole op into local memory and turn every field access of the hot loop into an LDL.
template <int KMAX>
struct Slots {
    int k;
    unsigned mask;   // bit j set: pair j is streamed
    int slot[KMAX > 0 ? KMAX : 1];
    __device__ __forceinline__ void load(const dd4ml_state* st, bool use_memory) {
        k = use_memory ? (int)st->k[0] : 0;
        if (k > KMAX) k = KMAX;
        mask = (1u << k) - 1u;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) slot[j] = (j < k) ? (int)st->order[j] : 0;
    }
    __device__ __forceinline__ bool live(int j) const { return (mask >> j) & 1u; }
    __device__ __forceinline__ int at(int j) const {   // runtime j (finalize only)
        int r = slot[0];
#pragma unroll
        for (int i = 1; i < KMAX; ++i) r = (j == i) ? slot[i] : r;
        return r;
    }
};

// ------------------------------------------------------------------ reduction finish
// warp shuffle -> shared memory -> per-CTA partials -> ticket; the last CTA sums the partials
// with a fixed tree (bit-reproducible for a given grid) and runs the op's k-space routine.
template <class Op, int NW = NWARP>
__device__ __forceinline__ void finish_reduce(Op& op, double* acc, double* ws) {
    constexpr int NR = Op::NRED;
    constexpr int NRA = NR > 0 ? NR : 1;
    if constexpr (NR > 0) {
        __shared__ double s_red[NW][NRA];
        __shared__ double s_sum[NRA];
        __shared__ int s_last;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            double v = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double u = __shfl_xor_sync(0xffffffffu, v, o);
                v = Op::is_max(j) ? fmax(v, u) : v + u;
            }
            if (lane == 0) s_red[warp][j] = v;
        }
        __syncthreads();
        double* partials = ws;
        unsigned int* ticket = reinterpret_cast<unsigned int*>(ws + WS_PARTIALS);
        if (threadIdx.x < NR) {
            const int j = threadIdx.x;
            double v = s_red[0][j];
            for (int w = 1; w < NW; ++w) v = Op::is_max(j) ? fmax(v, s_red[w][j]) : v + s_red[w][j];
            partials[(int64_t)blockIdx.x * NR + j] = v;
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int t = atomicAdd(ticket, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            // one warp per column: lanes stride over the CTAs' partials, then a butterfly -- a fixed
            // tree for a given grid (bit-reproducible), and ~G/32 dependent L2 reads instead of G
            for (int j = warp; j < NR; j += NW) {
                double v = 0.0;
                for (unsigned int b = lane; b < gridDim.x; b += 32) {
                    const double u = __ldcg(partials + (int64_t)b * NR + j);
                    v = Op::is_max(j) ? fmax(v, u) : v + u;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double u = __shfl_xor_sync(0xffffffffu, v, o);
                    v = Op::is_max(j) ? fmax(v, u) : v + u;
                }
                if (lane == 0) s_sum[j] = v;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                *ticket = 0u;
                op.finalize(s_sum, op.state());
            }
        }
    }
}

// programmatic dependent launch: let the k-space kernel that follows become resident now
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

// ------------------------------------------------------------------ the streaming kernel
template <class Op>
__global__ void __launch_bounds__(BLOCK) stream_kernel(Op op, int64_t n, int vec, double* ws) {
    constexpr int NR = Op::NRED;
    constexpr int NRA = NR > 0 ? NR : 1;
    double acc[NRA];
#pragma unroll
    for (int j = 0; j < NRA; ++j) acc[j] = 0.0;
    pdl_trigger();
    op.setup();
    if (op.active()) {
        const int64_t tid = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
        const int64_t nthr = (int64_t)gridDim.x * BLOCK;
        if (vec) {
            const int64_t nv = n >> 2;
            for (int64_t v = tid; v < nv; v += nthr) op.template body<4>(v << 2, acc);
            const int64_t t = (nv << 2) + tid;
            if (t < n) op.template body<1>(t, acc);
        } else {
            for (int64_t i = tid; i < n; i += nthr) op.template body<1>(i, acc);
        }
    }
    finish_reduce(op, acc, ws);
}

static_assert(BLOCK >= MAXRED, "one thread per reduction column in the finish");

// ------------------------------------------------------------------ launch helpers
inline int grid_for(int64_t n, int per_sm) {
    const int64_t work = (n / 4 + BLOCK * 4 - 1) / (BLOCK * 4);  // >= 4 vector iterations per thread when possible
    int64_t cap = (int64_t)NUM_SMS * per_sm;
    if (cap > MAXGRID) cap = MAXGRID;
    int64_t g = work < 1 ? 1 : work;
    if (g > cap) g = cap;
    return (int)g;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <class Op>
inline int launch(Op op, int64_t n, bool vec, double* ws, int per_sm, cudaStream_t s) {
    static_assert(Op::NRED <= MAXRED, "too many fused reductions");
    const int grid = grid_for(n, per_sm);
    stream_kernel<Op><<<grid, BLOCK, 0, s>>>(op, n, vec ? 1 : 0, ws);
    return (int)cudaGetLastError();
}

}  // namespace


#define DT_SWITCH3(ht, vt, wt, CALL)                                              \
    if (ht == 0 && vt == 0 && wt == 0) { CALL(float, float, float) }              \
    else if (ht == 0 && vt == 0 && wt == 1) { CALL(float, float, double) }        \
    else if (ht == 0 && vt == 1 && wt == 1) { CALL(float, double, double) }       \
    else if (ht == 1 && vt == 1 && wt == 1) { CALL(double, double, double) }      \
    else return DD4ML_E_DTYPE;

#define DT_SWITCH2(ht, vt, CALL)                                                  \
    if (ht == 0 && vt == 0) { CALL(float, float) }                                \
    else if (ht == 0 && vt == 1) { CALL(float, double) }                          \
    else if (ht == 1 && vt == 1) { CALL(double, double) }                         \
    else return DD4ML_E_DTYPE;


// variants that assign a return code instead of returning (the entry point goes on to enqueue
// the k-space kernel)
#define DT_SWITCH3_RC(ht, vt, wt, CALL)                                           \
    if (ht == 0 && vt == 0 && wt == 0) { CALL(float, float, float) }              \
    else if (ht == 0 && vt == 0 && wt == 1) { CALL(float, float, double) }        \
    else if (ht == 0 && vt == 1 && wt == 1) { CALL(float, double, double) }       \
    else if (ht == 1 && vt == 1 && wt == 1) { CALL(double, double, double) }

#define DT_SWITCH2_RC(ht, vt, CALL)                                               \
    if (ht == 0 && vt == 0) { CALL(float, float) }                                \
    else if (ht == 0 && vt == 1) { CALL(float, double) }                          \
    else if (ht == 1 && vt == 1) { CALL(double, double) }
