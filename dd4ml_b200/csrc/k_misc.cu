// dd4ml_b200 -- state block plumbing, multi-tensor gather / segmented copy, introspection.
#include "stream.cuh"
#include "state_table.h"

namespace {

constexpr int GATHER_MAX = 96;
struct GatherTable {
    const void* src[GATHER_MAX];
    int64_t off[GATHER_MAX];
    int64_t len[GATHER_MAX];
    int count;
};
template <class T>
__global__ void __launch_bounds__(BLOCK) gather_kernel(T* dst, GatherTable tb) {
    // blockIdx.y = tensor, blockIdx.x strides over its elements
    const int t = blockIdx.y;
    const T* s = reinterpret_cast<const T*>(tb.src[t]);
    T* d = dst + tb.off[t];
    const int64_t len = tb.len[t];
    for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < len; i += (int64_t)gridDim.x * BLOCK) d[i] = s[i];
}

constexpr int SEG_MAX = 96;
struct SegTable {
    int64_t dst_off[SEG_MAX];
    int64_t src_off[SEG_MAX];
    int64_t len[SEG_MAX];
    int count;
};
template <class T>
__global__ void __launch_bounds__(BLOCK) seg_copy_kernel(T* dst, const T* src, SegTable tb) {
    const int t = blockIdx.y;
    const T* s = src + tb.src_off[t];
    T* d = dst + tb.dst_off[t];
    const int64_t len = tb.len[t];
    for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < len; i += (int64_t)gridDim.x * BLOCK) d[i] = s[i];
}

// dst[r] = src[idx[r]] for rows of `units` U-byte words: one thread per word, consecutive threads on
// consecutive words of a row (coalesced); rows with an index outside [0, src_rows) are skipped
template <class U>
__global__ void __launch_bounds__(BLOCK) gather_rows_kernel(U* __restrict__ dst, const U* __restrict__ src,
                                                            const int64_t* __restrict__ idx, int64_t nrows,
                                                            int64_t units, int64_t src_rows) {
    const int64_t total = nrows * units;
    for (int64_t c = (int64_t)blockIdx.x * BLOCK + threadIdx.x; c < total; c += (int64_t)gridDim.x * BLOCK) {
        const int64_t r = c / units, k = c - r * units;
        const int64_t sr = __ldg(idx + r);
        if (sr >= 0 && sr < src_rows) dst[c] = __ldg(src + sr * units + k);
    }
}

// OBS.phiBar_f / phiBar_fg / Newton (obs.py:184-254) as stand-alone calls: one thread, the same
// routines the fused solve uses (kspace.h).  out = {f, g} (modes 0, 1) or {sigma, iterations,
// degenerate evaluations} (mode 2).
template <class T>
__global__ void obs_secular_kernel(int mode, double x0, const T* Lam, const T* a, int m, double delta, double tol,
                                   double* out) {
    T L[DD4ML_MAXM + 1], A[DD4ML_MAXM + 1];
    for (int j = 0; j < m; ++j) { L[j] = Lam[j]; A[j] = a[j]; }
    if (mode == 2) {
        int it = 0, nd = 0;
        const T x = ks::newton<T>((T)x0, L, A, m, (T)delta, (T)tol, it, nd);
        out[0] = (double)x; out[1] = (double)it; out[2] = (double)nd;
    } else {
        T f, g;
        bool deg;
        ks::phibar_fg<T>((T)x0, L, A, m, (T)delta, (T)tol, f, g, deg);
        out[0] = (double)f; out[1] = (double)g; out[2] = deg ? 1.0 : 0.0;
    }
}

constexpr int SET_MAX = 32;
struct SetTable {
    int off[SET_MAX];
    double val[SET_MAX];
    int count;
};
__global__ void state_set_kernel(double* st, SetTable tb) {
    const int i = threadIdx.x;
    if (i < tb.count) st[tb.off[i]] = tb.val[i];
}
__global__ void state_init_kernel(double* st, int nst, double* ws, int nws) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nst; i += gridDim.x * blockDim.x) st[i] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nws; i += gridDim.x * blockDim.x) ws[i] = 0.0;
}
// dst[dst_off] = clamp(src[src_off] * scale, lo, hi): radius bookkeeping between optimizer state
// blocks (apts_base.py:228-237, apts_p.py:169,236) without a host round trip
__global__ void state_link_kernel(double* dst, int dst_off, const double* src, int src_off, double scale, double lo,
                                  double hi) {
    double v = src[src_off] * scale;
    v = v < lo ? lo : v;
    v = v > hi ? hi : v;
    dst[dst_off] = v;
}
__global__ void state_defaults_kernel(dd4ml_state* st, double gamma0) {
    st->gamma[0] = gamma0;
    st->spare[0] = 0.0;
    st->k[0] = 0.0;
    st->clip_scale[0] = 1.0;
}


}  // namespace

extern "C" {

const char* dd4ml_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case DD4ML_E_DTYPE: return "dd4ml_b200: unsupported dtype combination";
        case DD4ML_E_ARG: return "dd4ml_b200: bad argument";
        case DD4ML_E_ALIGN: return "dd4ml_b200: misaligned pointer";
        default: return cudaGetErrorString((cudaError_t)code);
    }
}

int dd4ml_version(void) { return 100; }
int dd4ml_state_doubles(void) { return (int)(sizeof(dd4ml_state) / sizeof(double)); }
int dd4ml_workspace_doubles(void) { return WS_DOUBLES; }
int dd4ml_num_fields(void) { return DD4ML_NUM_FIELDS; }
int dd4ml_max_mem_length(void) { return DD4ML_MAXM; }
int dd4ml_field(int i, const char** name, int* offset, int* count) {
    if (i < 0 || i >= DD4ML_NUM_FIELDS) return DD4ML_E_ARG;
    *name = DD4ML_FIELD_TABLE[i].name;
    *offset = DD4ML_FIELD_TABLE[i].offset;
    *count = DD4ML_FIELD_TABLE[i].count;
    return 0;
}

int dd4ml_state_init(double* state, double* ws, double gamma0, void* stream) {
    if (!state || !ws) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    state_init_kernel<<<32, 256, 0, s>>>(state, dd4ml_state_doubles(), ws, WS_DOUBLES);
    state_defaults_kernel<<<1, 1, 0, s>>>((dd4ml_state*)state, gamma0);
    return (int)cudaGetLastError();
}

int dd4ml_state_set(double* state, int count, const int* offsets, const double* values, void* stream) {
    if (!state || count < 0 || count > SET_MAX) return DD4ML_E_ARG;
    SetTable tb;
    tb.count = count;
    const int nst = dd4ml_state_doubles();
    for (int i = 0; i < count; ++i) {
        if (offsets[i] < 0 || offsets[i] >= nst) return DD4ML_E_ARG;
        tb.off[i] = offsets[i];
        tb.val[i] = values[i];
    }
    state_set_kernel<<<1, SET_MAX, 0, (cudaStream_t)stream>>>(state, tb);
    return (int)cudaGetLastError();
}

int dd4ml_state_link(double* dst, int dst_off, const double* src, int src_off, double scale, double lo, double hi,
                     void* stream) {
    const int nst = dd4ml_state_doubles();
    if (!dst || !src || dst_off < 0 || dst_off >= nst || src_off < 0 || src_off >= nst) return DD4ML_E_ARG;
    state_link_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dst, dst_off, src, src_off, scale, lo, hi);
    return (int)cudaGetLastError();
}

int dd4ml_gather(void* dst, int count, const void* const* srcs, const int64_t* offsets,
                 const int64_t* lens, int dtype, void* stream) {
    if (!dst || count < 0 || (dtype != 0 && dtype != 1)) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    for (int base = 0; base < count; base += GATHER_MAX) {
        GatherTable tb;
        tb.count = (count - base < GATHER_MAX) ? count - base : GATHER_MAX;
        int64_t maxlen = 1;
        for (int i = 0; i < tb.count; ++i) {
            tb.src[i] = srcs[base + i];
            tb.off[i] = offsets[base + i];
            tb.len[i] = lens[base + i];
            if (tb.len[i] > maxlen) maxlen = tb.len[i];
        }
        int gx = (int)((maxlen + BLOCK * 8 - 1) / (BLOCK * 8));
        if (gx < 1) gx = 1;
        if (gx > 64) gx = 64;
        dim3 grid(gx, tb.count);
        if (dtype == 0) gather_kernel<float><<<grid, BLOCK, 0, s>>>((float*)dst, tb);
        else gather_kernel<double><<<grid, BLOCK, 0, s>>>((double*)dst, tb);
    }
    return (int)cudaGetLastError();
}

int dd4ml_seg_copy(void* dst, const void* src, int count, const int64_t* dst_off,
                   const int64_t* src_off, const int64_t* lens, int64_t n_dst, int zero_fill,
                   int dtype, void* stream) {
    if (!dst || !src || count < 0 || (dtype != 0 && dtype != 1)) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (zero_fill) {
        cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)n_dst * (dtype == 0 ? 4 : 8), s);
        if (e != cudaSuccess) return (int)e;
    }
    for (int base = 0; base < count; base += SEG_MAX) {
        SegTable tb;
        tb.count = (count - base < SEG_MAX) ? count - base : SEG_MAX;
        int64_t maxlen = 1;
        for (int i = 0; i < tb.count; ++i) {
            tb.dst_off[i] = dst_off[base + i];
            tb.src_off[i] = src_off[base + i];
            tb.len[i] = lens[base + i];
            if (tb.len[i] > maxlen) maxlen = tb.len[i];
        }
        int gx = (int)((maxlen + BLOCK * 8 - 1) / (BLOCK * 8));
        if (gx < 1) gx = 1;
        if (gx > 64) gx = 64;
        dim3 grid(gx, tb.count);
        if (dtype == 0) seg_copy_kernel<float><<<grid, BLOCK, 0, s>>>((float*)dst, (const float*)src, tb);
        else seg_copy_kernel<double><<<grid, BLOCK, 0, s>>>((double*)dst, (const double*)src, tb);
    }
    return (int)cudaGetLastError();
}

int dd4ml_obs_secular(int mode, double x0, const void* Lam, const void* a, int m, double delta, double tol,
                      int dtype, double* out, void* stream) {
    if (!Lam || !a || !out || m < 1 || m > DD4ML_MAXM + 1 || mode < 0 || mode > 2) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == 0) obs_secular_kernel<float><<<1, 1, 0, s>>>(mode, x0, (const float*)Lam, (const float*)a, m, delta, tol, out);
    else if (dtype == 1) obs_secular_kernel<double><<<1, 1, 0, s>>>(mode, x0, (const double*)Lam, (const double*)a, m, delta, tol, out);
    else return DD4ML_E_DTYPE;
    return (int)cudaGetLastError();
}

int dd4ml_gather_rows(void* dst, const void* src, const int64_t* idx, int64_t nrows, int64_t row_bytes,
                      int64_t src_rows, void* stream) {
    if (!dst || !src || !idx || nrows < 0 || row_bytes <= 0 || src_rows < 0) return DD4ML_E_ARG;
    if (nrows == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const uintptr_t a = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | (uintptr_t)row_bytes;
    const int64_t total_bytes = nrows * row_bytes;
    auto grid_for_units = [&](int64_t unit) {
        int64_t g = (total_bytes / unit + BLOCK - 1) / BLOCK;
        const int64_t cap = (int64_t)NUM_SMS * 8;
        return (int)(g < 1 ? 1 : (g > cap ? cap : g));
    };
    if ((a & 15u) == 0)
        gather_rows_kernel<uint4><<<grid_for_units(16), BLOCK, 0, s>>>((uint4*)dst, (const uint4*)src, idx, nrows, row_bytes / 16, src_rows);
    else if ((a & 7u) == 0)
        gather_rows_kernel<uint2><<<grid_for_units(8), BLOCK, 0, s>>>((uint2*)dst, (const uint2*)src, idx, nrows, row_bytes / 8, src_rows);
    else if ((a & 3u) == 0)
        gather_rows_kernel<uint32_t><<<grid_for_units(4), BLOCK, 0, s>>>((uint32_t*)dst, (const uint32_t*)src, idx, nrows, row_bytes / 4, src_rows);
    else
        gather_rows_kernel<unsigned char><<<grid_for_units(1), BLOCK, 0, s>>>((unsigned char*)dst, (const unsigned char*)src, idx, nrows, row_bytes, src_rows);
    return (int)cudaGetLastError();
}

}  // extern "C"
