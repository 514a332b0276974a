// TRAdam.step (optimizers/tradam.py:80-111): Adam moments + clip-to-radius in two fused passes.
//
//   pass 1 (tradam_moments): m, v updated in place; s = m_hat / (sqrt(v_hat) + eps) is NOT stored,
//           only its length (max|s| for the inf-norm, s.s -- not its root, as coded in
//           tradam.py:101 -- for the 2-norm); the last CTA derives the scale lr/len or lr.
//   pass 2 (tradam_apply):   s is recomputed from m, v; step = scale * s; w -= step.
// 9 n-vector accesses instead of the reference's ~25 (12 ATen launches with temporaries).
// The element arithmetic is done in the buffers' own dtype with the reference's operation order
// and explicit round-to-nearest intrinsics (no contraction), so given equal inputs m, v, s and
// the step agree with the CPU reference to the last bit up to its own fused-multiply-add choices.
#include "stream.cuh"

namespace {

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }

// s = (m / bc1) / (sqrt(v / bc2) + eps), every operation rounded to T (tradam.py:91-96)
template <class T>
__device__ __forceinline__ T adam_dir(T m, T v, T bc1, T bc2, T eps) {
    return div_rn(div_rn(m, bc1), add_rn(sqrt_rn(div_rn(v, bc2)), eps));
}

template <class VT>
struct AdamMomentsOp {
    static constexpr int NRED = 2;
    __host__ __device__ static constexpr bool is_max(int j) { return j == 1; }
    dd4ml_state* st;
    const VT* g;
    VT* m;
    VT* v;
    VT b1, om1, b2, om2, bc1, bc2, eps;
    double lr;
    int norm_inf;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() {}
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) {
        Raw<V, VT> gr, mr, vr;
        ldr_ro<V, VT>(gr, g, i);
        ldr_rw<V, VT>(mr, m, i);
        ldr_rw<V, VT>(vr, v, i);
        double mo[V], vo[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const VT ge = gr.x[e];
            const VT mn = fma_rn(ge, om1, mul_rn(mr.x[e], b1));                     // m.mul_(b1).add_(g, alpha=1-b1)
            const VT vn = add_rn(mul_rn(vr.x[e], b2), mul_rn(mul_rn(om2, ge), ge));  // v.mul_(b2).addcmul_(g, g, value=1-b2)
            const double s = (double)adam_dir<VT>(mn, vn, bc1, bc2, eps);
            acc[0] += s * s;
            acc[1] = fmax(acc[1], fabs(s));
            mo[e] = (double)mn;
            vo[e] = (double)vn;
        }
        st_v<V, VT>(m, i, mo);
        st_v<V, VT>(v, i, vo);
    }
    __device__ void finalize(const double* s, dd4ml_state* t) {
        const double len = (double)(VT)(norm_inf ? s[1] : s[0]);   // .item() of a VT tensor
        t->pnorm[0] = len;
        t->aux[0] = (len > lr) ? lr / len : lr;                     // tradam.py:104-107
    }
};

template <class VT, class WT>
struct AdamApplyOp {
    static constexpr int NRED = 0;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    dd4ml_state* st;
    const VT* m;
    const VT* v;
    WT* w;
    VT* step;   // nullable
    VT bc1, bc2, eps;
    VT scale;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() { scale = (VT)st->aux[0]; }
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double*) {
        Raw<V, VT> mr, vr;
        Raw<V, WT> wr;
        ldr_ro<V, VT>(mr, m, i);
        ldr_ro<V, VT>(vr, v, i);
        ldr_rw<V, WT>(wr, w, i);
        double so[V], wo[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const VT sp = mul_rn(adam_dir<VT>(mr.x[e], vr.x[e], bc1, bc2, eps), scale);
            so[e] = (double)sp;
            wo[e] = (double)add_rn((WT)wr.x[e], -(WT)sp);     // p.add_(step * -1)
        }
        if (step != nullptr) st_v<V, VT>(step, i, so);
        st_v<V, WT>(w, i, wo);
    }
    __device__ void finalize(const double*, dd4ml_state*) {}
};

}  // namespace

extern "C" {

int dd4ml_tradam_moments(double* state, double* ws, const void* g, void* m, void* v, double beta1, double beta2,
                         double bc1, double bc2, double eps, double lr, int norm_inf, int64_t n, int vt,
                         void* stream) {
    if (!state || !ws || !g || !m || !v || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(g) && aligned16(m) && aligned16(v);
    if (vt == 0) {
        AdamMomentsOp<float> op{(dd4ml_state*)state, (const float*)g, (float*)m, (float*)v, (float)beta1,
                                (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)bc1, (float)bc2,
                                (float)eps, lr, norm_inf};
        return launch(op, n, vec, ws, 4, s);
    }
    if (vt == 1) {
        AdamMomentsOp<double> op{(dd4ml_state*)state, (const double*)g, (double*)m, (double*)v, beta1, 1.0 - beta1,
                                 beta2, 1.0 - beta2, bc1, bc2, eps, lr, norm_inf};
        return launch(op, n, vec, ws, 4, s);
    }
    return DD4ML_E_DTYPE;
}

int dd4ml_tradam_apply(double* state, double* ws, const void* m, const void* v, void* w, void* step, double bc1,
                       double bc2, double eps, int64_t n, int vt, int wt, void* stream) {
    if (!state || !m || !v || !w || n < 0) return DD4ML_E_ARG;
    (void)ws;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(m) && aligned16(v) && aligned16(w) && (!step || aligned16(step));
    if (vt == 0 && wt == 0) {
        AdamApplyOp<float, float> op{(dd4ml_state*)state, (const float*)m, (const float*)v, (float*)w, (float*)step,
                                     (float)bc1, (float)bc2, (float)eps, 0.f};
        return launch(op, n, vec, nullptr, 4, s);
    }
    if (vt == 0 && wt == 1) {
        AdamApplyOp<float, double> op{(dd4ml_state*)state, (const float*)m, (const float*)v, (double*)w, (float*)step,
                                      (float)bc1, (float)bc2, (float)eps, 0.f};
        return launch(op, n, vec, nullptr, 4, s);
    }
    if (vt == 1 && wt == 1) {
        AdamApplyOp<double, double> op{(dd4ml_state*)state, (const double*)m, (const double*)v, (double*)w,
                                       (double*)step, bc1, bc2, eps, 0.0};
        return launch(op, n, vec, nullptr, 4, s);
    }
    return DD4ML_E_DTYPE;
}

}  // extern "C"
