// lssr1_trial (w = w_k + alpha p) and lssr1_eval (g.p, ||g||, keep g): per-evaluation kernels of
// the strong-Wolfe line search (lssr1_tr.py:247-278).
#include "stream.cuh"

namespace {

template <class VT, class WT>
struct TrialOp {
    static constexpr int NRED = 0;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    WT* w;
    const WT* base;
    const VT* p;
    double alpha;
    __device__ dd4ml_state* state() const { return nullptr; }
    __device__ void setup() {}
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double*) {
        double b[V], pv[V];
        ld_ro<V, WT>(base, i, b);
        ld_ro<V, VT>(p, i, pv);
#pragma unroll
        for (int e = 0; e < V; ++e) b[e] += rnd<VT>(alpha * pv[e]);
        st_v<V, WT>(w, i, b);
    }
    __device__ void finalize(const double*, dd4ml_state*) {}
};

template <class VT>
struct EvalOp {
    static constexpr int NRED = 3;
    __host__ __device__ static constexpr bool is_max(int j) { return j == 2; }
    dd4ml_state* st;
    const VT* g;
    const VT* p;
    VT* gs;
    const void* loss;
    int lt;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() {}
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) {
        double gv[V], pv[V];
        ld_ro<V, VT>(g, i, gv);
        ld_ro<V, VT>(p, i, pv);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            acc[0] += gv[e] * pv[e];
            acc[1] += gv[e] * gv[e];
            acc[2] = fmax(acc[2], fabs(gv[e]));
        }
        if (gs != nullptr) st_v<V, VT>(gs, i, gv);
    }
    __device__ void finalize(const double* s, dd4ml_state* t) {
        t->deriv[0] = s[0];
        t->out_gg[0] = s[1];
        t->out_ginf[0] = s[2];
        t->eval_loss[0] = ld_scalar(loss, lt);
    }
};

}  // namespace

extern "C" {

int dd4ml_lssr1_trial(void* w, const void* base, const void* p, double alpha, int64_t n, int vt,
                      int wt, void* stream) {
    if (!w || !base || !p || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(w) && aligned16(base) && aligned16(p);
    if (vt == 0 && wt == 0) { TrialOp<float, float> op{(float*)w, (const float*)base, (const float*)p, alpha}; return launch(op, n, vec, nullptr, 4, s); }
    if (vt == 0 && wt == 1) { TrialOp<float, double> op{(double*)w, (const double*)base, (const float*)p, alpha}; return launch(op, n, vec, nullptr, 4, s); }
    if (vt == 1 && wt == 1) { TrialOp<double, double> op{(double*)w, (const double*)base, (const double*)p, alpha}; return launch(op, n, vec, nullptr, 4, s); }
    return DD4ML_E_DTYPE;
}

int dd4ml_lssr1_eval(double* state, double* ws, const void* g, const void* p, void* g_save,
                     const void* loss, int lt, int64_t n, int vt, void* stream) {
    if (!state || !ws || !g || !p || !loss || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(g) && aligned16(p) && (!g_save || aligned16(g_save));
    if (vt == 0) { EvalOp<float> op{(dd4ml_state*)state, (const float*)g, (const float*)p, (float*)g_save, loss, lt}; return launch(op, n, vec, ws, 4, s); }
    if (vt == 1) { EvalOp<double> op{(dd4ml_state*)state, (const double*)g, (const double*)p, (double*)g_save, loss, lt}; return launch(op, n, vec, ws, 4, s); }
    return DD4ML_E_DTYPE;
}

}  // extern "C"
