// dd4ml_b200 kernels -- see stream.cuh for the shared streaming-kernel template.
#include "stream.cuh"

namespace {

// first pass of LSSR1_TR.step
template <class HT, class VT, class WT, int KMAX>
struct LssrBeginOp {
    static constexpr int NRED = 7 + 6 * KMAX;
    __host__ __device__ static constexpr bool is_max(int j) { return j == 1; }
    dd4ml_state* st;
    const WT* w;
    WT* old_w;
    const VT* g;
    VT* prev_g;
    HT* S;
    HT* Y;
    int64_t ld;
    int has_prev;
    const void* lossp;
    int lt;
    Slots<KMAX> sl;
    int spare;
    bool pair;
    __device__ void setup() {
        sl.load(st, st->second_order[0] != 0.0);
        spare = (int)st->spare[0];
        pair = has_prev && st->second_order[0] != 0.0;
    }
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) {
        double wv[V], gv[V];
        ld_ro<V, WT>(w, i, wv);
        ld_ro<V, VT>(g, i, gv);
#pragma unroll
        for (int e = 0; e < V; ++e) { acc[0] += gv[e] * gv[e]; acc[1] = fmax(acc[1], fabs(gv[e])); }
        double sv[V], yv[V];
        if (pair) {
            double ow[V], pg[V];
            ld_rw<V, WT>(old_w, i, ow);
            ld_rw<V, VT>(prev_g, i, pg);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                sv[e] = rnd<HT>(rnd<WT>(wv[e] - ow[e]));
                yv[e] = rnd<HT>(rnd<VT>(gv[e] - pg[e]));
                acc[2] += sv[e] * sv[e];
                acc[3] += sv[e] * yv[e];
                acc[4] += yv[e] * yv[e];
                acc[5] += sv[e] * gv[e];
                acc[6] += yv[e] * gv[e];
            }
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < sl.k) {
                double a[V], b[V];
                ld_ro<V, HT>(S + (int64_t)sl.slot[j] * ld, i, a);
                ld_ro<V, HT>(Y + (int64_t)sl.slot[j] * ld, i, b);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    acc[7 + j] += a[e] * gv[e];
                    acc[7 + KMAX + j] += b[e] * gv[e];
                }
                if (pair) {
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        acc[7 + 2 * KMAX + j] += a[e] * sv[e];
                        acc[7 + 3 * KMAX + j] += b[e] * sv[e];
                        acc[7 + 4 * KMAX + j] += a[e] * yv[e];
                        acc[7 + 5 * KMAX + j] += b[e] * yv[e];
                    }
                }
            }
        if (pair) {
            st_v<V, HT>(S + (int64_t)spare * ld, i, sv);
            st_v<V, HT>(Y + (int64_t)spare * ld, i, yv);
        }
        st_v<V, WT>(old_w, i, wv);
        st_v<V, VT>(prev_g, i, gv);
    }
    __device__ void finalize(const double* s, ks::Scratch* scr) {
        st->gg[0] = s[0];
        st->ginf[0] = s[1];
        st->loss[0] = ld_scalar(lossp, lt);
        for (int j = 0; j < sl.k; ++j) { st->Sg[sl.slot[j]] = s[7 + j]; st->Yg[sl.slot[j]] = s[7 + KMAX + j]; }
        st->pair_tried[0] = 0.0;
        st->pair_ok[0] = 0.0;
        st->err[0] = 0.0;
        // lssr1_tr.py:509: the convergence test precedes the memory update
        const double gn = (double)(VT)sqrt(s[0]);
        if (pair && gn > st->tol[0]) {
            st->c_ss[0] = s[2]; st->c_sy[0] = s[3]; st->c_yy[0] = s[4];
            st->c_sg[0] = s[5]; st->c_yg[0] = s[6];
            for (int j = 0; j < sl.k; ++j) {
                st->c_Ss[sl.slot[j]] = s[7 + 2 * KMAX + j];
                st->c_Ys[sl.slot[j]] = s[7 + 3 * KMAX + j];
                st->c_Sy[sl.slot[j]] = s[7 + 4 * KMAX + j];
                st->c_Yy[sl.slot[j]] = s[7 + 5 * KMAX + j];
            }
            const double tol = st->tol[0];
            if (sqrt(s[2]) > tol && sqrt(s[4]) > tol) {  // lssr1_tr.py:521
                st->pair_tried[0] = 1.0;
                if (ks::lsr1_try_update(st, scr)) { st->Sg[spare] = s[5]; st->Yg[spare] = s[6]; }
            }
        }
        const double ninf = st->norm_inf[0];
        st->norm_inf[0] = 0.0;  // LSSR1_TR forces the 2-norm (lssr1_tr.py:94)
        if (st->second_order[0] != 0.0 && st->k[0] > 0.0) ks::tr_solve<HT>(st, 1, scr);
        else ks::tr_solve<VT>(st, 1, scr);
        st->norm_inf[0] = ninf;
    }
};

template <class VT, class WT>
struct TrialOp {  // w = base + alpha p
    static constexpr int NRED = 0;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    WT* w;
    const WT* base;
    const VT* p;
    double alpha;
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
    __device__ void finalize(const double*, ks::Scratch*) {}
};

template <class VT>
struct EvalOp {  // deriv = g . p ; ||g|| ; g_save = g
    static constexpr int NRED = 3;
    __host__ __device__ static constexpr bool is_max(int j) { return j == 2; }
    dd4ml_state* st;
    const VT* g;
    const VT* p;
    VT* gs;
    const void* loss;
    int lt;
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
    __device__ void finalize(const double* s, ks::Scratch*) {
        st->deriv[0] = s[0];
        st->out_gg[0] = s[1];
        st->out_ginf[0] = s[2];
        st->eval_loss[0] = ld_scalar(loss, lt);
    }
};

}  // namespace

extern "C" {

int dd4ml_lssr1_begin(double* state, double* ws, const void* w, void* old_w, const void* g,
                      void* prev_g, int has_prev, void* S, void* Y, const void* loss, int lt,
                      int64_t n, int64_t ld, int vt, int ht, int wt, void* stream) {
    if (!state || !ws || !w || !old_w || !g || !prev_g || !S || !Y || !loss || n < 0) return DD4ML_E_ARG;
    if (lt != 0 && lt != 1) return DD4ML_E_DTYPE;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(w) && aligned16(old_w) && aligned16(g) && aligned16(prev_g) && aligned16(S) &&
                     aligned16(Y) && ld % 4 == 0;
#define CALL(HT, VT, WT)                                                                          \
    {                                                                                             \
        LssrBeginOp<HT, VT, WT, 8> op{(dd4ml_state*)state, (const WT*)w, (WT*)old_w, (const VT*)g, \
            (VT*)prev_g, (HT*)S, (HT*)Y, ld, has_prev, loss, lt};                                           \
        return launch(op, n, vec, ws, 1, s);                                                      \
    }
    DT_SWITCH3(ht, vt, wt, CALL)
#undef CALL
}

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
