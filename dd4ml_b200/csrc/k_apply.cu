// tr_apply: p = cg g + sum_j (cS_j S_j + cY_j Y_j); optional w += p; max|p|.
#include "tma.cuh"

namespace {

template <class HT, class VT, class WT, int KMAX>
struct ApplyOp {
    static constexpr int NRED = 1;
    static constexpr int F = 2, NS_MAX = F + 2 * KMAX, TMA_CW = 8, TMA_PW = 4, TMA_V = 4, TMA_SPLIT = 1;
    __host__ __device__ static constexpr bool is_max(int) { return true; }
    dd4ml_state* st;
    const VT* g;
    const HT* S;
    const HT* Y;
    VT* p;
    WT* w;
    int64_t ld;
    Slots<KMAX> sl;
    double cg, cS[KMAX > 0 ? KMAX : 1], cY[KMAX > 0 ? KMAX : 1];
    bool conv;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() {
        sl.load(st, true);
        cg = st->cg[0];
        conv = st->converged[0] != 0.0;
        unsigned m = 0;  // stream only slots with non-zero coefficients (first-order step: none)
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            cS[j] = cY[j] = 0.0;
            if (j < sl.k) {
                cS[j] = st->cS[sl.slot[j]];
                cY[j] = st->cY[sl.slot[j]];
                if (cS[j] != 0.0 || cY[j] != 0.0) m |= 1u << j;
            }
        }
        sl.mask = m;
    }
    __device__ bool active() const { return true; }
    __device__ bool use_tma() const { return !conv; }
    __device__ void stream_desc(int s, const void*& p, int& esz) const {
        if (s == 0) { p = g; esz = (int)sizeof(VT); }
        else if (s == 1) { p = w; esz = (int)sizeof(WT); }
        else PairRegs<4, HT, KMAX>::desc(s - F, S, Y, ld, sl, p, esz);
    }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) { GlobalLoader l; body_l<V>(l, i, acc); }
    template <int V, class L>
    __device__ __forceinline__ void body_l(const L& l, int64_t i, double* acc) {
        double pv[V];
        if (conv) {
#pragma unroll
            for (int e = 0; e < V; ++e) pv[e] = 0.0;
            st_v<V, VT>(p, i, pv);
            return;
        }
        Raw<V, VT> gr;
        Raw<V, WT> wr;
        PairRegs<V, HT, KMAX> pr;
        l.template get<V, VT>(0, g, i, false, gr);
        if (w != nullptr) l.template get<V, WT>(1, w, i, true, wr);
        pr.load(l, F, S, Y, ld, sl, i);
#pragma unroll
        for (int e = 0; e < V; ++e) pv[e] = cg * (double)gr.x[e];
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (sl.live(j)) {
#pragma unroll
                for (int e = 0; e < V; ++e) pv[e] += cS[j] * (double)pr.s[j].x[e] + cY[j] * (double)pr.y[j].x[e];
            }
#pragma unroll
        for (int e = 0; e < V; ++e) { pv[e] = rnd<VT>(pv[e]); acc[0] = fmax(acc[0], fabs(pv[e])); }
        st_v<V, VT>(p, i, pv);
        if (w != nullptr) {
            double wv[V];
#pragma unroll
            for (int e = 0; e < V; ++e) wv[e] = (double)wr.x[e] + pv[e];
            st_v<V, WT>(w, i, wv);
        }
    }
    __device__ void finalize(const double* s, dd4ml_state* t) { t->pmax[0] = s[0]; }
};

}  // namespace

namespace {
template <class HT, class VT, class WT>
int apply_t(double* state, double* ws, const void* g, const void* S, const void* Y, void* p, void* w, int64_t n,
            int64_t ld, int kcap, bool vec, cudaStream_t s) {
    constexpr bool FAST = sizeof(HT) == 4 && sizeof(VT) == 4 && sizeof(WT) == 4;
    if (!S) {
        ApplyOp<HT, VT, WT, 0> op{(dd4ml_state*)state, (const VT*)g, nullptr, nullptr, (VT*)p, (WT*)w, ld};
        return launch(op, n, vec, ws, 4, s);
    }
    auto run = [&](auto kc) {
        constexpr int K = decltype(kc)::value;
        ApplyOp<HT, VT, WT, K> op{(dd4ml_state*)state, (const VT*)g, (const HT*)S, (const HT*)Y, (VT*)p, (WT*)w, ld};
        return launch_sel<FAST>(op, n, vec, ws, 3, TILE * (int)(sizeof(VT) + sizeof(WT) + 2 * K * sizeof(HT)), true, s);
    };
    DD4ML_RUN_K(FAST, kcap, run)
}
}  // namespace

extern "C" int dd4ml_tr_apply(double* state, double* ws, const void* g, const void* S, const void* Y, void* p,
                              void* w, int64_t n, int64_t ld, int vt, int ht, int wt, int kcap, void* stream) {
    if (!state || !ws || !g || !p || n < 0 || kcap > DD4ML_MAXM) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(g) && aligned16(p) && (!w || aligned16(w)) &&
                     (!S || (aligned16(S) && aligned16(Y) && ld % 4 == 0));
#define CALL(HT, VT, WT) return apply_t<HT, VT, WT>(state, ws, g, S, Y, p, w, n, ld, kcap, vec, s);
    DT_SWITCH3(ht, vt, wt, CALL)
#undef CALL
}

extern "C" int dd4ml_tr_propose(double*, double*, const void*, const void*, const void*, int64_t, int64_t, int,
                                int, int, int, void*);
extern "C" int dd4ml_lsr1_B(double* state, double* ws, const void* v, const void* S, const void* Y, void* out,
                            int64_t n, int64_t ld, int vt, int ht, int kcap, void* stream) {
    if (!S || !Y) return DD4ML_E_ARG;
    int rc = dd4ml_tr_propose(state, ws, v, S, Y, n, ld, vt, ht, 2, kcap, stream);
    if (rc) return rc;
    return dd4ml_tr_apply(state, ws, v, S, Y, out, nullptr, n, ld, vt, ht, vt, kcap, stream);
}
