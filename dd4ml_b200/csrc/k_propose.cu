// tr_propose: ||g||, S^T g, Y^T g in one pass + k-space sub-problem solve in the k-space kernel launched behind it.
#include "tma.cuh"

namespace {

template <class HT, class VT, int KMAX>
struct ProposeOp {
    static constexpr int NRED = 2 + 2 * KMAX;
    static constexpr int F = 1, NS_MAX = F + 2 * KMAX, TMA_CW = 8, TMA_PW = 4, TMA_V = 4, TMA_SPLIT = 1;
    __host__ __device__ static constexpr bool is_max(int j) { return j == 1; }
    dd4ml_state* st;
    const VT* g;
    const HT* S;
    const HT* Y;
    int64_t ld;
    int mode;
    Slots<KMAX> sl;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() { sl.load(st, mode >= 2 || st->second_order[0] != 0.0); }
    __device__ bool active() const { return true; }
    __device__ bool use_tma() const { return true; }
    __device__ void stream_desc(int s, const void*& p, int& esz) const {
        if (s == 0) { p = g; esz = (int)sizeof(VT); }
        else PairRegs<4, HT, KMAX>::desc(s - F, S, Y, ld, sl, p, esz);
    }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) { GlobalLoader l; body_l<V>(l, i, acc); }
    template <int V, class L>
    __device__ __forceinline__ void body_l(const L& l, int64_t i, double* acc) {
        Raw<V, VT> gr;
        PairRegs<V, HT, KMAX> pr;
        l.template get<V, VT>(0, g, i, false, gr);
        pr.load(l, F, S, Y, ld, sl, i);
        double gv[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            gv[e] = (double)gr.x[e];
            acc[0] += gv[e] * gv[e];
            acc[1] = fmax(acc[1], fabs(gv[e]));
        }
        if constexpr (KMAX > 6 && BothF32<HT, VT>::value && V > 1) {
#pragma unroll
            for (int j = 0; j < KMAX; ++j)
                if (sl.live(j)) {
                    acc[2 + j] += (double)dot32<V>(pr.s[j].x, gr.x);
                    acc[2 + KMAX + j] += (double)dot32<V>(pr.y[j].x, gr.x);
                }
        } else {
#pragma unroll
            for (int j = 0; j < KMAX; ++j)
                if (sl.live(j)) {
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        acc[2 + j] += (double)pr.s[j].x[e] * gv[e];
                        acc[2 + KMAX + j] += (double)pr.y[j].x[e] * gv[e];
                    }
                }
        }
    }
    __device__ void finalize(const double* s, dd4ml_state* t) {   // the solve itself: k-space kernel (KT_SOLVE / KT_B / KT_SMW)
        t->gg[0] = s[0];
        t->ginf[0] = s[1];
        for (int j = 0; j < sl.k; ++j) { t->Sg[sl.at(j)] = s[2 + j]; t->Yg[sl.at(j)] = s[2 + KMAX + j]; }
    }
};

}  // namespace

namespace {
template <class HT, class VT>
int propose_t(double* state, double* ws, const void* g, const void* S, const void* Y, int64_t n, int64_t ld, int mode,
              int kcap, bool vec, cudaStream_t s) {
    constexpr bool FAST = sizeof(HT) == 4 && sizeof(VT) == 4;
    if (!S) {
        ProposeOp<HT, VT, 0> op{(dd4ml_state*)state, (const VT*)g, nullptr, nullptr, ld, mode};
        return launch(op, n, vec, ws, 4, s);
    }
    auto run = [&](auto kc) {
        constexpr int K = decltype(kc)::value;
        ProposeOp<HT, VT, K> op{(dd4ml_state*)state, (const VT*)g, (const HT*)S, (const HT*)Y, ld, mode};
        return launch_sel<FAST>(op, n, vec, ws, 3, TILE * (int)(sizeof(VT) + 2 * K * sizeof(HT)), true, s);   // measured: TMA wins for every K
    };
    DD4ML_RUN_K(FAST, kcap, run)
}
}  // namespace

extern "C" int dd4ml_tr_propose(double* state, double* ws, const void* g, const void* S, const void* Y,
                                int64_t n, int64_t ld, int vt, int ht, int mode, int kcap, void* stream) {
    if (!state || !ws || !g || n < 0 || kcap > DD4ML_MAXM) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(g) && (!S || (aligned16(S) && aligned16(Y) && ld % 4 == 0));
    int rc = DD4ML_E_DTYPE;
#define CALL(HT, VT) rc = propose_t<HT, VT>(state, ws, g, S, Y, n, ld, mode, kcap, vec, s);
    DT_SWITCH2_RC(ht, vt, CALL)
#undef CALL
    if (rc) return rc;
    KTask t{};
    t.kind = mode == 2 ? KT_B : (mode == 3 ? KT_SMW : KT_SOLVE);
    t.mode = mode;
    t.ht = ht; t.vt = vt;
    return dd4ml_ks_launch(state, t, s, kcap);
}
