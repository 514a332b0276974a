// lssr1_begin: first pass of LSSR1_TR.step (lssr1_tr.py:493-587).
#include "tma.cuh"

namespace {

template <class HT, class VT, class WT, int KMAX>
struct LssrBeginOp {
    static constexpr int NRED = 7 + 6 * KMAX;
    // KMAX = 10: 24 streams x 4 KB would leave two stages in the ring.  Two consumer groups of four warps take
    // alternate 2 KB-per-stream tiles instead: four stages, and a thread still owns 4 consecutive elements
    // (67 accumulators: one float32 -> float64 conversion per accumulator per 4 elements, see heavy.cuh)
    static constexpr int F = 4, NS_MAX = F + 2 * KMAX, TMA_CW = 8, TMA_PW = 4, TMA_V = 4, TMA_SPLIT = KMAX > 6 ? 2 : 1;
    __host__ __device__ static constexpr bool is_max(int j) { return j == 1; }
    dd4ml_state* st;
    const WT* w;
    const WT* old_w;      // previous iterate / gradient: read ...
    WT* old_w_out;        // ... and written to a second buffer: the reference keeps old_wk / prev_grad when
    const VT* g;          // ||g|| <= tol (it returns before storing them, lssr1_tr.py:509-526), which is only
    const VT* prev_g;     // known after this pass -- the host swaps the buffers unless the step converged
    VT* prev_g_out;
    HT* S;
    HT* Y;
    int64_t ld;
    int has_prev;
    const void* lossp;
    int lt;
    double rad[3];   // delta, min_delta, max_delta pushed with this launch (rad[0] < 0: keep the state's)
    Slots<KMAX> sl;
    int spare;
    bool pair;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() {
        sl.load(st, st->second_order[0] != 0.0);
        spare = (int)st->spare[0];
        pair = has_prev && st->second_order[0] != 0.0;
    }
    __device__ bool active() const { return true; }
    __device__ bool use_tma() const { return true; }
    __device__ void stream_desc(int s, const void*& q, int& esz) const {
        if (s == 0) { q = w; esz = (int)sizeof(WT); }
        else if (s == 1) { q = g; esz = (int)sizeof(VT); }
        else if (s == 2) { q = pair ? (const void*)old_w : nullptr; esz = (int)sizeof(WT); }
        else if (s == 3) { q = pair ? (const void*)prev_g : nullptr; esz = (int)sizeof(VT); }
        else PairRegs<4, HT, KMAX>::desc(s - F, S, Y, ld, sl, q, esz);
    }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) { GlobalLoader l; body_l<V>(l, i, acc); }
    template <int V, class L>
    __device__ __forceinline__ void body_l(const L& l, int64_t i, double* acc) {
        Raw<V, WT> wr, owr;
        Raw<V, VT> gr, pgr;
        PairRegs<V, HT, KMAX> pr;
        constexpr bool kF32Tiles = KMAX > 6 && BothF32<HT, VT>::value && BothF32<WT, WT>::value && V > 1;
        constexpr bool kJustInTime = kF32Tiles && L::kStaged;   // shared-memory operands: load pair j when it is used
        l.template get<V, WT>(0, w, i, false, wr);
        l.template get<V, VT>(1, g, i, false, gr);
        if (pair) {
            l.template get<V, WT>(2, old_w, i, true, owr);
            l.template get<V, VT>(3, prev_g, i, true, pgr);
        }
        if constexpr (!kJustInTime) pr.load(l, F, S, Y, ld, sl, i);
        double wv[V], gv[V], sv[V], yv[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            wv[e] = (double)wr.x[e];
            gv[e] = (double)gr.x[e];
            acc[0] += gv[e] * gv[e];
            acc[1] = fmax(acc[1], fabs(gv[e]));
        }
        if (pair) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                // differences in the vectors' own dtype (what the reference computes), then the
                // L-SR1 dtype: one conversion per element instead of three
                sv[e] = (double)(HT)(WT)(wr.x[e] - owr.x[e]);
                yv[e] = (double)(HT)(VT)(gr.x[e] - pgr.x[e]);
                acc[2] += sv[e] * sv[e];
                acc[3] += sv[e] * yv[e];
                acc[4] += yv[e] * yv[e];
                acc[5] += sv[e] * gv[e];
                acc[6] += yv[e] * gv[e];
            }
        }
        if constexpr (kF32Tiles) {
            float sf[V], yf[V];
#pragma unroll
            for (int e = 0; e < V; ++e) { sf[e] = pair ? (float)sv[e] : 0.f; yf[e] = pair ? (float)yv[e] : 0.f; }   // exact: sv, yv are float32 values
#pragma unroll
            for (int j = 0; j < KMAX; ++j)
                if (sl.live(j)) {
                    if constexpr (kJustInTime) {
                        l.template get<V, HT>(F + j, S, i, false, pr.s[j]);
                        l.template get<V, HT>(F + KMAX + j, Y, i, false, pr.y[j]);
                    }
                    acc[7 + j] += (double)dot32<V>(pr.s[j].x, gr.x);
                    acc[7 + KMAX + j] += (double)dot32<V>(pr.y[j].x, gr.x);
                    if (pair) {
                        acc[7 + 2 * KMAX + j] += (double)dot32<V>(pr.s[j].x, sf);
                        acc[7 + 3 * KMAX + j] += (double)dot32<V>(pr.y[j].x, sf);
                        acc[7 + 4 * KMAX + j] += (double)dot32<V>(pr.s[j].x, yf);
                        acc[7 + 5 * KMAX + j] += (double)dot32<V>(pr.y[j].x, yf);
                    }
                }
        } else {
#pragma unroll
            for (int j = 0; j < KMAX; ++j)
                if (sl.live(j)) {
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const double a = (double)pr.s[j].x[e], b = (double)pr.y[j].x[e];
                        acc[7 + j] += a * gv[e];
                        acc[7 + KMAX + j] += b * gv[e];
                        if (pair) {
                            acc[7 + 2 * KMAX + j] += a * sv[e];
                            acc[7 + 3 * KMAX + j] += b * sv[e];
                            acc[7 + 4 * KMAX + j] += a * yv[e];
                            acc[7 + 5 * KMAX + j] += b * yv[e];
                        }
                    }
                }
        }
        if (pair) {
            st_v<V, HT>(S + (int64_t)spare * ld, i, sv);
            st_v<V, HT>(Y + (int64_t)spare * ld, i, yv);
        }
        st_v<V, WT>(old_w_out, i, wv);
        st_v<V, VT>(prev_g_out, i, gv);
    }
    __device__ void finalize(const double* s, dd4ml_state* t) {   // chain + solve: k-space kernel (KT_BEGIN)
        if (rad[0] >= 0.0) { t->delta[0] = rad[0]; t->min_delta[0] = rad[1]; t->max_delta[0] = rad[2]; }
        t->gg[0] = s[0];
        t->ginf[0] = s[1];
        t->loss[0] = ld_scalar(lossp, lt);
        for (int j = 0; j < sl.k; ++j) { t->Sg[sl.at(j)] = s[7 + j]; t->Yg[sl.at(j)] = s[7 + KMAX + j]; }
        t->pair_tried[0] = 0.0;
        t->pair_ok[0] = 0.0;
        t->err[0] = 0.0;
        if (pair) {
            t->c_ss[0] = s[2]; t->c_sy[0] = s[3]; t->c_yy[0] = s[4];
            t->c_sg[0] = s[5]; t->c_yg[0] = s[6];
            for (int j = 0; j < sl.k; ++j) {
                t->c_Ss[sl.at(j)] = s[7 + 2 * KMAX + j];
                t->c_Ys[sl.at(j)] = s[7 + 3 * KMAX + j];
                t->c_Sy[sl.at(j)] = s[7 + 4 * KMAX + j];
                t->c_Yy[sl.at(j)] = s[7 + 5 * KMAX + j];
            }
        }
    }
};

}  // namespace

namespace {
template <class HT, class VT, class WT>
int begin_t(double* state, double* ws, const void* w, const void* old_w, void* old_w_out, const void* g,
            const void* prev_g, void* prev_g_out, int has_prev, void* S, void* Y, const void* loss, int lt, int64_t n,
            int64_t ld, int kcap, bool vec, double delta, double min_delta, double max_delta, cudaStream_t s) {
    constexpr bool FAST = sizeof(HT) == 4 && sizeof(VT) == 4 && sizeof(WT) == 4;
    auto run = [&](auto kc) {
        constexpr int K = decltype(kc)::value;
        LssrBeginOp<HT, VT, WT, K> op{(dd4ml_state*)state, (const WT*)w, (const WT*)old_w, (WT*)old_w_out,
                                      (const VT*)g, (const VT*)prev_g, (VT*)prev_g_out,
                                      (HT*)S, (HT*)Y, ld, has_prev, loss, lt, {delta, min_delta, max_delta}};
        return launch_sel<FAST>(op, n, vec, ws, K <= 4 ? 2 : 1,
                                TILE * (int)(2 * sizeof(WT) + 2 * sizeof(VT) + 2 * K * sizeof(HT)), true, s);
    };
    DD4ML_RUN_K(FAST, kcap, run)
}
}  // namespace

extern "C" int dd4ml_lssr1_begin(double* state, double* ws, const void* w, const void* old_w, void* old_w_out,
                                 const void* g, const void* prev_g, void* prev_g_out, int has_prev, void* S,
                                 void* Y, const void* loss, int lt, int64_t n, int64_t ld, int vt, int ht, int wt,
                                 int kcap, double delta, double min_delta, double max_delta, void* stream) {
    if (!state || !ws || !w || !old_w || !old_w_out || !g || !prev_g || !prev_g_out || !S || !Y || !loss || n < 0 ||
        kcap > DD4ML_MAXM)
        return DD4ML_E_ARG;
    if (lt != 0 && lt != 1) return DD4ML_E_DTYPE;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(w) && aligned16(old_w) && aligned16(old_w_out) && aligned16(g) && aligned16(prev_g) &&
                     aligned16(prev_g_out) && aligned16(S) && aligned16(Y) && ld % 4 == 0;
    int rc = DD4ML_E_DTYPE;
#define CALL(HT, VT, WT) \
    rc = begin_t<HT, VT, WT>(state, ws, w, old_w, old_w_out, g, prev_g, prev_g_out, has_prev, S, Y, loss, lt, n, ld, \
                             kcap, vec, delta, min_delta, max_delta, s);
    DT_SWITCH3_RC(ht, vt, wt, CALL)
#undef CALL
    if (rc) return rc;
    KTask t{};
    t.kind = KT_BEGIN;
    t.ht = ht; t.vt = vt; t.lt = lt;
    t.has_prev = has_prev;
    return dd4ml_ks_launch(state, t, s, kcap);
}
