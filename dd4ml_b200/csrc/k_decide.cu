// tr_decide (after the trial evaluation of TR.step) and the standalone LSR1.update_memory.
#include "tma.cuh"

namespace {

template <class HT, class VT, class WT, int KMAX, bool MEM>
struct DecideOp {
    static constexpr int NRED = 5 + 4 * KMAX;
    static constexpr int F = 3, NS_MAX = F + 2 * KMAX, TMA_CW = 8, TMA_PW = 4, TMA_V = 4, TMA_SPLIT = 1;
    __host__ __device__ static constexpr bool is_max(int j) { return j == 4; }
    dd4ml_state* st;
    const void* loss;
    const void* trial;
    int lt;
    const VT* gt;
    const VT* go;
    VT* gout;
    const VT* p;
    WT* w;
    HT* S;
    HT* Y;
    int64_t ld;
    void* loss_out;      // nullable: device-side select of the returned loss (sync-free TR)
    int copy_old;        // rejected/converged: g_out <- g_old (only needed when they are different buffers)
    Slots<KMAX> sl;
    bool accept, conv;
    double rho, f0, f1;
    int spare;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() {
        sl.load(st, MEM);
        spare = (int)st->spare[0];
        f0 = ld_scalar(loss, lt);
        f1 = ld_scalar(trial, lt);
        conv = st->converged[0] != 0.0;
        if (conv) { accept = false; rho = 0.0; return; }
        double r = 0.0;   // (not &rho: the op must stay in registers)
        accept = (lt == 0) ? ks::tr_accept<float>(f0, f1, st->pred[0], st->tol[0], st->nu_dec[0], &r)
                           : ks::tr_accept<double>(f0, f1, st->pred[0], st->tol[0], st->nu_dec[0], &r);
        rho = r;
    }
    __device__ bool active() const { return !conv || copy_old; }
    __device__ bool use_tma() const { return accept; }
    __device__ void stream_desc(int s, const void*& q, int& esz) const {
        esz = (int)sizeof(VT);
        if (s == 0) q = p;
        else if (s == 1) q = gt;
        else if (s == 2) q = go;
        else if (MEM) PairRegs<4, HT, KMAX>::desc(s - F, S, Y, ld, sl, q, esz);
        else q = nullptr;
    }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) { GlobalLoader l; body_l<V>(l, i, acc); }
    template <int V, class L>
    __device__ __forceinline__ void body_l(const L& l, int64_t i, double* acc) {
        Raw<V, VT> prw;
        if (!accept) {  // tr.py:216 undo (always register staged)
            if (!conv) {
                Raw<V, WT> wr;
                ldr_ro<V, VT>(prw, p, i);
                ldr_rw<V, WT>(wr, w, i);
                double wv[V];
#pragma unroll
                for (int e = 0; e < V; ++e) wv[e] = (double)wr.x[e] - (double)prw.x[e];
                st_v<V, WT>(w, i, wv);
            }
            if (copy_old) {
                double ov[V];
                ld_ro<V, VT>(go, i, ov);
                st_v<V, VT>(gout, i, ov);
            }
            return;
        }
        Raw<V, VT> tr, orw;
        PairRegs<V, HT, KMAX> pr;
        l.template get<V, VT>(0, p, i, false, prw);
        l.template get<V, VT>(1, gt, i, false, tr);
        l.template get<V, VT>(2, go, i, true, orw);
        if (MEM) pr.load(l, F, S, Y, ld, sl, i);
        double tv[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            tv[e] = (double)tr.x[e];
            acc[3] += tv[e] * tv[e];
            acc[4] = fmax(acc[4], fabs(tv[e]));
        }
        if (MEM) {
            double sv[V], yv[V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                sv[e] = (double)(HT)prw.x[e];
                yv[e] = (double)(HT)(VT)(tr.x[e] - orw.x[e]);
                acc[0] += sv[e] * sv[e];
                acc[1] += sv[e] * yv[e];
                acc[2] += yv[e] * yv[e];
            }
            if constexpr (KMAX > 6 && BothF32<HT, VT>::value && V > 1) {
                float sf[V], yf[V];
#pragma unroll
                for (int e = 0; e < V; ++e) { sf[e] = (float)sv[e]; yf[e] = (float)yv[e]; }   // exact: float32 values
#pragma unroll
                for (int j = 0; j < KMAX; ++j)
                    if (sl.live(j)) {
                        acc[5 + j] += (double)dot32<V>(pr.s[j].x, sf);
                        acc[5 + KMAX + j] += (double)dot32<V>(pr.y[j].x, sf);
                        acc[5 + 2 * KMAX + j] += (double)dot32<V>(pr.s[j].x, yf);
                        acc[5 + 3 * KMAX + j] += (double)dot32<V>(pr.y[j].x, yf);
                    }
            } else {
#pragma unroll
                for (int j = 0; j < KMAX; ++j)
                    if (sl.live(j)) {
#pragma unroll
                        for (int e = 0; e < V; ++e) {
                            const double a = (double)pr.s[j].x[e], b = (double)pr.y[j].x[e];
                            acc[5 + j] += a * sv[e];
                            acc[5 + KMAX + j] += b * sv[e];
                            acc[5 + 2 * KMAX + j] += a * yv[e];
                            acc[5 + 3 * KMAX + j] += b * yv[e];
                        }
                    }
            }
            st_v<V, HT>(S + (int64_t)spare * ld, i, sv);
            st_v<V, HT>(Y + (int64_t)spare * ld, i, yv);
        }
        st_v<V, VT>(gout, i, tv);
    }
    __device__ void finalize(const double* s, dd4ml_state* t) {   // radius / L-SR1 update: k-space kernel (KT_DECIDE)
        t->loss[0] = f0;
        t->trial_loss[0] = f1;
        if (loss_out != nullptr) st_scalar(loss_out, lt, accept ? f1 : f0);
        if (conv) {  // tr.py:147: nothing happened
            t->accept[0] = 0.0;
            t->out_gn[0] = t->gn[0];
            return;
        }
        if (accept) {
            t->c_ss[0] = s[0]; t->c_sy[0] = s[1]; t->c_yy[0] = s[2];
            t->out_gg[0] = s[3]; t->out_ginf[0] = s[4];
            for (int j = 0; j < sl.k; ++j) {
                t->c_Ss[sl.at(j)] = s[5 + j];
                t->c_Ys[sl.at(j)] = s[5 + KMAX + j];
                t->c_Sy[sl.at(j)] = s[5 + 2 * KMAX + j];
                t->c_Yy[sl.at(j)] = s[5 + 3 * KMAX + j];
            }
        }
    }
};

// standalone LSR1.update_memory(s, y) (lsr1.py:53-144)
template <class HT, class VT, int KMAX>
struct UpdateOp {
    static constexpr int NRED = 3 + 4 * KMAX;
    static constexpr int F = 2, NS_MAX = F + 2 * KMAX, TMA_CW = 8, TMA_PW = 4, TMA_V = 4, TMA_SPLIT = 1;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    dd4ml_state* st;
    const VT* s_in;
    const VT* y_in;
    HT* S;
    HT* Y;
    int64_t ld;
    Slots<KMAX> sl;
    int spare;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() { sl.load(st, true); spare = (int)st->spare[0]; }
    __device__ bool active() const { return true; }
    __device__ bool use_tma() const { return true; }
    __device__ void stream_desc(int s, const void*& q, int& esz) const {
        esz = (int)sizeof(VT);
        if (s == 0) q = s_in;
        else if (s == 1) q = y_in;
        else PairRegs<4, HT, KMAX>::desc(s - F, S, Y, ld, sl, q, esz);
    }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) { GlobalLoader l; body_l<V>(l, i, acc); }
    template <int V, class L>
    __device__ __forceinline__ void body_l(const L& l, int64_t i, double* acc) {
        Raw<V, VT> sr, yr;
        PairRegs<V, HT, KMAX> pr;
        l.template get<V, VT>(0, s_in, i, false, sr);
        l.template get<V, VT>(1, y_in, i, false, yr);
        pr.load(l, F, S, Y, ld, sl, i);
        double sv[V], yv[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            sv[e] = rnd<HT>((double)sr.x[e]);
            yv[e] = rnd<HT>((double)yr.x[e]);
            acc[0] += sv[e] * sv[e];
            acc[1] += sv[e] * yv[e];
            acc[2] += yv[e] * yv[e];
        }
        if constexpr (KMAX > 6 && BothF32<HT, VT>::value && V > 1) {
            float sf[V], yf[V];
#pragma unroll
            for (int e = 0; e < V; ++e) { sf[e] = (float)sv[e]; yf[e] = (float)yv[e]; }       // exact: float32 values
#pragma unroll
            for (int j = 0; j < KMAX; ++j)
                if (sl.live(j)) {
                    acc[3 + j] += (double)dot32<V>(pr.s[j].x, sf);
                    acc[3 + KMAX + j] += (double)dot32<V>(pr.y[j].x, sf);
                    acc[3 + 2 * KMAX + j] += (double)dot32<V>(pr.s[j].x, yf);
                    acc[3 + 3 * KMAX + j] += (double)dot32<V>(pr.y[j].x, yf);
                }
        } else {
#pragma unroll
            for (int j = 0; j < KMAX; ++j)
                if (sl.live(j)) {
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const double a = (double)pr.s[j].x[e], b = (double)pr.y[j].x[e];
                        acc[3 + j] += a * sv[e];
                        acc[3 + KMAX + j] += b * sv[e];
                        acc[3 + 2 * KMAX + j] += a * yv[e];
                        acc[3 + 3 * KMAX + j] += b * yv[e];
                    }
                }
        }
        st_v<V, HT>(S + (int64_t)spare * ld, i, sv);
        st_v<V, HT>(Y + (int64_t)spare * ld, i, yv);
    }
    __device__ void finalize(const double* s, dd4ml_state* t) {   // acceptance chain: k-space kernel (KT_UPDATE)
        t->c_ss[0] = s[0]; t->c_sy[0] = s[1]; t->c_yy[0] = s[2];
        for (int j = 0; j < sl.k; ++j) {
            t->c_Ss[sl.at(j)] = s[3 + j];
            t->c_Ys[sl.at(j)] = s[3 + KMAX + j];
            t->c_Sy[sl.at(j)] = s[3 + 2 * KMAX + j];
            t->c_Yy[sl.at(j)] = s[3 + 3 * KMAX + j];
        }
    }
};

}  // namespace

namespace {
template <class HT, class VT, class WT>
int decide_t(double* state, double* ws, const void* loss, const void* trial_loss, int lt, const void* g_trial,
             const void* g_old, void* g_out, const void* p, void* w, void* S, void* Y, void* loss_out, int copy_old,
             int64_t n, int64_t ld, int kcap, bool vec, cudaStream_t s) {
    constexpr bool FAST = sizeof(HT) == 4 && sizeof(VT) == 4 && sizeof(WT) == 4;
    if (!S) {
        DecideOp<HT, VT, WT, 0, false> op{(dd4ml_state*)state, loss, trial_loss, lt, (const VT*)g_trial, (const VT*)g_old,
                                          (VT*)g_out, (const VT*)p, (WT*)w, nullptr, nullptr, ld, loss_out, copy_old};
        return launch(op, n, vec, ws, 4, s);
    }
    auto run = [&](auto kc) {
        constexpr int K = decltype(kc)::value;
        DecideOp<HT, VT, WT, K, true> op{(dd4ml_state*)state, loss, trial_loss, lt, (const VT*)g_trial, (const VT*)g_old,
                                         (VT*)g_out, (const VT*)p, (WT*)w, (HT*)S, (HT*)Y, ld, loss_out, copy_old};
        return launch_sel<FAST>(op, n, vec, ws, 2, TILE * (int)(3 * sizeof(VT) + 2 * K * sizeof(HT)), true, s);
    };
    DD4ML_RUN_K(FAST, kcap, run)
}

template <class HT, class VT>
int update_t(double* state, double* ws, const void* s_in, const void* y_in, void* S, void* Y, int64_t n, int64_t ld,
             int kcap, bool vec, cudaStream_t s) {
    constexpr bool FAST = sizeof(HT) == 4 && sizeof(VT) == 4;
    auto run = [&](auto kc) {
        constexpr int K = decltype(kc)::value;
        UpdateOp<HT, VT, K> op{(dd4ml_state*)state, (const VT*)s_in, (const VT*)y_in, (HT*)S, (HT*)Y, ld};
        return launch_sel<FAST>(op, n, vec, ws, 2, TILE * (int)(2 * sizeof(VT) + 2 * K * sizeof(HT)), true, s);
    };
    DD4ML_RUN_K(FAST, kcap, run)
}
}  // namespace

extern "C" {

int dd4ml_tr_decide(double* state, double* ws, const void* loss, const void* trial_loss, int lt,
                    const void* g_trial, const void* g_old, void* g_out, const void* p, void* w,
                    void* S, void* Y, void* loss_out, int64_t n, int64_t ld, int vt, int ht, int wt, int kcap,
                    void* stream) {
    if (!state || !ws || !loss || !trial_loss || !g_trial || !g_old || !g_out || !p || !w || n < 0 || kcap > DD4ML_MAXM)
        return DD4ML_E_ARG;
    if (lt != 0 && lt != 1) return DD4ML_E_DTYPE;
    cudaStream_t s = (cudaStream_t)stream;
    const int copy_old = (loss_out != nullptr && g_old != g_out) ? 1 : 0;   // device-side select mode only
    const bool vec = aligned16(g_trial) && aligned16(g_old) && aligned16(g_out) && aligned16(p) && aligned16(w) &&
                     (!S || (aligned16(S) && aligned16(Y) && ld % 4 == 0));
    int rc = DD4ML_E_DTYPE;
#define CALL(HT, VT, WT)                                                                                         \
    rc = decide_t<HT, VT, WT>(state, ws, loss, trial_loss, lt, g_trial, g_old, g_out, p, w, S, Y, loss_out, copy_old, \
                              n, ld, kcap, vec, s);
    DT_SWITCH3_RC(ht, vt, wt, CALL)
#undef CALL
    if (rc) return rc;
    KTask t{};
    t.kind = KT_DECIDE;
    t.ht = ht; t.vt = vt; t.lt = lt;
    t.mem = S ? 1 : 0;
    return dd4ml_ks_launch(state, t, s, kcap);
}

int dd4ml_lsr1_update(double* state, double* ws, const void* s_in, const void* y_in, void* S, void* Y,
                      int64_t n, int64_t ld, int vt, int ht, int kcap, void* stream) {
    if (!state || !ws || !s_in || !y_in || !S || !Y || n < 0 || kcap > DD4ML_MAXM) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(s_in) && aligned16(y_in) && aligned16(S) && aligned16(Y) && ld % 4 == 0;
    int rc = DD4ML_E_DTYPE;
#define CALL(HT, VT) rc = update_t<HT, VT>(state, ws, s_in, y_in, S, Y, n, ld, kcap, vec, s);
    DT_SWITCH2_RC(ht, vt, CALL)
#undef CALL
    if (rc) return rc;
    KTask t{};
    t.kind = KT_UPDATE;
    t.ht = ht; t.vt = vt;
    return dd4ml_ks_launch(state, t, s, kcap);
}

}  // extern "C"
