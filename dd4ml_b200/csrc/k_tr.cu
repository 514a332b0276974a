// dd4ml_b200 kernels -- see stream.cuh for the shared streaming-kernel template.
#include "stream.cuh"

namespace {

// pass 1 of TR.step: ||g||, S^T g, Y^T g  -> sub-problem solve (mode 0/1) or B-coefs (mode 2)
template <class HT, class VT, int KMAX>
struct ProposeOp {
    static constexpr int NRED = 2 + 2 * KMAX;
    __host__ __device__ static constexpr bool is_max(int j) { return j == 1; }
    dd4ml_state* st;
    const VT* g;
    const HT* S;
    const HT* Y;
    int64_t ld;
    int mode;
    Slots<KMAX> sl;
    __device__ void setup() { sl.load(st, mode == 2 || st->second_order[0] != 0.0); }
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) {
        double gv[V];
        ld_ro<V, VT>(g, i, gv);
#pragma unroll
        for (int e = 0; e < V; ++e) { acc[0] += gv[e] * gv[e]; acc[1] = fmax(acc[1], fabs(gv[e])); }
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < sl.k) {
                double sv[V], yv[V];
                ld_ro<V, HT>(S + (int64_t)sl.slot[j] * ld, i, sv);
                ld_ro<V, HT>(Y + (int64_t)sl.slot[j] * ld, i, yv);
#pragma unroll
                for (int e = 0; e < V; ++e) { acc[2 + j] += sv[e] * gv[e]; acc[2 + KMAX + j] += yv[e] * gv[e]; }
            }
    }
    __device__ void finalize(const double* s, ks::Scratch* scr) {
        st->gg[0] = s[0];
        st->ginf[0] = s[1];
        for (int j = 0; j < sl.k; ++j) { st->Sg[sl.slot[j]] = s[2 + j]; st->Yg[sl.slot[j]] = s[2 + KMAX + j]; }
        if (mode == 2) ks::lsr1_B_coefs(st, scr);
        else if (sl.k > 0) ks::tr_solve<HT>(st, mode, scr);   // OBS works in the dtype of Psi (obs.py:88)
        else ks::tr_solve<VT>(st, mode, scr);
    }
};

// pass 2: p = cg g + sum cS_j S_j + cY_j Y_j ; optional w += p ; max|p|
template <class HT, class VT, class WT, int KMAX>
struct ApplyOp {
    static constexpr int NRED = 1;
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
    __device__ void setup() {
        sl.load(st, true);
        cg = st->cg[0];
        conv = st->converged[0] != 0.0;
        int kk = 0;  // keep only slots with non-zero coefficients (first-order step: none)
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < sl.k) {
                const double a = st->cS[sl.slot[j]], b = st->cY[sl.slot[j]];
                if (a != 0.0 || b != 0.0) { cS[kk] = a; cY[kk] = b; sl.slot[kk] = sl.slot[j]; ++kk; }
            }
        sl.k = kk;
    }
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) {
        double pv[V];
        if (conv) {
#pragma unroll
            for (int e = 0; e < V; ++e) pv[e] = 0.0;
            st_v<V, VT>(p, i, pv);
            return;
        }
        double gv[V];
        ld_ro<V, VT>(g, i, gv);
#pragma unroll
        for (int e = 0; e < V; ++e) pv[e] = cg * gv[e];
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < sl.k) {
                double sv[V], yv[V];
                ld_ro<V, HT>(S + (int64_t)sl.slot[j] * ld, i, sv);
                ld_ro<V, HT>(Y + (int64_t)sl.slot[j] * ld, i, yv);
#pragma unroll
                for (int e = 0; e < V; ++e) pv[e] += cS[j] * sv[e] + cY[j] * yv[e];
            }
#pragma unroll
        for (int e = 0; e < V; ++e) { pv[e] = rnd<VT>(pv[e]); acc[0] = fmax(acc[0], fabs(pv[e])); }
        st_v<V, VT>(p, i, pv);
        if (w != nullptr) {
            double wv[V];
            ld_rw<V, WT>(w, i, wv);
#pragma unroll
            for (int e = 0; e < V; ++e) wv[e] += pv[e];
            st_v<V, WT>(w, i, wv);
        }
    }
    __device__ void finalize(const double* s, ks::Scratch*) { st->pmax[0] = s[0]; }
};

// after the trial evaluation of TR.step
template <class HT, class VT, class WT, int KMAX, bool MEM>
struct DecideOp {
    static constexpr int NRED = 5 + 4 * KMAX;
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
    Slots<KMAX> sl;
    bool accept;
    double rho, f0, f1;
    int spare;
    __device__ void setup() {
        sl.load(st, MEM);
        spare = (int)st->spare[0];
        f0 = ld_scalar(loss, lt);
        f1 = ld_scalar(trial, lt);
        if (st->converged[0] != 0.0) { accept = false; rho = 0.0; return; }
        accept = (lt == 0) ? ks::tr_accept<float>(f0, f1, st->pred[0], st->tol[0], st->nu_dec[0], &rho)
                           : ks::tr_accept<double>(f0, f1, st->pred[0], st->tol[0], st->nu_dec[0], &rho);
    }
    __device__ bool active() const { return st->converged[0] == 0.0; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) {
        double pv[V];
        ld_ro<V, VT>(p, i, pv);
        if (!accept) {  // tr.py:216 undo
            double wv[V];
            ld_rw<V, WT>(w, i, wv);
#pragma unroll
            for (int e = 0; e < V; ++e) wv[e] -= pv[e];
            st_v<V, WT>(w, i, wv);
            return;
        }
        double tv[V], ov[V];
        ld_ro<V, VT>(gt, i, tv);
        ld_rw<V, VT>(go, i, ov);
#pragma unroll
        for (int e = 0; e < V; ++e) { acc[3] += tv[e] * tv[e]; acc[4] = fmax(acc[4], fabs(tv[e])); }
        if (MEM) {
            double sv[V], yv[V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                sv[e] = rnd<HT>(pv[e]);
                yv[e] = rnd<HT>(rnd<VT>(tv[e] - ov[e]));
                acc[0] += sv[e] * sv[e];
                acc[1] += sv[e] * yv[e];
                acc[2] += yv[e] * yv[e];
            }
#pragma unroll
            for (int j = 0; j < KMAX; ++j)
                if (j < sl.k) {
                    double a[V], b[V];
                    ld_ro<V, HT>(S + (int64_t)sl.slot[j] * ld, i, a);
                    ld_ro<V, HT>(Y + (int64_t)sl.slot[j] * ld, i, b);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        acc[5 + j] += a[e] * sv[e];
                        acc[5 + KMAX + j] += b[e] * sv[e];
                        acc[5 + 2 * KMAX + j] += a[e] * yv[e];
                        acc[5 + 3 * KMAX + j] += b[e] * yv[e];
                    }
                }
            st_v<V, HT>(S + (int64_t)spare * ld, i, sv);
            st_v<V, HT>(Y + (int64_t)spare * ld, i, yv);
        }
        st_v<V, VT>(gout, i, tv);
    }
    __device__ void finalize(const double* s, ks::Scratch* scr) {
        st->loss[0] = f0;
        st->trial_loss[0] = f1;
        if (st->converged[0] != 0.0) {  // tr.py:147: nothing happened
            st->accept[0] = 0.0;
            st->out_gn[0] = st->gn[0];
            return;
        }
        if (accept) {
            st->c_ss[0] = s[0]; st->c_sy[0] = s[1]; st->c_yy[0] = s[2];
            st->out_gg[0] = s[3]; st->out_ginf[0] = s[4];
            for (int j = 0; j < sl.k; ++j) {
                st->c_Ss[sl.slot[j]] = s[5 + j];
                st->c_Ys[sl.slot[j]] = s[5 + KMAX + j];
                st->c_Sy[sl.slot[j]] = s[5 + 2 * KMAX + j];
                st->c_Yy[sl.slot[j]] = s[5 + 3 * KMAX + j];
            }
        }
        const double so = st->second_order[0];
        if (!MEM) st->second_order[0] = 0.0;
        if (lt == 0) ks::tr_finish<float>(st, accept, rho, scr); else ks::tr_finish<double>(st, accept, rho, scr);
        st->second_order[0] = so;
        st->solve_valid[0] = 0.0;
    }
};

// standalone LSR1.update_memory(s, y) (lsr1.py:53-144)
template <class HT, class VT, int KMAX>
struct UpdateOp {
    static constexpr int NRED = 3 + 4 * KMAX;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    dd4ml_state* st;
    const VT* s_in;
    const VT* y_in;
    HT* S;
    HT* Y;
    int64_t ld;
    Slots<KMAX> sl;
    int spare;
    __device__ void setup() { sl.load(st, true); spare = (int)st->spare[0]; }
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) {
        double sv[V], yv[V];
        ld_ro<V, VT>(s_in, i, sv);
        ld_ro<V, VT>(y_in, i, yv);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            sv[e] = rnd<HT>(sv[e]);
            yv[e] = rnd<HT>(yv[e]);
            acc[0] += sv[e] * sv[e];
            acc[1] += sv[e] * yv[e];
            acc[2] += yv[e] * yv[e];
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < sl.k) {
                double a[V], b[V];
                ld_ro<V, HT>(S + (int64_t)sl.slot[j] * ld, i, a);
                ld_ro<V, HT>(Y + (int64_t)sl.slot[j] * ld, i, b);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    acc[3 + j] += a[e] * sv[e];
                    acc[3 + KMAX + j] += b[e] * sv[e];
                    acc[3 + 2 * KMAX + j] += a[e] * yv[e];
                    acc[3 + 3 * KMAX + j] += b[e] * yv[e];
                }
            }
        st_v<V, HT>(S + (int64_t)spare * ld, i, sv);
        st_v<V, HT>(Y + (int64_t)spare * ld, i, yv);
    }
    __device__ void finalize(const double* s, ks::Scratch* scr) {
        st->c_ss[0] = s[0]; st->c_sy[0] = s[1]; st->c_yy[0] = s[2];
        for (int j = 0; j < sl.k; ++j) {
            st->c_Ss[sl.slot[j]] = s[3 + j];
            st->c_Ys[sl.slot[j]] = s[3 + KMAX + j];
            st->c_Sy[sl.slot[j]] = s[3 + 2 * KMAX + j];
            st->c_Yy[sl.slot[j]] = s[3 + 3 * KMAX + j];
        }
        st->err[0] = 0.0;
        st->pair_tried[0] = 1.0;
        ks::lsr1_try_update(st, scr);
        st->solve_valid[0] = 0.0;
    }
};

}  // namespace

extern "C" {

int dd4ml_tr_propose(double* state, double* ws, const void* g, const void* S, const void* Y,
                     int64_t n, int64_t ld, int vt, int ht, int mode, void* stream) {
    if (!state || !ws || !g || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(g) && (!S || (aligned16(S) && aligned16(Y) && ld % 4 == 0));
#define CALL(HT, VT)                                                                         \
    if (!S) {                                                                                \
        ProposeOp<HT, VT, 0> op{(dd4ml_state*)state, (const VT*)g, nullptr, nullptr, ld, mode}; \
        return launch(op, n, vec, ws, 4, s);                                                 \
    } else {                                                                                 \
        ProposeOp<HT, VT, 8> op{(dd4ml_state*)state, (const VT*)g, (const HT*)S, (const HT*)Y, ld, mode}; \
        return launch(op, n, vec, ws, 3, s);                                                 \
    }
    DT_SWITCH2(ht, vt, CALL)
#undef CALL
}

int dd4ml_tr_apply(double* state, double* ws, const void* g, const void* S, const void* Y, void* p,
                   void* w, int64_t n, int64_t ld, int vt, int ht, int wt, void* stream) {
    if (!state || !ws || !g || !p || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(g) && aligned16(p) && (!w || aligned16(w)) &&
                     (!S || (aligned16(S) && aligned16(Y) && ld % 4 == 0));
#define CALL(HT, VT, WT)                                                                         \
    if (!S) {                                                                                    \
        ApplyOp<HT, VT, WT, 0> op{(dd4ml_state*)state, (const VT*)g, nullptr, nullptr, (VT*)p, (WT*)w, ld}; \
        return launch(op, n, vec, ws, 4, s);                                                     \
    } else {                                                                                     \
        ApplyOp<HT, VT, WT, 8> op{(dd4ml_state*)state, (const VT*)g, (const HT*)S, (const HT*)Y, (VT*)p, (WT*)w, ld}; \
        return launch(op, n, vec, ws, 2, s);                                                     \
    }
    DT_SWITCH3(ht, vt, wt, CALL)
#undef CALL
}

int dd4ml_tr_decide(double* state, double* ws, const void* loss, const void* trial_loss, int lt,
                    const void* g_trial, const void* g_old, void* g_out, const void* p, void* w,
                    void* S, void* Y, int64_t n, int64_t ld, int vt, int ht, int wt, void* stream) {
    if (!state || !ws || !loss || !trial_loss || !g_trial || !g_old || !g_out || !p || !w || n < 0) return DD4ML_E_ARG;
    if (lt != 0 && lt != 1) return DD4ML_E_DTYPE;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(g_trial) && aligned16(g_old) && aligned16(g_out) && aligned16(p) && aligned16(w) &&
                     (!S || (aligned16(S) && aligned16(Y) && ld % 4 == 0));
#define CALL(HT, VT, WT)                                                                          \
    if (!S) {                                                                                     \
        DecideOp<HT, VT, WT, 0, false> op{(dd4ml_state*)state, loss, trial_loss, lt, (const VT*)g_trial, \
            (const VT*)g_old, (VT*)g_out, (const VT*)p, (WT*)w, nullptr, nullptr, ld};            \
        return launch(op, n, vec, ws, 4, s);                                                      \
    } else {                                                                                      \
        DecideOp<HT, VT, WT, 8, true> op{(dd4ml_state*)state, loss, trial_loss, lt, (const VT*)g_trial, \
            (const VT*)g_old, (VT*)g_out, (const VT*)p, (WT*)w, (HT*)S, (HT*)Y, ld};              \
        return launch(op, n, vec, ws, 2, s);                                                      \
    }
    DT_SWITCH3(ht, vt, wt, CALL)
#undef CALL
}

int dd4ml_lsr1_B(double* state, double* ws, const void* v, const void* S, const void* Y, void* out,
                 int64_t n, int64_t ld, int vt, int ht, void* stream) {
    if (!S || !Y) return DD4ML_E_ARG;
    int rc = dd4ml_tr_propose(state, ws, v, S, Y, n, ld, vt, ht, 2, stream);
    if (rc) return rc;
    return dd4ml_tr_apply(state, ws, v, S, Y, out, nullptr, n, ld, vt, ht, vt, stream);
}

int dd4ml_lsr1_update(double* state, double* ws, const void* s_in, const void* y_in, void* S, void* Y,
                      int64_t n, int64_t ld, int vt, int ht, void* stream) {
    if (!state || !ws || !s_in || !y_in || !S || !Y || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(s_in) && aligned16(y_in) && aligned16(S) && aligned16(Y) && ld % 4 == 0;
#define CALL(HT, VT)                                                                               \
    {                                                                                              \
        UpdateOp<HT, VT, 8> op{(dd4ml_state*)state, (const VT*)s_in, (const VT*)y_in, (HT*)S, (HT*)Y, ld}; \
        return launch(op, n, vec, ws, 2, s);                                                       \
    }
    DT_SWITCH2(ht, vt, CALL)
#undef CALL
}

}  // extern "C"
