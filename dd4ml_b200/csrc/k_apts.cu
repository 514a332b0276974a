// dd4ml_b200 kernels -- see stream.cuh for the shared streaming-kernel template.
#include "stream.cuh"

namespace {

template <class VT>
struct ResidualOp {
    static constexpr int NRED = 0;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    const VT* Gg;
    const VT* Gl;
    VT* g0;
    VT* gl0;
    VT* resid;
    __device__ dd4ml_state* state() const { return nullptr; }
    __device__ void setup() {}
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double*) {
        double a[V], b[V], r[V];
        ld_ro<V, VT>(Gg, i, a);
        ld_ro<V, VT>(Gl, i, b);
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = a[e] - b[e];
        st_v<V, VT>(g0, i, a);
        st_v<V, VT>(gl0, i, b);
        if (resid != nullptr) st_v<V, VT>(resid, i, r);
    }
    __device__ void finalize(const double*, dd4ml_state*) {}
};

template <class VT, class WT>
struct FocOp {
    static constexpr int NRED = 2;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    dd4ml_state* st;
    const WT* wl;
    const WT* wg;
    const VT* resid;
    const void* loss_in;
    void* loss_out;
    int lt;
    double tol;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() {}
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) {
        double a[V], b[V], r[V];
        ld_ro<V, WT>(wl, i, a);
        ld_ro<V, WT>(wg, i, b);
        ld_ro<V, VT>(resid, i, r);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const double d = rnd<WT>(a[e] - b[e]);
            acc[0] += d * d;
            acc[1] += r[e] * d;
        }
    }
    __device__ void finalize(const double* s, dd4ml_state*) {
        st->foc_dd[0] = s[0];
        st->foc_rd[0] = s[1];
        const double l = ld_scalar(loss_in, lt);
        st_scalar(loss_out, lt, (sqrt(s[0]) > tol) ? l + s[1] : l);
    }
};

// BT: dtype of the reference's pre-allocated `_step_buffer` / coalesced all-reduce buffer, i.e. the
// dtype the model had when the optimizer was CONSTRUCTED (apts_d.py:112,129-141, apts_p.py:108).  It
// differs from WT when the Trainer converts the model afterwards (trainer.py:1203-1211): the step is
// then rounded to BT and summed over the subdomains in BT.
template <class WT, class BT>
struct PackOp {
    static constexpr int NRED = 0;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    BT* buf;
    const WT* wl;
    const WT* w0;
    const void* loss0;
    const void* loss1;
    int lt;
    double weight, evals;
    int64_t n;
    __device__ dd4ml_state* state() const { return nullptr; }
    __device__ void setup() {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const double red = ld_scalar(loss0, lt) - ld_scalar(loss1, lt);
            buf[n] = (BT)((lt == 0 ? (double)(float)red : red) * weight);
            buf[n + 1] = (BT)evals;
        }
    }
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double*) {
        double a[V], b[V];
        ld_ro<V, WT>(wl, i, a);
        ld_ro<V, WT>(w0, i, b);
#pragma unroll
        for (int e = 0; e < V; ++e) a[e] = rnd<WT>(a[e] - b[e]);
        st_v<V, BT>(buf, i, a);
    }
    __device__ void finalize(const double*, dd4ml_state*) {}
};

template <class WT, class BT>
struct AptsTrialOp {
    static constexpr int NRED = 0;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    WT* w;
    const WT* w0;
    const void* buf;      // mode 0: the summed step (BT); mode 1: the merged parameters (WT)
    int mode;
    double scale, clampv;
    __device__ dd4ml_state* state() const { return nullptr; }
    __device__ void setup() {}
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double*) {
        double a[V], b[V];
        ld_ro<V, WT>(w0, i, a);
        if (mode == 0) ld_ro<V, BT>(reinterpret_cast<const BT*>(buf), i, b);
        else ld_ro<V, WT>(reinterpret_cast<const WT*>(buf), i, b);
        const double cl = rnd<BT>(clampv);   // clamp_ on a BT tensor compares against the bound in BT
#pragma unroll
        for (int e = 0; e < V; ++e) {
            double s;
            if (mode == 0) s = rnd<BT>(b[e] * scale);
            else { s = rnd<BT>(rnd<WT>(b[e] - a[e])); s = fmin(fmax(s, -cl), cl); }
            a[e] += s;
        }
        st_v<V, WT>(w, i, a);
    }
    __device__ void finalize(const double*, dd4ml_state*) {}
};

// The additive-Schwarz combination fused with its exchange (apts_d.py:139-150 + apts_base.py:445): every
// rank reads the packed steps [w_loc - w0 || pred || evals] of ALL subdomains straight out of its peers'
// HBM over NVLink (symmetric-memory buffers, one per rank), sums them in rank order -- fixed order and
// float64 accumulation, so every rank forms bit-identical sums -- and applies w = w0 + scale * sum in the
// same pass.  Replaces pack -> ncclAllReduce -> apts_trial (two extra passes over the n-vector and the
// collective's launch + protocol latency); the caller brackets it with a device-side barrier on the
// symmetric-memory signal pads.
constexpr int MAX_PEERS = 16;
template <class BT> struct PeerTable { const BT* buf[MAX_PEERS]; int n; };

template <class WT, class BT>
struct CombineOp {
    static constexpr int NRED = 0;
    __host__ __device__ static constexpr bool is_max(int) { return false; }
    WT* w;
    const WT* w0;
    PeerTable<BT> peers;
    BT* tail;             // [sum of pred, sum of evals]
    double scale;
    int64_t n;
    __device__ dd4ml_state* state() const { return nullptr; }
    __device__ void setup() {
        if (blockIdx.x == 0 && threadIdx.x < 2) {
            double s = 0.0;
            for (int r = 0; r < peers.n; ++r) s += (double)peers.buf[r][n + threadIdx.x];
            tail[threadIdx.x] = (BT)s;
        }
    }
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double*) {
        double a[V], s[V];
        ld_ro<V, WT>(w0, i, a);
#pragma unroll
        for (int e = 0; e < V; ++e) s[e] = 0.0;
#pragma unroll 4
        for (int r = 0; r < peers.n; ++r) {
            double b[V];
            ld_rw<V, BT>(peers.buf[r], i, b);      // peer memory: plain (coherent) loads, not the read-only path
#pragma unroll
            for (int e = 0; e < V; ++e) s[e] += b[e];
        }
#pragma unroll
        for (int e = 0; e < V; ++e) a[e] += rnd<BT>(rnd<BT>(s[e]) * scale);
        st_v<V, WT>(w, i, a);
    }
    __device__ void finalize(const double*, dd4ml_state*) {}
};

template <class VT, class WT>
struct ControlOp {
    static constexpr int NRED = 2;
    __host__ __device__ static constexpr bool is_max(int j) { return j == 1; }
    dd4ml_state* st;
    const void* f0p;
    const void* f1p;
    int lt;
    const void* predp;
    int pt;
    WT* w;
    const WT* w0;
    VT* g0;
    const VT* G;
    void* loss_out;
    const void* auxp;
    bool accept;
    double f0, f1, pred;
    __device__ dd4ml_state* state() const { return st; }
    __device__ void setup() {
        f0 = ld_scalar(f0p, lt);
        f1 = ld_scalar(f1p, lt);
        pred = ld_scalar(predp, pt);
        accept = (lt == 0) ? ks::apts_accept<float>(f0, f1, pred, st->tol[0], st->nu_dec[0])
                           : ks::apts_accept<double>(f0, f1, pred, st->tol[0], st->nu_dec[0]);
    }
    __device__ bool active() const { return true; }
    template <int V>
    __device__ __forceinline__ void body(int64_t i, double* acc) {
        if (accept) {
            double t[V];
            ld_ro<V, VT>(G, i, t);
#pragma unroll
            for (int e = 0; e < V; ++e) { acc[0] += t[e] * t[e]; acc[1] = fmax(acc[1], fabs(t[e])); }
            st_v<V, VT>(g0, i, t);
        } else {
            double a[V];
            ld_ro<V, WT>(w0, i, a);
            st_v<V, WT>(w, i, a);
        }
    }
    __device__ void finalize(const double* s, dd4ml_state*) {
        const bool acc = (lt == 0) ? ks::apts_control<float>(st, f0, f1, pred) : ks::apts_control<double>(st, f0, f1, pred);
        if (acc) { st->out_gg[0] = s[0]; st->out_ginf[0] = s[1]; }
        st->aux[0] = auxp ? ld_scalar(auxp, pt) : 0.0;
        st_scalar(loss_out, lt, acc ? f1 : f0);
    }
};

}  // namespace

extern "C" {

int dd4ml_apts_residual(const void* Gg, const void* Gl, void* g0, void* gl0, void* resid,
                        int64_t n, int vt, void* stream) {
    if (!Gg || !Gl || !g0 || !gl0 || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(Gg) && aligned16(Gl) && aligned16(g0) && aligned16(gl0) && (!resid || aligned16(resid));
    if (vt == 0) { ResidualOp<float> op{(const float*)Gg, (const float*)Gl, (float*)g0, (float*)gl0, (float*)resid}; return launch(op, n, vec, nullptr, 4, s); }
    if (vt == 1) { ResidualOp<double> op{(const double*)Gg, (const double*)Gl, (double*)g0, (double*)gl0, (double*)resid}; return launch(op, n, vec, nullptr, 4, s); }
    return DD4ML_E_DTYPE;
}

int dd4ml_apts_foc(double* state, double* ws, const void* w_loc, const void* w_glob,
                   const void* resid, const void* loss_in, void* loss_out, int lt, double tol,
                   int64_t n, int vt, int wt, void* stream) {
    if (!state || !ws || !w_loc || !w_glob || !resid || !loss_in || !loss_out || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(w_loc) && aligned16(w_glob) && aligned16(resid);
    if (vt == 0 && wt == 0) { FocOp<float, float> op{(dd4ml_state*)state, (const float*)w_loc, (const float*)w_glob, (const float*)resid, loss_in, loss_out, lt, tol}; return launch(op, n, vec, ws, 4, s); }
    if (vt == 0 && wt == 1) { FocOp<float, double> op{(dd4ml_state*)state, (const double*)w_loc, (const double*)w_glob, (const float*)resid, loss_in, loss_out, lt, tol}; return launch(op, n, vec, ws, 4, s); }
    if (vt == 1 && wt == 1) { FocOp<double, double> op{(dd4ml_state*)state, (const double*)w_loc, (const double*)w_glob, (const double*)resid, loss_in, loss_out, lt, tol}; return launch(op, n, vec, ws, 4, s); }
    return DD4ML_E_DTYPE;
}

int dd4ml_apts_pack(void* buf, const void* w_loc, const void* w0, const void* loss0,
                    const void* loss1, int lt, double weight, double evals, int64_t n, int wt,
                    int bt, void* stream) {
    if (!buf || !w_loc || !w0 || !loss0 || !loss1 || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(buf) && aligned16(w_loc) && aligned16(w0);
    if (wt == 0 && bt == 0) { PackOp<float, float> op{(float*)buf, (const float*)w_loc, (const float*)w0, loss0, loss1, lt, weight, evals, n}; return launch(op, n, vec, nullptr, 4, s); }
    if (wt == 1 && bt == 1) { PackOp<double, double> op{(double*)buf, (const double*)w_loc, (const double*)w0, loss0, loss1, lt, weight, evals, n}; return launch(op, n, vec, nullptr, 4, s); }
    if (wt == 1 && bt == 0) { PackOp<double, float> op{(float*)buf, (const double*)w_loc, (const double*)w0, loss0, loss1, lt, weight, evals, n}; return launch(op, n, vec, nullptr, 4, s); }
    return DD4ML_E_DTYPE;
}

int dd4ml_apts_trial(void* w, const void* w0, const void* buf, int mode, double scale,
                     double clampv, int64_t n, int wt, int bt, void* stream) {
    if (!w || !w0 || !buf || n < 0 || (mode != 0 && mode != 1)) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(w) && aligned16(w0) && aligned16(buf);
    if (wt == 0 && bt == 0) { AptsTrialOp<float, float> op{(float*)w, (const float*)w0, buf, mode, scale, clampv}; return launch(op, n, vec, nullptr, 4, s); }
    if (wt == 1 && bt == 1) { AptsTrialOp<double, double> op{(double*)w, (const double*)w0, buf, mode, scale, clampv}; return launch(op, n, vec, nullptr, 4, s); }
    if (wt == 1 && bt == 0) { AptsTrialOp<double, float> op{(double*)w, (const double*)w0, buf, mode, scale, clampv}; return launch(op, n, vec, nullptr, 4, s); }
    return DD4ML_E_DTYPE;
}

int dd4ml_apts_combine(void* w, const void* w0, const void* const* peer_bufs, int nranks, double scale,
                       void* tail_out, int64_t n, int wt, int bt, void* stream) {
    if (!w || !w0 || !peer_bufs || !tail_out || nranks < 1 || nranks > MAX_PEERS || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    bool vec = aligned16(w) && aligned16(w0);
    for (int r = 0; r < nranks; ++r) {
        if (!peer_bufs[r]) return DD4ML_E_ARG;
        vec = vec && aligned16(peer_bufs[r]);
    }
#define COMBINE(WT, BT)                                                                                   \
    {                                                                                                     \
        PeerTable<BT> pt;                                                                                 \
        pt.n = nranks;                                                                                    \
        for (int r = 0; r < MAX_PEERS; ++r) pt.buf[r] = r < nranks ? (const BT*)peer_bufs[r] : nullptr;   \
        CombineOp<WT, BT> op{(WT*)w, (const WT*)w0, pt, (BT*)tail_out, scale, n};                         \
        return launch(op, n, vec, nullptr, 4, s);                                                         \
    }
    if (wt == 0 && bt == 0) COMBINE(float, float)
    if (wt == 1 && bt == 1) COMBINE(double, double)
    if (wt == 1 && bt == 0) COMBINE(double, float)
#undef COMBINE
    return DD4ML_E_DTYPE;
}

int dd4ml_apts_control(double* state, double* ws, const void* f0, const void* ftrial, int lt,
                       const void* pred, int pt, void* w, const void* w0, void* g0, const void* G,
                       void* loss_out, const void* aux, int64_t n, int vt, int wt, void* stream) {
    if (!state || !ws || !f0 || !ftrial || !pred || !w || !w0 || !g0 || !G || !loss_out || n < 0) return DD4ML_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(w) && aligned16(w0) && aligned16(g0) && aligned16(G);
    if (vt == 0 && wt == 0) { ControlOp<float, float> op{(dd4ml_state*)state, f0, ftrial, lt, pred, pt, (float*)w, (const float*)w0, (float*)g0, (const float*)G, loss_out, aux}; return launch(op, n, vec, ws, 4, s); }
    if (vt == 0 && wt == 1) { ControlOp<float, double> op{(dd4ml_state*)state, f0, ftrial, lt, pred, pt, (double*)w, (const double*)w0, (float*)g0, (const float*)G, loss_out, aux}; return launch(op, n, vec, ws, 4, s); }
    if (vt == 1 && wt == 1) { ControlOp<double, double> op{(dd4ml_state*)state, f0, ftrial, lt, pred, pt, (double*)w, (const double*)w0, (double*)g0, (const double*)G, loss_out, aux}; return launch(op, n, vec, ws, 4, s); }
    return DD4ML_E_DTYPE;
}

}  // extern "C"
