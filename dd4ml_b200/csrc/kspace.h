// k-space routines of the trust-region step: everything that happens between the
// streaming reductions over the n-vectors and the streaming output pass.
//
// All inputs are O(k^2) numbers (Gram blocks of the L-SR1 pairs, Psi^T g, g^T g)
// held in the device state block; the routines run in ONE thread of the last CTA of
// the reduction kernel ("small m x m inner solve kept in one CTA").  The file is
// plain C++ guarded for host compilation as well, so tests/hostshim can run the
// same source on the CPU against the oracle.
//
// Reference behaviour restated here (cruzas/DD4ML):
//   LSR1.precompute / B / update_memory      optimizers/lsr1.py:53-198
//   OBS.solve_tr_subproblem / Newton / SMW   solvers/obs.py:46-254
//   solve_tr_first_order / second_order      utility/optimizer_utils.py:197-307
//   TR.step ratio test and radius update     optimizers/tr.py:186-221
//   LSSR1_TR clip-to-radius / descent flip   optimizers/lssr1_tr.py:563-587
// Linear algebra (Cholesky, LU, symmetric eigen-decomposition) runs in double;
// the quantities the reference thresholds or iterates on in its working dtype
// (Lambda, a_j, the Newton iteration, rho) are rounded to that dtype T first.
#pragma once
#include <math.h>

#include "state.h"

#ifdef __CUDACC__
#define KS_FN __host__ __device__
#else
#define KS_FN
#endif
#if defined(__CUDA_ARCH__)
#define KS_CLOCK() ((double)clock64())
#define KS_RSQRT(x) rsqrt(x)
#else
#define KS_CLOCK() 0.0
#define KS_RSQRT(x) (1.0 / sqrt(x))
#endif

namespace ks {

constexpr int M = DD4ML_MAXM;
constexpr int MS = DD4ML_MAXS;
constexpr double OBS_TOL = 1e-6;  // obs.py:29

KS_FN inline bool finite(double x) { return x == x && fabs(x) <= 1.79769313486231570e308; }
KS_FN inline float ks_sqrt(float x) { return sqrtf(x); }
KS_FN inline double ks_sqrt(double x) { return sqrt(x); }

// ---- dense helpers on row-major k x k blocks with leading dimension M ----------
// R upper triangular with R^T R = G (torch.linalg.cholesky(upper=True)).
KS_FN inline bool chol_upper(const double* G, int k, double* R) {
    for (int i = 0; i < k * M; ++i) R[i] = 0.0;
    for (int j = 0; j < k; ++j) {
        double s = G[j * M + j];
        for (int i = 0; i < j; ++i) s -= R[i * M + j] * R[i * M + j];
        if (!(s > 0.0) || !finite(s)) return false;
        double d = sqrt(s);
        R[j * M + j] = d;
        for (int c = j + 1; c < k; ++c) {
            double t = G[j * M + c];
            for (int i = 0; i < j; ++i) t -= R[i * M + j] * R[i * M + c];
            R[j * M + c] = t / d;
        }
    }
    return true;
}

// In-place LU with partial pivoting (torch.linalg.solve = getrf/getrs).
KS_FN inline bool lu_factor(double* A, int k, int* piv) {
    for (int c = 0; c < k; ++c) {
        int p = c;
        double best = fabs(A[c * M + c]);
        for (int r = c + 1; r < k; ++r) {
            double v = fabs(A[r * M + c]);
            if (v > best) { best = v; p = r; }
        }
        piv[c] = p;
        if (best == 0.0 || !finite(best)) return false;
        if (p != c)
            for (int j = 0; j < k; ++j) { double t = A[c * M + j]; A[c * M + j] = A[p * M + j]; A[p * M + j] = t; }
        double inv = 1.0 / A[c * M + c];
        for (int r = c + 1; r < k; ++r) {
            double f = A[r * M + c] * inv;
            A[r * M + c] = f;
            for (int j = c + 1; j < k; ++j) A[r * M + j] -= f * A[c * M + j];
        }
    }
    return true;
}

KS_FN inline void lu_solve(const double* LU, const int* piv, int k, double* b) {
    for (int c = 0; c < k; ++c) {
        int p = piv[c];
        if (p != c) { double t = b[c]; b[c] = b[p]; b[p] = t; }
    }
    for (int r = 1; r < k; ++r) {
        double s = b[r];
        for (int j = 0; j < r; ++j) s -= LU[r * M + j] * b[j];
        b[r] = s;
    }
    for (int r = k - 1; r >= 0; --r) {
        double s = b[r];
        for (int j = r + 1; j < k; ++j) s -= LU[r * M + j] * b[j];
        b[r] = s / LU[r * M + r];
    }
}

// Cyclic Jacobi for a symmetric k x k matrix.  On exit w ascending, V[:, j] the
// eigenvector of w[j], each with its largest-|.| component positive (lowest index
// on ties): the documented sign convention of this implementation.
KS_FN inline int jacobi_eigh(double* A, int k, double* w, double* V) {
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) V[i * M + j] = (i == j) ? 1.0 : 0.0;
    int sweeps = 0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < k; ++i) {
            diag += A[i * M + i] * A[i * M + i];
            for (int j = i + 1; j < k; ++j) off += A[i * M + j] * A[i * M + j];
        }
        // off-diagonal mass below 1e-26 of the diagonal mass: eigenvalues converged to ~1e-26
        // relative (Jacobi converges quadratically), eigenvectors to ~1e-13
        if (off <= 1e-26 * diag || off == 0.0) break;
        ++sweeps;
        for (int p = 0; p < k - 1; ++p)
            for (int q = p + 1; q < k; ++q) {
                double apq = A[p * M + q];
                if (apq == 0.0) continue;
                // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (aqq - app) / (2 apq),
                // written with one division and one square root
                const double d = A[q * M + q] - A[p * M + p];
                const double a2 = 2.0 * apq;
                const double sg = ((d >= 0.0) == (a2 >= 0.0)) ? 1.0 : -1.0;
                const double t = sg * fabs(a2) / (fabs(d) + sqrt(d * d + a2 * a2));
                const double c = KS_RSQRT(t * t + 1.0), s = t * c;
                // Each update loads its two vectors into registers first (constant trip count M,
                // predicated on r < k): the loads issue back to back instead of one dependent
                // shared-memory round trip per element -- the rotation is ~4x faster for k = 10.
                double u[M], v[M];
#pragma unroll
                for (int r = 0; r < M; ++r)   // columns p, q of A
                    if (r < k) { u[r] = A[r * M + p]; v[r] = A[r * M + q]; }
#pragma unroll
                for (int r = 0; r < M; ++r)
                    if (r < k) { A[r * M + p] = c * u[r] - s * v[r]; A[r * M + q] = s * u[r] + c * v[r]; }
#pragma unroll
                for (int r = 0; r < M; ++r)   // rows p, q of A
                    if (r < k) { u[r] = A[p * M + r]; v[r] = A[q * M + r]; }
#pragma unroll
                for (int r = 0; r < M; ++r)
                    if (r < k) { A[p * M + r] = c * u[r] - s * v[r]; A[q * M + r] = s * u[r] + c * v[r]; }
#pragma unroll
                for (int r = 0; r < M; ++r)
                    if (r < k) { u[r] = V[r * M + p]; v[r] = V[r * M + q]; }
#pragma unroll
                for (int r = 0; r < M; ++r)
                    if (r < k) { V[r * M + p] = c * u[r] - s * v[r]; V[r * M + q] = s * u[r] + c * v[r]; }
            }
    }
    for (int i = 0; i < k; ++i) w[i] = A[i * M + i];
    for (int i = 0; i < k - 1; ++i) {  // selection sort, ascending
        int m = i;
        for (int j = i + 1; j < k; ++j)
            if (w[j] < w[m]) m = j;
        if (m != i) {
            double t = w[i]; w[i] = w[m]; w[m] = t;
            for (int r = 0; r < k; ++r) { double u = V[r * M + i]; V[r * M + i] = V[r * M + m]; V[r * M + m] = u; }
        }
    }
    for (int j = 0; j < k; ++j) {
        int im = 0;
        double best = -1.0;
        for (int r = 0; r < k; ++r)
            if (fabs(V[r * M + j]) > best) { best = fabs(V[r * M + j]); im = r; }
        if (V[im * M + j] < 0.0)
            for (int r = 0; r < k; ++r) V[r * M + j] = -V[r * M + j];
    }
    return sweeps;
}

// ---- secular equation, in the reference's working dtype T (obs.py:184-254) -----
template <class T>
KS_FN inline void phibar_fg(T sigma, const T* Lam, const T* a, int m, T delta, T tol, T& f, T& g, bool& deg) {
    deg = false;
    for (int j = 0; j < m; ++j) {
        T D = Lam[j] + sigma;
        if (fabs(a[j]) < tol || fabs(D) < tol) deg = true;
    }
    if (deg) {  // obs.py:242-245
        f = T(-1) / delta;
        g = T(1000000.0);
        return;
    }
    T n2 = T(0), s3 = T(0);
    for (int j = 0; j < m; ++j) {
        T D = Lam[j] + sigma;
        T p = a[j] / D;
        n2 += p * p;
        s3 += (a[j] * a[j]) / (D * D * D);
    }
    T nrm = ks_sqrt(n2);
    f = T(1) / nrm - T(1) / delta;
    g = s3 / (nrm * nrm * nrm);
}

template <class T>
KS_FN inline T phibar_f(T sigma, const T* Lam, const T* a, int m, T delta, T tol) {
    T n2 = T(0);
    for (int j = 0; j < m; ++j) {
        T D = Lam[j] + sigma;
        if (fabs(a[j]) < tol || fabs(D) < tol) return T(-1) / delta;
    }
    for (int j = 0; j < m; ++j) {
        T p = a[j] / (Lam[j] + sigma);
        n2 += p * p;
    }
    return T(1) / ks_sqrt(n2) - T(1) / delta;
}

template <class T>
KS_FN inline T newton(T x0, const T* Lam, const T* a, int m, T delta, T tol, int& iters, int& ndeg) {
    T x = x0, f, g;
    bool deg;
    bool adeg = false;
    for (int j = 0; j < m; ++j) adeg = adeg || (fabs(a[j]) < tol);
    if (adeg) {
        // some |a_j| < tol: phibar_fg returns (-1/delta, 1/tol) for EVERY sigma (obs.py:242-245),
        // so the reference's loop is sigma <- sigma - f/g, 200 times (or not at all when
        // |f| <= tol).  Same fp32/fp64 rounding sequence, without re-evaluating phibar.
        f = T(-1) / delta;
        g = T(1000000.0);
        iters = 0;
        ndeg = 1;
        const T inc = f / g;
        if (fabs(f) > tol)
            for (; iters < 200; ++iters) x = x - inc;
        ndeg += iters;
        return x;
    }
    phibar_fg<T>(x, Lam, a, m, delta, tol, f, g, deg);
    iters = 0;
    ndeg = deg ? 1 : 0;
    while (fabs(f) > tol && iters < 200) {  // obs.py:217
        x = x - f / g;
        phibar_fg<T>(x, Lam, a, m, delta, tol, f, g, deg);
        ndeg += deg ? 1 : 0;
        ++iters;
    }
    return x;
}

// ---- logical (oldest..newest) views of the Gram blocks --------------------------
struct Compact {
    int k;
    double gamma;
    double SS[M * M], SY[M * M], YY[M * M];
    double Minv[M * M];  // diag(S^T Y) + L + L^T - gamma S^T S + tol*||.||_F I   (lsr1.py:168-179)
    double G[M * M];     // Psi^T Psi
};

// Work arrays of the k-space routines.  On the device this lives in shared memory of
// the last CTA (keeps the per-thread stack frame of the streaming kernels small).
struct Scratch {
    Compact C;
    double R[M * M], LU[M * M], MR[M * M], RMR[M * M], U[M * M], RinvU[M * M], W[M * M];
    double D[M], b[M], x[M], y[M], q[M], wv[M], col[M];
    int piv[M], pw[M];
};

KS_FN inline void build_compact(const dd4ml_state* st, Compact& c) {
    const int k = (int)st->k[0];
    const double gam = st->gamma[0];
    c.k = k;
    c.gamma = gam;
    for (int i = 0; i < k; ++i) {
        const int pi = (int)st->order[i];
        for (int j = 0; j < k; ++j) {
            const int pj = (int)st->order[j];
            c.SS[i * M + j] = st->SS[pi * MS + pj];
            c.SY[i * M + j] = st->SY[pi * MS + pj];
            c.YY[i * M + j] = st->YY[pi * MS + pj];
        }
    }
    double fro = 0.0;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            double sy = (i >= j) ? c.SY[i * M + j] : c.SY[j * M + i];
            double v = sy - gam * c.SS[i * M + j];
            c.Minv[i * M + j] = v;
            fro += v * v;
            c.G[i * M + j] = c.YY[i * M + j] - gam * (c.SY[i * M + j] + c.SY[j * M + i]) + gam * gam * c.SS[i * M + j];
        }
    const double reg = st->tol[0] * sqrt(fro);
    for (int i = 0; i < k; ++i) c.Minv[i * M + i] += reg;
}

KS_FN inline double dotk(const double* a, const double* b, int k) {
    double s = 0.0;
    for (int i = 0; i < k; ++i) s += a[i] * b[i];
    return s;
}
KS_FN inline double quadk(const double* A, const double* x, int k) {
    double s = 0.0;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) s += x[i] * A[i * M + j] * x[j];
    return s;
}

// p = cg*g + Psi*c  ->  p^T p, g^T p via the Gram identities
struct StepK {
    double cg;
    double c[M];
};
KS_FN inline double step_pp(const Compact& C, const double* b, double gg, const StepK& s) {
    return s.cg * s.cg * gg + 2.0 * s.cg * dotk(s.c, b, C.k) + quadk(C.G, s.c, C.k);
}
KS_FN inline double step_gp(const Compact& C, const double* b, double gg, const StepK& s) {
    return s.cg * gg + dotk(s.c, b, C.k);
}

// register-resident variants for k <= KS_SMALL_MAX (kspace_small.h, included at the end of this file):
// fully unrolled, compile-time sized; k = 5 (TR's hard-coded memory length) runs 3x faster than
// through the generic shared-memory-scratch routines (185 k -> ~60 k cycles)
constexpr int KS_SMALL_MAX = 5;   // (6 doubles the compile time of every kernel file for a rare memory length)
template <class T, int K>
KS_FN bool second_order_s(dd4ml_state* st, double gn, double gg, double delta, double& cg_out, double (&c_out)[M],
                          double& pred, double& pp, double& gp);
template <int K>
KS_FN bool update_us_uu_s(dd4ml_state* st, double ss, double sy, double yy, double& us, double& uu);

// =================================================================================
// The trust-region sub-problem.  mode 0: TR (tr.py:150-176), mode 1: LSSR1_TR
// (additionally clip to the radius and flip to a descent direction,
// lssr1_tr.py:563-587).  Reads gg, ginf, Sg, Yg and the L-SR1 blocks from the
// state; writes the step descriptor (cg, cS, cY) and the report fields.
// =================================================================================
template <class T>
KS_FN inline void tr_solve(dd4ml_state* st, int mode, Scratch* ws) {
    const double delta = st->delta[0], tol = st->tol[0], gg = st->gg[0];
    const int k = (st->second_order[0] != 0.0) ? (int)st->k[0] : 0;
    const double c_s0 = KS_CLOCK();
    st->cyc_eigh[0] = st->cyc_newton[0] = st->jacobi_sweeps[0] = 0.0;
    for (int i = 0; i < MS; ++i) { st->cS[i] = 0.0; st->cY[i] = 0.0; }
    st->cg[0] = 0.0;
    st->err[0] = DD4ML_ERR_NONE;
    st->sigma[0] = st->sigma0[0] = st->tau[0] = 0.0;
    st->newton_iters[0] = st->newton_deg[0] = 0.0;
    st->dogleg_branch[0] = 0.0;
    st->flip[0] = 0.0;
    st->clip_scale[0] = 1.0;
    st->lam_min[0] = st->gBg[0] = 0.0;
    st->converged[0] = 0.0;
    st->solve_valid[0] = 1.0;

    const T gnT = (T)((st->norm_inf[0] != 0.0) ? st->ginf[0] : sqrt(gg));
    const double gn = (double)gnT;
    st->gn[0] = gn;
    if (gn <= tol) {  // tr.py:147 / lssr1_tr.py:509
        st->converged[0] = 1.0;
        st->pred[0] = st->pnorm[0] = st->gp[0] = st->psq[0] = st->dphi0[0] = 0.0;
        st->case_id[0] = DD4ML_CASE_FIRST_ORDER;
        return;
    }

    double pred, pp, gp;
    Compact& C = ws->C;
    C.k = 0;
    double* b = ws->b;
    StepK S;
    for (int i = 0; i < M; ++i) S.c[i] = 0.0;

    if (k == 0) {  // optimizer_utils.py:197-215
        const T scale = (T(1) / gnT) * (T)delta;  // python: delta / gn  ==  gn.reciprocal() * delta
        S.cg = -(double)scale;
        pred = (double)((T)delta * gnT);
        pp = S.cg * S.cg * gg;
        gp = S.cg * gg;
        st->case_id[0] = DD4ML_CASE_FIRST_ORDER;
    } else if (k <= KS_SMALL_MAX) {
        bool ok;
        switch (k) {
            case 1: ok = second_order_s<T, 1>(st, gn, gg, delta, S.cg, S.c, pred, pp, gp); break;
            case 2: ok = second_order_s<T, 2>(st, gn, gg, delta, S.cg, S.c, pred, pp, gp); break;
            case 3: ok = second_order_s<T, 3>(st, gn, gg, delta, S.cg, S.c, pred, pp, gp); break;
            case 4: ok = second_order_s<T, 4>(st, gn, gg, delta, S.cg, S.c, pred, pp, gp); break;
            default: ok = second_order_s<T, 5>(st, gn, gg, delta, S.cg, S.c, pred, pp, gp); break;
        }
        if (!ok) { st->solve_valid[0] = 0.0; return; }
        C.k = k;
        C.gamma = st->gamma[0];
    } else {
        build_compact(st, C);
        const double gam = C.gamma;
        bool ok = finite(gg) && finite(delta) && finite(gam);
        for (int i = 0; i < k; ++i) {
            const int pi = (int)st->order[i];
            b[i] = st->Yg[pi] - gam * st->Sg[pi];  // Psi^T g
            ok = ok && finite(b[i]);
            for (int j = 0; j < k; ++j) ok = ok && finite(C.G[i * M + j]) && finite(C.Minv[i * M + j]);
        }
        if (!ok) { st->err[0] = DD4ML_ERR_NONFINITE; st->solve_valid[0] = 0.0; return; }
        if (delta < 0) { st->err[0] = DD4ML_ERR_NEG_DELTA; st->solve_valid[0] = 0.0; return; }

        // ---- OBS (obs.py:61-95) ----
        double *R = ws->R, *LU = ws->LU, *MR = ws->MR, *RMR = ws->RMR, *U = ws->U, *D = ws->D, *RinvU = ws->RinvU;
        int* piv = ws->piv;
        if (!chol_upper(C.G, k, R)) { st->err[0] = DD4ML_ERR_CHOLESKY; st->solve_valid[0] = 0.0; return; }
        for (int i = 0; i < k * M; ++i) LU[i] = C.Minv[i];
        if (!lu_factor(LU, k, piv)) { st->err[0] = DD4ML_ERR_SINGULAR; st->solve_valid[0] = 0.0; return; }
        for (int c = 0; c < k; ++c) {  // MR[:, c] = Minv \ R^T[:, c]
            double* col = ws->col;
            for (int r = 0; r < k; ++r) col[r] = R[c * M + r];
            lu_solve(LU, piv, k, col);
            for (int r = 0; r < k; ++r) MR[r * M + c] = col[r];
        }
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                double s = 0.0;
                for (int r = 0; r < k; ++r) s += R[i * M + r] * MR[r * M + j];
                RMR[i * M + j] = s;
            }
        for (int i = 0; i < k; ++i)
            for (int j = i + 1; j < k; ++j) {
                double s = 0.5 * (RMR[i * M + j] + RMR[j * M + i]);
                RMR[i * M + j] = RMR[j * M + i] = s;
            }
        const double c_e0 = KS_CLOCK();
        st->jacobi_sweeps[0] = jacobi_eigh(RMR, k, D, U);
        st->cyc_eigh[0] = KS_CLOCK() - c_e0;
        T Lam[M + 1], a[M + 1];
        const T tolT = (T)OBS_TOL, gamT = (T)gam, delT = (T)delta;
        for (int i = 0; i < k; ++i) Lam[i] = (T)D[i] + gamT;
        Lam[k] = gamT;
        for (int i = 0; i <= k; ++i)
            if (fabs(Lam[i]) < tolT) Lam[i] = T(0);
        const T lam_min = (Lam[0] < gamT) ? Lam[0] : gamT;
        st->lam_min[0] = (double)lam_min;
        for (int c = 0; c < k; ++c) {  // RinvU[:, c] = R \ U[:, c]
            for (int r = k - 1; r >= 0; --r) {
                double s = U[r * M + c];
                for (int j = r + 1; j < k; ++j) s -= R[r * M + j] * RinvU[j * M + c];
                RinvU[r * M + c] = s / R[r * M + r];
            }
        }
        double gpgp = 0.0;
        for (int i = 0; i < k; ++i) {
            double s = 0.0;
            for (int r = 0; r < k; ++r) s += RinvU[r * M + i] * b[r];
            a[i] = (T)s;
            gpgp += s * s;
        }
        {
            double d = gg - gpgp;
            a[k] = (T)sqrt(d > 0.0 ? d : 0.0);
        }
        T tauT;
        T h2 = T(0);
        for (int i = 0; i <= k; ++i) { T h = a[i] / Lam[i]; h2 += h * h; }
        if (lam_min > T(0) && ks_sqrt(h2) <= delT) {  // case A, obs.py:97
            tauT = gamT;
            st->case_id[0] = DD4ML_CASE_A;
        } else {
            if (lam_min <= T(0) && phibar_f<T>(-lam_min, Lam, a, k + 1, delT, tolT) >= T(0)) {
                st->err[0] = DD4ML_ERR_HARD_CASE;  // obs.py:100-158, unreachable for gamma > 0
                st->solve_valid[0] = 0.0;
                return;
            }
            T x0 = T(0);
            if (!(lam_min > T(0))) {  // obs.py:162-168
                T sh = a[0] / delT - Lam[0];
                for (int i = 1; i <= k; ++i) { T v = a[i] / delT - Lam[i]; if (v > sh) sh = v; }
                x0 = (sh > -lam_min) ? sh : -lam_min;
            }
            int it, nd;
            const double c_n0 = KS_CLOCK();
            T sig = newton<T>(x0, Lam, a, k + 1, delT, tolT, it, nd);
            st->cyc_newton[0] = KS_CLOCK() - c_n0;
            if (!finite((double)sig)) sig = newton<T>(T(0), Lam, a, k + 1, delT, tolT, it, nd);
            st->sigma0[0] = (double)x0;
            st->sigma[0] = (double)sig;
            st->newton_iters[0] = it;
            st->newton_deg[0] = nd;
            st->case_id[0] = DD4ML_CASE_C;
            tauT = gamT + sig;
        }
        const double tau = (double)tauT;
        st->tau[0] = tau;
        // SMW step exactly as coded (obs.py:175-182): p = -g/tau + Psi (tau Minv + G)^{-1} Psi^T g
        double *W = ws->W, *wv = ws->wv;
        for (int i = 0; i < k * M; ++i) W[i] = tau * C.Minv[i] + C.G[i];
        int* pw = ws->pw;
        if (!lu_factor(W, k, pw)) { st->err[0] = DD4ML_ERR_SINGULAR; st->solve_valid[0] = 0.0; return; }
        for (int i = 0; i < k; ++i) wv[i] = b[i];
        lu_solve(W, pw, k, wv);
        StepK Pb;
        Pb.cg = (double)(T(-1) / tauT);
        for (int i = 0; i < M; ++i) Pb.c[i] = (i < k) ? wv[i] : 0.0;

        // ---- Cauchy point and dogleg (optimizer_utils.py:259-296) ----
        double* x = ws->x;
        for (int i = 0; i < k; ++i) x[i] = b[i];
        lu_solve(LU, piv, k, x);
        const double gBg = gam * gg + dotk(b, x, k);
        st->gBg[0] = gBg;
        S = Pb;
        if (st->dogleg[0] != 0.0 && gBg > 0.0) {
            const double alpha = gn * gn / gBg;
            const double cc = (alpha * sqrt(gg) >= delta) ? -(delta / gn) : -alpha;
            const double pb_norm = sqrt(step_pp(C, b, gg, Pb));
            const double pc_norm = fabs(cc) * sqrt(gg);
            if (pb_norm <= delta) {
                st->dogleg_branch[0] = 1.0;
            } else if (pc_norm >= delta) {
                S.cg = cc;
                for (int i = 0; i < k; ++i) S.c[i] = 0.0;
                st->dogleg_branch[0] = 2.0;
            } else {
                StepK d = Pb;
                d.cg = Pb.cg - cc;
                const double qa = step_pp(C, b, gg, d);
                const double qb = 2.0 * cc * step_gp(C, b, gg, d);
                const double qc = cc * cc * gg - delta * delta;
                const double disc = qb * qb - 4.0 * qa * qc;
                if (disc < 0.0) {
                    S.cg = cc;
                    for (int i = 0; i < k; ++i) S.c[i] = 0.0;
                    st->dogleg_branch[0] = 3.0;
                } else {
                    const double t = (-qb + sqrt(disc)) / (2.0 * qa);
                    S.cg = cc + t * d.cg;
                    for (int i = 0; i < k; ++i) S.c[i] = t * d.c[i];
                    st->dogleg_branch[0] = 4.0;
                }
            }
        }
        // ---- predicted reduction -(g^T p + 1/2 p^T B p) (optimizer_utils.py:298-305) ----
        pp = step_pp(C, b, gg, S);
        gp = step_gp(C, b, gg, S);
        double *q = ws->q, *y = ws->y;
        for (int i = 0; i < k; ++i) {
            double s = S.cg * b[i];
            for (int j = 0; j < k; ++j) s += C.G[i * M + j] * S.c[j];
            q[i] = s;
        }
        for (int i = 0; i < k; ++i) y[i] = q[i];
        lu_solve(LU, piv, k, y);
        const double pBp = gam * pp + dotk(q, y, k);
        pred = -(gp + 0.5 * pBp);
    }

    if (mode == 1) {  // lssr1_tr.py:563-587 (the momentum term is identically zero)
        if (pp > tol) {
            const double nrm = sqrt((double)(T)pp);
            const double sc = (delta / nrm < 1.0) ? delta / nrm : 1.0;
            S.cg *= sc;
            for (int i = 0; i < C.k; ++i) S.c[i] *= sc;
            pp *= sc * sc;
            gp *= sc;
            st->clip_scale[0] = sc;
        }
        if (gp > 0.0) {
            S.cg = -S.cg;
            for (int i = 0; i < C.k; ++i) S.c[i] = -S.c[i];
            pred = -pred;
            gp = -gp;
            st->flip[0] = 1.0;
        }
    }
    st->cg[0] = S.cg;
    for (int i = 0; i < C.k; ++i) {
        const int pi = (int)st->order[i];
        st->cY[pi] = S.c[i];
        st->cS[pi] = -C.gamma * S.c[i];
    }
    st->pred[0] = pred;
    st->psq[0] = pp;
    st->pnorm[0] = sqrt(pp > 0.0 ? pp : 0.0);
    st->gp[0] = gp;
    st->dphi0[0] = gp;
    st->cyc_solve[0] = KS_CLOCK() - c_s0;
}

// =================================================================================
// LSR1.update_memory acceptance chain (lsr1.py:66-144) on the candidate pair whose
// reductions sit in c_ss, c_sy, c_yy, c_Ss, c_Ys, c_Sy, c_Yy (PHYSICAL slot index)
// and whose vectors were written to the spare slot.  On acceptance the spare slot
// joins the ring (dropping the oldest pair at capacity), the Gram blocks grow by one
// row/column and gamma <- max(y^T y / y^T s, tol).
// =================================================================================
KS_FN inline bool lsr1_try_update(dd4ml_state* st, Scratch* ws) {
    const double tol = st->tol[0];
    const double ss = st->c_ss[0], sy = st->c_sy[0], yy = st->c_yy[0];
    const int k = (int)st->k[0], m = (int)st->mem_len[0];
    const double gam = st->gamma[0];
    st->pair_ok[0] = 0.0;
    const double sn = sqrt(ss), yn = sqrt(yy);
    if (!(sn > tol) || !(yn > tol)) return false;        // :72
    if (!(fabs(sy) > tol * (sn * yn))) return false;     // :78
    double us, uu;
    if (k == 0) {  // B = gamma I
        us = sy - gam * ss;
        uu = yy - 2.0 * gam * sy + gam * gam * ss;
    } else if (k <= KS_SMALL_MAX) {
        bool ok;
        switch (k) {
            case 1: ok = update_us_uu_s<1>(st, ss, sy, yy, us, uu); break;
            case 2: ok = update_us_uu_s<2>(st, ss, sy, yy, us, uu); break;
            case 3: ok = update_us_uu_s<3>(st, ss, sy, yy, us, uu); break;
            case 4: ok = update_us_uu_s<4>(st, ss, sy, yy, us, uu); break;
            default: ok = update_us_uu_s<5>(st, ss, sy, yy, us, uu); break;
        }
        if (!ok) return false;
    } else {
        Compact& C = ws->C;
        build_compact(st, C);
        double *bs = ws->b, *by = ws->y, *x = ws->x, *LU = ws->LU;
        int* piv = ws->piv;
        for (int i = 0; i < k; ++i) {
            const int pi = (int)st->order[i];
            bs[i] = st->c_Ys[pi] - gam * st->c_Ss[pi];   // Psi^T s
            by[i] = st->c_Yy[pi] - gam * st->c_Sy[pi];   // Psi^T y
            x[i] = bs[i];
        }
        for (int i = 0; i < k * M; ++i) LU[i] = C.Minv[i];
        if (!lu_factor(LU, k, piv)) { st->err[0] = DD4ML_ERR_SINGULAR; return false; }
        lu_solve(LU, piv, k, x);                         // x = Minv \ Psi^T s
        us = sy - gam * ss - dotk(bs, x, k);
        uu = yy + gam * gam * ss + quadk(C.G, x, k) - 2.0 * gam * sy - 2.0 * dotk(by, x, k)
             + 2.0 * gam * dotk(bs, x, k);
    }
    const double un = sqrt(uu > 0.0 ? uu : 0.0);
    if (!(fabs(us) > tol * sn * un)) return false;       // :86
    double gc = yy / sy;
    if (gc < tol) gc = tol;                              // :90
    const double pc2 = yy - 2.0 * gc * sy + gc * gc * ss;
    if (!(sqrt(pc2 > 0.0 ? pc2 : 0.0) > tol * yn)) return false;  // :93
    // ---- accept ----
    const int ns = (int)st->spare[0];
    int newspare;
    int kk = k;
    if (k >= m) {  // drop the oldest pair; its slot becomes the spare
        newspare = (int)st->order[0];
        for (int i = 0; i + 1 < k; ++i) st->order[i] = st->order[i + 1];
        st->order[k - 1] = ns;
    } else {
        st->order[k] = ns;
        kk = k + 1;
        // first unused physical slot
        bool used[MS];
        for (int i = 0; i < MS; ++i) used[i] = false;
        for (int i = 0; i < kk; ++i) used[(int)st->order[i]] = true;
        newspare = 0;
        while (newspare < MS && used[newspare]) ++newspare;
    }
    for (int i = 0; i < kk; ++i) {
        const int pj = (int)st->order[i];
        if (pj == ns) continue;
        st->SS[ns * MS + pj] = st->SS[pj * MS + ns] = st->c_Ss[pj];
        st->YY[ns * MS + pj] = st->YY[pj * MS + ns] = st->c_Yy[pj];
        st->SY[pj * MS + ns] = st->c_Sy[pj];  // S_j^T y
        st->SY[ns * MS + pj] = st->c_Ys[pj];  // s^T Y_j
    }
    st->SS[ns * MS + ns] = ss;
    st->SY[ns * MS + ns] = sy;
    st->YY[ns * MS + ns] = yy;
    st->k[0] = kk;
    st->spare[0] = newspare;
    st->gamma[0] = gc;
    st->gen[0] += 1.0;
    st->pair_ok[0] = 1.0;
    return true;
}

// LSR1.B (lsr1.py:181-198): out = gamma v + Psi (Minv \ Psi^T v), expressed as the
// coefficients (cg, cS, cY) of the output pass.  S^T v, Y^T v are read from Sg, Yg.
KS_FN inline void lsr1_B_coefs(dd4ml_state* st, Scratch* ws) {
    const int k = (int)st->k[0];
    const double gam = st->gamma[0];
    for (int i = 0; i < MS; ++i) { st->cS[i] = 0.0; st->cY[i] = 0.0; }
    st->cg[0] = gam;
    st->converged[0] = 0.0;
    st->solve_valid[0] = 0.0;
    st->err[0] = DD4ML_ERR_NONE;
    if (k == 0) return;
    Compact& C = ws->C;
    build_compact(st, C);
    double* x = ws->x;
    for (int i = 0; i < k; ++i) {
        const int pi = (int)st->order[i];
        x[i] = st->Yg[pi] - gam * st->Sg[pi];
    }
    for (int i = 0; i < k * M; ++i) ws->LU[i] = C.Minv[i];
    if (!lu_factor(ws->LU, k, ws->piv)) { st->err[0] = DD4ML_ERR_SINGULAR; return; }
    lu_solve(ws->LU, ws->piv, k, x);
    for (int i = 0; i < k; ++i) {
        const int pi = (int)st->order[i];
        st->cY[pi] = x[i];
        st->cS[pi] = -gam * x[i];
    }
}

// OBS.ComputeSBySMW (obs.py:175-182) for a caller-supplied tau (read from st->tau):
// p = -g/tau + Psi ((tau Minv + Psi^T Psi) \ Psi^T g), as coefficients of the output pass
// (as coded in the reference, i.e. without the paper's 1/tau on the Psi term).
KS_FN inline void smw_coefs(dd4ml_state* st, Scratch* ws) {
    const int k = (int)st->k[0];
    const double gam = st->gamma[0], tau = st->tau[0];
    for (int i = 0; i < MS; ++i) { st->cS[i] = 0.0; st->cY[i] = 0.0; }
    st->cg[0] = -1.0 / tau;
    st->converged[0] = 0.0;
    st->solve_valid[0] = 0.0;
    st->err[0] = DD4ML_ERR_NONE;
    if (k == 0) return;
    Compact& C = ws->C;
    build_compact(st, C);
    double* x = ws->x;
    for (int i = 0; i < k; ++i) {
        const int pi = (int)st->order[i];
        x[i] = st->Yg[pi] - gam * st->Sg[pi];
    }
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) ws->W[i * M + j] = tau * C.Minv[i * M + j] + C.G[i * M + j];
    if (!lu_factor(ws->W, k, ws->pw)) { st->err[0] = DD4ML_ERR_SINGULAR; return; }
    lu_solve(ws->W, ws->pw, k, x);
    for (int i = 0; i < k; ++i) {
        const int pi = (int)st->order[i];
        st->cY[pi] = x[i];
        st->cS[pi] = -gam * x[i];
    }
}

// rho and the accept decision of TR.step (tr.py:186-191), in the loss dtype T.
template <class T>
KS_FN inline bool tr_accept(double loss, double trial, double pred, double tol, double nu_dec, double* rho_out) {
    if (fabs((double)(T)pred) < tol) {
        *rho_out = INFINITY;
        return true;
    }
    const T rho = ((T)loss - (T)trial) / (T)pred;
    *rho_out = (double)rho;
    return rho > (T)nu_dec;
}

// Radius update + L-SR1 memory update after the trial evaluation (tr.py:191-221).
// `accept` was decided by tr_accept; on acceptance the candidate reductions c_* and
// out_gg / out_ginf (norms of the trial gradient) have been filled by the kernel.
template <class T>
KS_FN inline void tr_finish(dd4ml_state* st, bool accept, double rho, Scratch* ws) {
    st->rho[0] = rho;
    st->accept[0] = accept ? 1.0 : 0.0;
    st->pair_tried[0] = 0.0;
    st->pair_ok[0] = 0.0;
    st->steps[0] += 1.0;
    const double delta = st->delta[0];
    if (accept) {
        if (st->second_order[0] != 0.0) {
            const double tol = st->tol[0];
            if (sqrt(st->c_ss[0]) > tol && sqrt(st->c_yy[0]) > tol) {  // tr.py:202
                st->pair_tried[0] = 1.0;
                lsr1_try_update(st, ws);
            }
        }
        if ((T)rho > (T)st->nu_inc[0] && (T)st->pnorm[0] >= (T)(0.9 * delta)) {  // tr.py:207
            const double d = st->inc_factor[0] * delta;
            st->delta[0] = (st->max_delta[0] < d) ? st->max_delta[0] : d;
        }
        st->out_gn[0] = (double)(T)((st->norm_inf[0] != 0.0) ? st->out_ginf[0] : sqrt(st->out_gg[0]));
    } else {
        const double d = st->dec_factor[0] * delta;
        st->delta[0] = (st->min_delta[0] > d) ? st->min_delta[0] : d;
        st->out_gn[0] = st->gn[0];
        st->out_gg[0] = st->gg[0];
        st->out_ginf[0] = st->ginf[0];
    }
}

// APTS_Base.control_step decision (apts_base.py:483-501): reject when pred < tol or
// rho < nu_dec; enlarge when rho > nu_inc.  delta is the APTS-level radius.
template <class T>
KS_FN inline bool apts_accept(double f0, double ftrial, double pred, double tol, double nu_dec) {
    double rho;
    if (fabs((double)(T)pred) < tol) rho = INFINITY;
    else rho = (double)(((T)f0 - (T)ftrial) / (T)pred);
    return !((T)pred < (T)tol || (T)rho < (T)nu_dec);
}

template <class T>
KS_FN inline bool apts_control(dd4ml_state* st, double f0, double ftrial, double pred) {
    const double tol = st->tol[0];
    double rho;
    if (fabs((double)(T)pred) < tol) rho = INFINITY;
    else rho = (double)(((T)f0 - (T)ftrial) / (T)pred);
    st->rho[0] = rho;
    st->pred[0] = pred;
    st->loss[0] = f0;
    st->trial_loss[0] = ftrial;
    const double delta = st->delta[0];
    if ((T)pred < (T)tol || (T)rho < (T)st->nu_dec[0]) {
        const double d = delta * st->dec_factor[0];
        st->delta[0] = (d > st->min_delta[0]) ? d : st->min_delta[0];
        st->accept[0] = 0.0;
        return false;
    }
    if ((T)rho > (T)st->nu_inc[0]) {
        const double d = delta * st->inc_factor[0];
        st->delta[0] = (d < st->max_delta[0]) ? d : st->max_delta[0];
    }
    st->accept[0] = 1.0;
    return true;
}

}  // namespace ks

#include "kspace_small.h"
