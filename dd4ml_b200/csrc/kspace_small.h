// Register-resident versions of the k-space routines for k <= 5 (KS_SMALL_MAX) stored pairs (the memory
// lengths of the reference's configs are 3 and 5).  Same arithmetic, same order of operations
// as the generic routines in kspace.h, but K is a template parameter: every loop unrolls, every
// matrix lives in registers, pivoting and sorting use predicated compare-exchanges instead of
// runtime indices.  The generic routines work out of shared memory one dependent load at a
// time (~100 cycles per scalar operation in a single thread); these run ~5x faster.
#pragma once
#include "kspace.h"

namespace ks {

template <int K>
KS_FN inline bool chol_upper_s(const double (&G)[K * K], double (&R)[K * K]) {
#pragma unroll
    for (int i = 0; i < K * K; ++i) R[i] = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double s = G[j * K + j];
#pragma unroll
        for (int i = 0; i < j; ++i) s -= R[i * K + j] * R[i * K + j];
        if (!(s > 0.0) || !finite(s)) return false;
        const double d = sqrt(s);
        R[j * K + j] = d;
#pragma unroll
        for (int c = j + 1; c < K; ++c) {
            double t = G[j * K + c];
#pragma unroll
            for (int i = 0; i < j; ++i) t -= R[i * K + j] * R[i * K + c];
            R[j * K + c] = t / d;
        }
    }
    return true;
}

template <int K>
KS_FN inline bool lu_factor_s(double (&A)[K * K], int (&piv)[K]) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
        int p = c;
        double best = fabs(A[c * K + c]);
#pragma unroll
        for (int r = c + 1; r < K; ++r) {
            const double v = fabs(A[r * K + c]);
            if (v > best) { best = v; p = r; }
        }
        piv[c] = p;
        if (best == 0.0 || !finite(best)) return false;
#pragma unroll
        for (int r = c + 1; r < K; ++r)
            if (p == r) {
#pragma unroll
                for (int j = 0; j < K; ++j) { const double t = A[c * K + j]; A[c * K + j] = A[r * K + j]; A[r * K + j] = t; }
            }
        const double inv = 1.0 / A[c * K + c];
#pragma unroll
        for (int r = c + 1; r < K; ++r) {
            const double f = A[r * K + c] * inv;
            A[r * K + c] = f;
#pragma unroll
            for (int j = c + 1; j < K; ++j) A[r * K + j] -= f * A[c * K + j];
        }
    }
    return true;
}

template <int K>
KS_FN inline void lu_solve_s(const double (&LU)[K * K], const int (&piv)[K], double (&b)[K]) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
#pragma unroll
        for (int r = c + 1; r < K; ++r)
            if (piv[c] == r) { const double t = b[c]; b[c] = b[r]; b[r] = t; }
    }
#pragma unroll
    for (int r = 1; r < K; ++r) {
        double s = b[r];
#pragma unroll
        for (int j = 0; j < r; ++j) s -= LU[r * K + j] * b[j];
        b[r] = s;
    }
#pragma unroll
    for (int r = K - 1; r >= 0; --r) {
        double s = b[r];
#pragma unroll
        for (int j = r + 1; j < K; ++j) s -= LU[r * K + j] * b[j];
        b[r] = s / LU[r * K + r];
    }
}

template <int K>
KS_FN inline int jacobi_s(double (&A)[K * K], double (&w)[K], double (&V)[K * K]) {
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) V[i * K + j] = (i == j) ? 1.0 : 0.0;
    int sweeps = 0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            diag += A[i * K + i] * A[i * K + i];
#pragma unroll
            for (int j = i + 1; j < K; ++j) off += A[i * K + j] * A[i * K + j];
        }
        if (off <= 1e-26 * diag || off == 0.0) break;
        ++sweeps;
#pragma unroll
        for (int p = 0; p < K - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < K; ++q) {
                const double apq = A[p * K + q];
                if (apq != 0.0) {
                    const double d = A[q * K + q] - A[p * K + p];
                    const double a2 = 2.0 * apq;
                    const double sg = ((d >= 0.0) == (a2 >= 0.0)) ? 1.0 : -1.0;
                    const double t = sg * fabs(a2) / (fabs(d) + sqrt(d * d + a2 * a2));
                    const double c = KS_RSQRT(t * t + 1.0), s = t * c;
#pragma unroll
                    for (int r = 0; r < K; ++r) {
                        const double arp = A[r * K + p], arq = A[r * K + q];
                        A[r * K + p] = c * arp - s * arq;
                        A[r * K + q] = s * arp + c * arq;
                    }
#pragma unroll
                    for (int r = 0; r < K; ++r) {
                        const double apr = A[p * K + r], aqr = A[q * K + r];
                        A[p * K + r] = c * apr - s * aqr;
                        A[q * K + r] = s * apr + c * aqr;
                    }
#pragma unroll
                    for (int r = 0; r < K; ++r) {
                        const double vrp = V[r * K + p], vrq = V[r * K + q];
                        V[r * K + p] = c * vrp - s * vrq;
                        V[r * K + q] = s * vrp + c * vrq;
                    }
                }
            }
    }
#pragma unroll
    for (int i = 0; i < K; ++i) w[i] = A[i * K + i];
    // ascending order by a fixed compare-exchange network (bubble passes)
#pragma unroll
    for (int pass = 0; pass < K - 1; ++pass)
#pragma unroll
        for (int j = 0; j < K - 1 - pass; ++j)
            if (w[j + 1] < w[j]) {
                const double t = w[j]; w[j] = w[j + 1]; w[j + 1] = t;
#pragma unroll
                for (int r = 0; r < K; ++r) { const double u = V[r * K + j]; V[r * K + j] = V[r * K + j + 1]; V[r * K + j + 1] = u; }
            }
    // sign convention: largest-|.| component of every eigenvector positive (lowest index on ties)
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double best = -1.0, val = 0.0;
#pragma unroll
        for (int r = 0; r < K; ++r)
            if (fabs(V[r * K + j]) > best) { best = fabs(V[r * K + j]); val = V[r * K + j]; }
        if (val < 0.0) {
#pragma unroll
            for (int r = 0; r < K; ++r) V[r * K + j] = -V[r * K + j];
        }
    }
    return sweeps;
}


// secular equation with the element count M1 = K+1 known at compile time (registers)
template <class T, int M1>
KS_FN inline void phibar_fg_s(T sigma, const T (&Lam)[M1], const T (&a)[M1], T delta, T tol, T& f, T& g, bool& deg) {
    deg = false;
#pragma unroll
    for (int j = 0; j < M1; ++j) {
        const T D = Lam[j] + sigma;
        if (fabs(a[j]) < tol || fabs(D) < tol) deg = true;
    }
    if (deg) {
        f = T(-1) / delta;
        g = T(1000000.0);
        return;
    }
    T n2 = T(0), s3 = T(0);
#pragma unroll
    for (int j = 0; j < M1; ++j) {
        const T D = Lam[j] + sigma;
        const T p = a[j] / D;
        n2 += p * p;
        s3 += (a[j] * a[j]) / (D * D * D);
    }
    const T nrm = ks_sqrt(n2);
    f = T(1) / nrm - T(1) / delta;
    g = s3 / (nrm * nrm * nrm);
}

template <class T, int M1>
KS_FN inline T phibar_f_s(T sigma, const T (&Lam)[M1], const T (&a)[M1], T delta, T tol) {
    bool deg = false;
#pragma unroll
    for (int j = 0; j < M1; ++j) {
        const T D = Lam[j] + sigma;
        if (fabs(a[j]) < tol || fabs(D) < tol) deg = true;
    }
    if (deg) return T(-1) / delta;
    T n2 = T(0);
#pragma unroll
    for (int j = 0; j < M1; ++j) {
        const T p = a[j] / (Lam[j] + sigma);
        n2 += p * p;
    }
    return T(1) / ks_sqrt(n2) - T(1) / delta;
}

template <class T, int M1>
KS_FN inline T newton_s(T x0, const T (&Lam)[M1], const T (&a)[M1], T delta, T tol, int& iters, int& ndeg) {
    T x = x0, f, g;
    bool deg;
    bool adeg = false;
#pragma unroll
    for (int j = 0; j < M1; ++j) adeg = adeg || (fabs(a[j]) < tol);
    if (adeg) {   // see newton() in kspace.h
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
    phibar_fg_s<T, M1>(x, Lam, a, delta, tol, f, g, deg);
    iters = 0;
    ndeg = deg ? 1 : 0;
    while (fabs(f) > tol && iters < 200) {
        x = x - f / g;
        phibar_fg_s<T, M1>(x, Lam, a, delta, tol, f, g, deg);
        ndeg += deg ? 1 : 0;
        ++iters;
    }
    return x;
}

template <int K>
struct CompactS {
    double gamma;
    double Minv[K * K], G[K * K];
};

template <int K>
KS_FN inline void build_compact_s(const dd4ml_state* st, CompactS<K>& c) {
    const double gam = st->gamma[0];
    c.gamma = gam;
    int ord[K];
#pragma unroll
    for (int i = 0; i < K; ++i) ord[i] = (int)st->order[i];
    double SS[K * K], SY[K * K], YY[K * K];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            SS[i * K + j] = st->SS[ord[i] * MS + ord[j]];
            SY[i * K + j] = st->SY[ord[i] * MS + ord[j]];
            YY[i * K + j] = st->YY[ord[i] * MS + ord[j]];
        }
    double fro = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double sy = (i >= j) ? SY[i * K + j] : SY[j * K + i];
            const double v = sy - gam * SS[i * K + j];
            c.Minv[i * K + j] = v;
            fro += v * v;
            c.G[i * K + j] = YY[i * K + j] - gam * (SY[i * K + j] + SY[j * K + i]) + gam * gam * SS[i * K + j];
        }
    const double reg = st->tol[0] * sqrt(fro);
#pragma unroll
    for (int i = 0; i < K; ++i) c.Minv[i * K + i] += reg;
}

template <int K>
KS_FN inline double dot_s(const double (&a)[K], const double (&b)[K]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) s += a[i] * b[i];
    return s;
}
template <int K>
KS_FN inline double quad_s(const double (&A)[K * K], const double (&x)[K]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) s += x[i] * A[i * K + j] * x[j];
    return s;
}

// second-order part of tr_solve for exactly K stored pairs.  Returns false on a numerical
// failure (st->err set).  Outputs the step in k-space (cg, c[]) and pred, p^T p, g^T p.
template <class T, int K>
KS_FN
#ifdef __CUDACC__
__noinline__
#endif
bool second_order_s(dd4ml_state* st, double gn, double gg, double delta, double& cg_out, double (&c_out)[M],
                    double& pred, double& pp, double& gp) {
    CompactS<K> C;
    build_compact_s<K>(st, C);
    const double gam = C.gamma;
    double b[K];
    bool ok = finite(gg) && finite(delta) && finite(gam);
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int pi = (int)st->order[i];
        b[i] = st->Yg[pi] - gam * st->Sg[pi];
        ok = ok && finite(b[i]);
#pragma unroll
        for (int j = 0; j < K; ++j) ok = ok && finite(C.G[i * K + j]) && finite(C.Minv[i * K + j]);
    }
    if (!ok) { st->err[0] = DD4ML_ERR_NONFINITE; return false; }
    if (delta < 0) { st->err[0] = DD4ML_ERR_NEG_DELTA; return false; }
    double R[K * K], LU[K * K], RMR[K * K], U[K * K], D[K], RinvU[K * K];
    int piv[K];
    if (!chol_upper_s<K>(C.G, R)) { st->err[0] = DD4ML_ERR_CHOLESKY; return false; }
#pragma unroll
    for (int i = 0; i < K * K; ++i) LU[i] = C.Minv[i];
    if (!lu_factor_s<K>(LU, piv)) { st->err[0] = DD4ML_ERR_SINGULAR; return false; }
    {
        double MR[K * K];
#pragma unroll
        for (int c = 0; c < K; ++c) {
            double col[K];
#pragma unroll
            for (int r = 0; r < K; ++r) col[r] = R[c * K + r];
            lu_solve_s<K>(LU, piv, col);
#pragma unroll
            for (int r = 0; r < K; ++r) MR[r * K + c] = col[r];
        }
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < K; ++r) s += R[i * K + r] * MR[r * K + j];
                RMR[i * K + j] = s;
            }
    }
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = i + 1; j < K; ++j) {
            const double s = 0.5 * (RMR[i * K + j] + RMR[j * K + i]);
            RMR[i * K + j] = RMR[j * K + i] = s;
        }
    const double c_e0 = KS_CLOCK();
    st->jacobi_sweeps[0] = jacobi_s<K>(RMR, D, U);
    st->cyc_eigh[0] = KS_CLOCK() - c_e0;
    T Lam[K + 1], a[K + 1];
    const T tolT = (T)OBS_TOL, gamT = (T)gam, delT = (T)delta;
#pragma unroll
    for (int i = 0; i < K; ++i) Lam[i] = (T)D[i] + gamT;
    Lam[K] = gamT;
#pragma unroll
    for (int i = 0; i <= K; ++i)
        if (fabs(Lam[i]) < tolT) Lam[i] = T(0);
    const T lam_min = (Lam[0] < gamT) ? Lam[0] : gamT;
    st->lam_min[0] = (double)lam_min;
#pragma unroll
    for (int c = 0; c < K; ++c) {
#pragma unroll
        for (int r = K - 1; r >= 0; --r) {
            double s = U[r * K + c];
#pragma unroll
            for (int j = r + 1; j < K; ++j) s -= R[r * K + j] * RinvU[j * K + c];
            RinvU[r * K + c] = s / R[r * K + r];
        }
    }
    double gpgp = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < K; ++r) s += RinvU[r * K + i] * b[r];
        a[i] = (T)s;
        gpgp += s * s;
    }
    {
        const double d = gg - gpgp;
        a[K] = (T)sqrt(d > 0.0 ? d : 0.0);
    }
    T tauT;
    T h2 = T(0);
#pragma unroll
    for (int i = 0; i <= K; ++i) { const T h = a[i] / Lam[i]; h2 += h * h; }
    if (lam_min > T(0) && ks_sqrt(h2) <= delT) {
        tauT = gamT;
        st->case_id[0] = DD4ML_CASE_A;
    } else {
        if (lam_min <= T(0) && phibar_f_s<T, K + 1>(-lam_min, Lam, a, delT, tolT) >= T(0)) {
            st->err[0] = DD4ML_ERR_HARD_CASE;
            return false;
        }
        T x0 = T(0);
        if (!(lam_min > T(0))) {
            T sh = a[0] / delT - Lam[0];
#pragma unroll
            for (int i = 1; i <= K; ++i) { const T v = a[i] / delT - Lam[i]; if (v > sh) sh = v; }
            x0 = (sh > -lam_min) ? sh : -lam_min;
        }
        int it, nd;
        const double c_n0 = KS_CLOCK();
        T sig = newton_s<T, K + 1>(x0, Lam, a, delT, tolT, it, nd);
        if (!finite((double)sig)) sig = newton_s<T, K + 1>(T(0), Lam, a, delT, tolT, it, nd);
        st->cyc_newton[0] = KS_CLOCK() - c_n0;
        st->sigma0[0] = (double)x0;
        st->sigma[0] = (double)sig;
        st->newton_iters[0] = it;
        st->newton_deg[0] = nd;
        st->case_id[0] = DD4ML_CASE_C;
        tauT = gamT + sig;
    }
    const double tau = (double)tauT;
    st->tau[0] = tau;
    double wv[K];
    {
        double W[K * K];
        int pw[K];
#pragma unroll
        for (int i = 0; i < K * K; ++i) W[i] = tau * C.Minv[i] + C.G[i];
        if (!lu_factor_s<K>(W, pw)) { st->err[0] = DD4ML_ERR_SINGULAR; return false; }
#pragma unroll
        for (int i = 0; i < K; ++i) wv[i] = b[i];
        lu_solve_s<K>(W, pw, wv);
    }
    const double cgb = (double)(T(-1) / tauT);
    double x[K];
#pragma unroll
    for (int i = 0; i < K; ++i) x[i] = b[i];
    lu_solve_s<K>(LU, piv, x);
    const double gBg = gam * gg + dot_s<K>(b, x);
    st->gBg[0] = gBg;
    double cg = cgb, c[K];
#pragma unroll
    for (int i = 0; i < K; ++i) c[i] = wv[i];
    auto ppf = [&](double cgv, const double (&cv)[K]) {
        return cgv * cgv * gg + 2.0 * cgv * dot_s<K>(cv, b) + quad_s<K>(C.G, cv);
    };
    auto gpf = [&](double cgv, const double (&cv)[K]) { return cgv * gg + dot_s<K>(cv, b); };
    if (st->dogleg[0] != 0.0 && gBg > 0.0) {
        const double alpha = gn * gn / gBg;
        const double cc = (alpha * sqrt(gg) >= delta) ? -(delta / gn) : -alpha;
        const double pb_norm = sqrt(ppf(cgb, wv));
        const double pc_norm = fabs(cc) * sqrt(gg);
        if (pb_norm <= delta) {
            st->dogleg_branch[0] = 1.0;
        } else if (pc_norm >= delta) {
            cg = cc;
#pragma unroll
            for (int i = 0; i < K; ++i) c[i] = 0.0;
            st->dogleg_branch[0] = 2.0;
        } else {
            const double dcg = cgb - cc;
            const double qa = ppf(dcg, wv);
            const double qb = 2.0 * cc * gpf(dcg, wv);
            const double qc = cc * cc * gg - delta * delta;
            const double disc = qb * qb - 4.0 * qa * qc;
            if (disc < 0.0) {
                cg = cc;
#pragma unroll
                for (int i = 0; i < K; ++i) c[i] = 0.0;
                st->dogleg_branch[0] = 3.0;
            } else {
                const double t = (-qb + sqrt(disc)) / (2.0 * qa);
                cg = cc + t * dcg;
#pragma unroll
                for (int i = 0; i < K; ++i) c[i] = t * wv[i];
                st->dogleg_branch[0] = 4.0;
            }
        }
    }
    pp = ppf(cg, c);
    gp = gpf(cg, c);
    double q[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double s = cg * b[i];
#pragma unroll
        for (int j = 0; j < K; ++j) s += C.G[i * K + j] * c[j];
        q[i] = s;
    }
    double y[K];
#pragma unroll
    for (int i = 0; i < K; ++i) y[i] = q[i];
    lu_solve_s<K>(LU, piv, y);
    const double pBp = gam * pp + dot_s<K>(q, y);
    pred = -(gp + 0.5 * pBp);
    cg_out = cg;
#pragma unroll
    for (int i = 0; i < K; ++i) c_out[i] = c[i];
    return true;
}

// u = y - B s of the update_memory chain for exactly K stored pairs: returns (u.s, ||u||^2)
template <int K>
KS_FN
#ifdef __CUDACC__
__noinline__
#endif
bool update_us_uu_s(dd4ml_state* st, double ss, double sy, double yy, double& us, double& uu) {
    CompactS<K> C;
    build_compact_s<K>(st, C);
    const double gam = C.gamma;
    double bs[K], by[K], x[K], LU[K * K];
    int piv[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int pi = (int)st->order[i];
        bs[i] = st->c_Ys[pi] - gam * st->c_Ss[pi];
        by[i] = st->c_Yy[pi] - gam * st->c_Sy[pi];
        x[i] = bs[i];
    }
#pragma unroll
    for (int i = 0; i < K * K; ++i) LU[i] = C.Minv[i];
    if (!lu_factor_s<K>(LU, piv)) { st->err[0] = DD4ML_ERR_SINGULAR; return false; }
    lu_solve_s<K>(LU, piv, x);
    us = sy - gam * ss - dot_s<K>(bs, x);
    uu = yy + gam * gam * ss + quad_s<K>(C.G, x) - 2.0 * gam * sy - 2.0 * dot_s<K>(by, x) + 2.0 * gam * dot_s<K>(bs, x);
    return true;
}

}  // namespace ks
