// k-space routines of the trust-region step: everything that happens between the streaming
// reductions over the n-vectors and the streaming output pass.
//
// All inputs are O(k^2) numbers (Gram blocks of the L-SR1 pairs, Psi^T g, g^T g) held in the device
// state block.  The routines run in the ONE-WARP k-space kernel (k_kspace.cu) that follows every
// reduction kernel ("small m x m inner solve kept in one CTA"): the k x k matrices live in shared
// memory and every O(k^2)/O(k^3) step is a *parallel region* over matrix elements / rows / columns /
// Jacobi rotation pairs, so the dependent chain is O(k) steps per factorisation instead of O(k^3)
// scalar operations in one thread.  The code is loop based and compact on purpose: it is executed
// once per launch right after the closure's kernels have evicted the instruction cache, so its
// footprint, not its flop count, sets its latency.
//
// Execution model (also what lets tests/hostshim compile this file for the CPU):
//   KS_PAR(i, n) { body }  KS_SYNC();      a parallel region: iterations are independent -- an iteration
//                                          never reads what another iteration of the same region writes
//   everything else                        uniform code, executed redundantly by all lanes; shared state is
//                                          written by the master lane only (KS_MASTER) and never read back
//                                          before the next KS_SYNC()
// On the device KS_PAR strides the lanes of the warp and KS_SYNC is __syncwarp(); on the host KS_PAR is a
// plain loop and KS_SYNC a no-op, which is the same program by the independence rule.
//
// Reference behaviour restated here (cruzas/DD4ML):
//   LSR1.precompute / B / update_memory      optimizers/lsr1.py:53-198
//   OBS.solve_tr_subproblem / Newton / SMW   solvers/obs.py:46-254
//   solve_tr_first_order / second_order      utility/optimizer_utils.py:197-307
//   TR.step ratio test and radius update     optimizers/tr.py:186-221
//   LSSR1_TR clip-to-radius / descent flip   optimizers/lssr1_tr.py:563-587
// Linear algebra (Cholesky, LU, symmetric eigen-decomposition) runs in double; the quantities the
// reference thresholds or iterates on in its working dtype (Lambda, a_j, the Newton iteration, rho) are
// rounded to that dtype T first.
#pragma once
#include <math.h>

#include "state.h"

#ifdef __CUDACC__
#define KS_FN __host__ __device__
#define KS_BIG __host__ __device__ __noinline__    // one copy of every large routine: code size is latency here
#define KS_LOOP _Pragma("unroll 1")
#else
#define KS_FN
#define KS_BIG
#define KS_LOOP
#endif
#if defined(__CUDA_ARCH__)
#define KS_CLOCK() ((double)clock64())
#define KS_RSQRT(x) rsqrt(x)
#define KS_TID ((int)threadIdx.x)
#define KS_NT ((int)blockDim.x)
#define KS_SYNC() __syncwarp()
#else
#define KS_CLOCK() 0.0
#define KS_RSQRT(x) (1.0 / sqrt(x))
#define KS_TID 0
#define KS_NT 1
#define KS_SYNC() ((void)0)
#endif
#define KS_PAR(i, n) KS_LOOP for (int i = KS_TID; i < (n); i += KS_NT)
#define KS_MASTER if (KS_TID == 0)
// 2-D region over a k x k block: slots e = i*M + j with the constant leading dimension M (division by a
// constant, no runtime integer division), entries j >= k idle.  Usage: KS_PAR2(i, j, k) { ... } KS_END2
#define KS_PAR2(i, j, k) KS_LOOP for (int e_ = KS_TID; e_ < (k) * M; e_ += KS_NT) { const int i = e_ / M, j = e_ - i * M; if (j < (k))
#define KS_END2 }
#ifdef __CUDACC__
#define KS_UNROLL _Pragma("unroll")
#else
#define KS_UNROLL
#endif

namespace ks {

constexpr int M = DD4ML_MAXM;
constexpr int MS = DD4ML_MAXS;
constexpr double OBS_TOL = 1e-6;  // obs.py:29

KS_FN inline bool finite(double x) { return x == x && fabs(x) <= 1.79769313486231570e308; }
// IEEE double division / square root expand to ~30 instructions each: one out-of-line copy
KS_BIG inline double ddiv(double a, double b) { return a / b; }
KS_BIG inline double dsqrt(double x) { return sqrt(x); }
// Round to the reference's working dtype (float32 or float64).  A float32 operation on float32
// operands equals the float64 operation rounded to float32 (53 >= 2*24 + 2 bits: no double-rounding
// error for + - * / sqrt), so `rT(a op b, f32)` IS the reference's float32 arithmetic.
KS_FN inline double rT(double x, bool f32) { return f32 ? (double)(float)x : x; }
KS_FN inline float ks_sqrt(float x) { return sqrtf(x); }
KS_FN inline double ks_sqrt(double x) { return sqrt(x); }

// Work arrays of the k-space routines: shared memory of the k-space kernel (row-major, leading
// dimension M).
struct Scratch {
    double Minv[M * M];  // diag(S^T Y) + L + L^T - gamma S^T S + tol*||.||_F I   (lsr1.py:168-179)
    double G[M * M];     // Psi^T Psi
    double R[M * M], LU[M * M], MR[M * M], RMR[M * M], U[M * M], V[M * M], RinvU[M * M], W[M * M];
    double D[M], Dw[M], b[M], x[M], z[M], wv[M], Gw[M], bs[M], by[M], ad[M + 1];
    double Ld[M + 1], Ad[M + 1];   // Lambda, a in the working dtype (T = double)
    float Lf[M + 1], Af[M + 1];    //                                 (T = float)
    double rc[M], rs[M];           // Jacobi rotations of one round
    int rp[M], rq[M], rank[M];
    int piv[M], pw[M];
};
template <class T> struct LamA;
template <> struct LamA<float> {
    KS_FN static float* lam(Scratch* ws) { return ws->Lf; }
    KS_FN static float* a(Scratch* ws) { return ws->Af; }
};
template <> struct LamA<double> {
    KS_FN static double* lam(Scratch* ws) { return ws->Ld; }
    KS_FN static double* a(Scratch* ws) { return ws->Ad; }
};

// sum_{i<k} a[i*sa] * b[i*sb].  ONE out-of-line copy with a runtime loop: the k-space kernel is bound by
// instruction issue / fetch of a single warp (ncu: ~5 cycles per instruction, half of the stall samples
// `no_instruction`), so the dynamic instruction count and the code footprint are what matter -- a
// predicated unroll to M terms executes M/k times the instructions and was measured slower.
KS_BIG inline double dots(const double* a, int sa, const double* b, int sb, int k) {
    double s = 0.0;
    KS_LOOP
    for (int i = 0; i < k; ++i) s += a[i * sa] * b[i * sb];
    return s;
}
KS_FN inline double dotk(const double* a, const double* b, int k) { return dots(a, 1, b, 1, k); }

// ---- dense helpers on row-major k x k blocks with leading dimension M (collective) -----------
// R upper triangular with R^T R = G (torch.linalg.cholesky(upper=True)).  One column per step:
// the pivot is uniform scalar work, the row to its right a parallel region.
KS_BIG inline bool chol_upper(const double* G, int k, double* R) {
    KS_PAR(e, k * M) R[e] = 0.0;
    KS_SYNC();
    KS_LOOP
    for (int j = 0; j < k; ++j) {
        const double s = G[j * M + j] - dots(R + j, M, R + j, M, j);
        if (!(s > 0.0) || !finite(s)) return false;
        const double d = dsqrt(s);
        KS_PAR(c, k) {
            if (c == j) R[j * M + j] = d;
            else if (c > j) R[j * M + c] = ddiv(G[j * M + c] - dots(R + j, M, R + c, M, j), d);
        }
        KS_SYNC();
    }
    return true;
}

// In-place LU with partial pivoting (torch.linalg.solve = getrf/getrs): pivot search uniform; row
// swap, multipliers and the rank-1 update of the trailing block are parallel regions.
KS_BIG inline bool lu_factor(double* A, int k, int* piv) {
    KS_LOOP
    for (int c = 0; c < k; ++c) {
        int p = c;
        double best = fabs(A[c * M + c]);
        KS_LOOP
        for (int r = c + 1; r < k; ++r) {
            const double v = fabs(A[r * M + c]);
            if (v > best) { best = v; p = r; }
        }
        if (best == 0.0 || !finite(best)) return false;
        KS_SYNC();                      // every lane has read column c before rows move
        KS_PAR(j, k) {
            if (j == 0) piv[c] = p;
            if (p != c) { const double t = A[c * M + j]; A[c * M + j] = A[p * M + j]; A[p * M + j] = t; }
        }
        KS_SYNC();
        const double inv = ddiv(1.0, A[c * M + c]);
        KS_PAR(r, k) { if (r > c) A[r * M + c] *= inv; }          // multipliers
        KS_SYNC();
        KS_PAR2(r, j, k) {              // trailing block
            if (r > c && j > c) A[r * M + j] -= A[r * M + c] * A[c * M + j];
        } KS_END2
        KS_SYNC();
    }
    return true;
}

// v <- A^{-1} v from the factorisation (per lane, not collective); v has stride ldv
KS_BIG inline void lu_solve_vec(const double* LU, const int* piv, int k, double* v, int ldv) {
    KS_LOOP
    for (int i = 0; i < k; ++i) {
        const int p = piv[i];
        if (p != i) { const double t = v[i * ldv]; v[i * ldv] = v[p * ldv]; v[p * ldv] = t; }
    }
    KS_LOOP
    for (int r = 1; r < k; ++r) v[r * ldv] -= dots(LU + r * M, 1, v, ldv, r);
    KS_LOOP
    for (int r = k - 1; r >= 0; --r) {
        const double s = v[r * ldv] - dots(LU + r * M + r + 1, 1, v + (r + 1) * ldv, ldv, k - r - 1);
        v[r * ldv] = ddiv(s, LU[r * M + r]);
    }
}

// X[:, c] <- A^{-1} X[:, c] for c < ncols (one right-hand side per lane).  Collective.
KS_FN inline void lu_solve_cols(const double* LU, const int* piv, int k, double* X, int ldx, int ncols) {
    KS_PAR(c, ncols) lu_solve_vec(LU, piv, k, X + c, ldx);
    KS_SYNC();
}

// Jacobi eigen-decomposition of a symmetric k x k matrix with the round-robin (tournament) ordering:
// the floor(k/2) rotations of a round touch disjoint index pairs and are applied together -- k-1 (k)
// rounds per sweep instead of k(k-1)/2 dependent rotations.  On exit w ascending, U[:, j] the
// eigenvector of w[j], each with its largest-|.| component positive (lowest index on ties): the
// documented sign convention of this implementation.  A is destroyed.
// The rotation angle is computed in float32 on exponent-normalised inputs (any angle keeps the
// transformation orthogonal as long as c = 1/sqrt(1 + t^2) and s = t c are formed in double; a 1e-7
// relative error of t only leaves 1e-7 of the pivot behind for the next sweep), which takes the
// double-precision division and square root off the dependent chain of every round.
KS_BIG inline int jacobi_eigh(double* A, int k, double* w, double* U, Scratch* ws) {
    double* V = ws->V;
    KS_PAR2(i, j, k) { V[i * M + j] = (i == j) ? 1.0 : 0.0; } KS_END2
    KS_SYNC();
    const int n = k + (k & 1);          // players of the tournament (one dummy when k is odd)
    const int half = n >> 1;
    int sweeps = 0;
    KS_LOOP
    for (int sweep = 0; sweep < 64 && k > 1; ++sweep) {
        KS_PAR(i, k) {                  // off-diagonal / diagonal mass, one row per lane
            double o = 0.0;         // (summed directly: total - diagonal would cancel at the 1e-26 threshold)
            KS_LOOP
            for (int j = 0; j < k; ++j)
                if (j != i) o += A[i * M + j] * A[i * M + j];
            ws->rc[i] = o;
            ws->rs[i] = A[i * M + i] * A[i * M + i];
        }
        KS_SYNC();
        double off = 0.0, diag = 0.0;
        KS_LOOP
        for (int i = 0; i < k; ++i) { off += ws->rc[i]; diag += ws->rs[i]; }
        KS_SYNC();                      // rc / rs are reused for the rotations
        // off-diagonal mass (each entry counted twice) below 1e-26 of the diagonal mass: eigenvalues
        // converged to ~1e-26 relative (Jacobi converges quadratically), eigenvectors to ~1e-13
        if (0.5 * off <= 1e-26 * diag || off == 0.0) break;
        ++sweeps;
        KS_LOOP
        for (int round = 0; round < n - 1; ++round) {
            KS_PAR(t, half) {           // rotation of pair t of this round
                int p, q;
                if (t == 0) { p = n - 1; q = round; }
                else {
                    p = round + t; if (p >= n - 1) p -= n - 1;
                    q = round - t; if (q < 0) q += n - 1;
                }
                if (p > q) { const int u = p; p = q; q = u; }
                double c = 1.0, s = 0.0;
                if (q < k) {
                    const double apq = A[p * M + q];
                    if (apq != 0.0) {
                        // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (aqq - app) / (2 apq)
                        const double d = A[q * M + q] - A[p * M + p];
                        const double a2 = 2.0 * apq;
                        int ex;
                        (void)frexp(fmax(fabs(d), fabs(a2)), &ex);
                        const float df = (float)ldexp(d, -ex), af = (float)ldexp(a2, -ex);
                        const float tf = fabsf(af) / (fabsf(df) + sqrtf(df * df + af * af));
                        const double tt = ((d >= 0.0) == (a2 >= 0.0)) ? (double)tf : -(double)tf;
                        c = KS_RSQRT(tt * tt + 1.0);
                        s = tt * c;
                    }
                }
                ws->rp[t] = p; ws->rq[t] = q; ws->rc[t] = c; ws->rs[t] = s;
            }
            KS_SYNC();
            KS_PAR2(t, r, k) {          // A <- A J, V <- V J: entry pairs (r, p), (r, q) of every row r
                if (t < half) {
                    const int p = ws->rp[t], q = ws->rq[t];
                    if (q < k) {
                        const double c = ws->rc[t], s = ws->rs[t];
                        const double arp = A[r * M + p], arq = A[r * M + q];
                        const double vrp = V[r * M + p], vrq = V[r * M + q];
                        A[r * M + p] = c * arp - s * arq;
                        A[r * M + q] = s * arp + c * arq;
                        V[r * M + p] = c * vrp - s * vrq;
                        V[r * M + q] = s * vrp + c * vrq;
                    }
                }
            } KS_END2
            KS_SYNC();
            KS_PAR2(t, r, k) {          // A <- J^T A: entry pairs (p, r), (q, r) of every column r
                if (t < half) {
                    const int p = ws->rp[t], q = ws->rq[t];
                    if (q < k) {
                        const double c = ws->rc[t], s = ws->rs[t];
                        const double apr = A[p * M + r], aqr = A[q * M + r];
                        A[p * M + r] = c * apr - s * aqr;
                        A[q * M + r] = s * apr + c * aqr;
                    }
                }
            } KS_END2
            KS_SYNC();
        }
    }
    // ascending order through ranks (stable on ties), then the sign convention per column
    KS_PAR(j, k) {
        const double wj = A[j * M + j];
        int r = 0;
        KS_LOOP
        for (int i = 0; i < k; ++i) {
            const double wi = A[i * M + i];
            r += (wi < wj || (wi == wj && i < j)) ? 1 : 0;
        }
        ws->rank[j] = r;
        w[r] = wj;
    }
    KS_SYNC();
    KS_PAR(j, k) {
        const int r = ws->rank[j];
        int im = 0;
        double best = -1.0;
        KS_LOOP
        for (int i = 0; i < k; ++i)
            if (fabs(V[i * M + j]) > best) { best = fabs(V[i * M + j]); im = i; }
        const double sg = (V[im * M + j] < 0.0) ? -1.0 : 1.0;
        KS_LOOP
        for (int i = 0; i < k; ++i) U[i * M + r] = sg * V[i * M + j];
    }
    KS_SYNC();
    return sweeps;
}

// ---- secular equation, in the reference's working dtype T (obs.py:184-254); per-lane, pure ----
template <class T>
struct Secular {
    const T* L;
    const T* A;
    int m;
    T delta, tol;
    KS_FN void load(const T* Lam, const T* a, int m_, T delta_, T tol_) { L = Lam; A = a; m = m_; delta = delta_; tol = tol_; }
    KS_FN bool degenerate(T sigma) const {
        bool deg = false;
        KS_LOOP
        for (int j = 0; j < m; ++j) {
            const T D = L[j] + sigma;
            if (fabs(A[j]) < tol || fabs(D) < tol) deg = true;
        }
        return deg;
    }
    KS_FN void fg(T sigma, T& f, T& g, bool& deg) const {
        deg = degenerate(sigma);
        if (deg) {  // obs.py:242-245
            f = T(-1) / delta;
            g = T(1000000.0);
            return;
        }
        T n2 = T(0), s3 = T(0);
        KS_LOOP
        for (int j = 0; j < m; ++j) {
            const T D = L[j] + sigma;
            const T p = A[j] / D;
            n2 += p * p;
            s3 += (A[j] * A[j]) / (D * D * D);
        }
        const T nrm = ks_sqrt(n2);
        f = T(1) / nrm - T(1) / delta;
        g = s3 / (nrm * nrm * nrm);
    }
    KS_FN T f_only(T sigma) const {
        if (degenerate(sigma)) return T(-1) / delta;
        T n2 = T(0);
        KS_LOOP
        for (int j = 0; j < m; ++j) {
            const T p = A[j] / (L[j] + sigma);
            n2 += p * p;
        }
        return T(1) / ks_sqrt(n2) - T(1) / delta;
    }
};

template <class T>
KS_BIG inline void phibar_fg(T sigma, const T* Lam, const T* a, int m, T delta, T tol, T& f, T& g, bool& deg) {
    Secular<T> sec;
    sec.load(Lam, a, m, delta, tol);
    sec.fg(sigma, f, g, deg);
}

template <class T>
KS_BIG inline T phibar_f(T sigma, const T* Lam, const T* a, int m, T delta, T tol) {
    Secular<T> sec;
    sec.load(Lam, a, m, delta, tol);
    return sec.f_only(sigma);
}

template <class T>
KS_BIG inline T newton(T x0, const T* Lam, const T* a, int m, T delta, T tol, int& iters, int& ndeg) {
    Secular<T> sec;
    sec.load(Lam, a, m, delta, tol);
    T x = x0, f, g;
    bool deg;
    bool adeg = false;
    KS_LOOP
    for (int j = 0; j < m; ++j) adeg = adeg || (fabs(sec.A[j]) < tol);
    if (adeg) {
        // some |a_j| < tol: phibar_fg returns (-1/delta, 1/tol) for EVERY sigma (obs.py:242-245),
        // so the reference's loop is sigma <- sigma - f/g, 200 times (or not at all when
        // |f| <= tol).  Same fp32/fp64 rounding sequence, without re-evaluating phibar.
        f = T(-1) / delta;
        g = T(1000000.0);
        iters = 0;
        ndeg = 1;
        const T inc = f / g;
        if (fabs(f) > tol) {
            KS_LOOP
            for (; iters < 200; ++iters) x = x - inc;
        }
        ndeg += iters;
        return x;
    }
    sec.fg(x, f, g, deg);
    iters = 0;
    ndeg = deg ? 1 : 0;
    KS_LOOP
    while (fabs(f) > tol && iters < 200) {  // obs.py:217
        x = x - f / g;
        sec.fg(x, f, g, deg);
        ndeg += deg ? 1 : 0;
        ++iters;
    }
    return x;
}

// ---- logical (oldest..newest) views of the Gram blocks: Minv and G = Psi^T Psi (collective) ------
KS_BIG inline void build_compact(const dd4ml_state* st, Scratch* ws) {
    const int k = (int)st->k[0];
    const double gam = st->gamma[0];
    KS_PAR2(i, j, k) {
        const int pi = (int)st->order[i], pj = (int)st->order[j];
        const double sy_ij = st->SY[pi * MS + pj], sy_ji = st->SY[pj * MS + pi];
        const double ss = st->SS[pi * MS + pj], yy = st->YY[pi * MS + pj];
        ws->Minv[i * M + j] = ((i >= j) ? sy_ij : sy_ji) - gam * ss;
        ws->G[i * M + j] = yy - gam * (sy_ij + sy_ji) + gam * gam * ss;
        if (st->use_mx[0] != 0.0) ws->Minv[i * M + j] = st->Mx[i * M + j];   // explicit Minv (already regularised)
    } KS_END2
    KS_SYNC();
    if (st->use_mx[0] != 0.0) return;
    KS_PAR(i, k) ws->Dw[i] = dots(ws->Minv + i * M, 1, ws->Minv + i * M, 1, k);   // row sums of squares
    KS_SYNC();
    double fro = 0.0;
    KS_LOOP
    for (int i = 0; i < k; ++i) fro += ws->Dw[i];
    const double reg = st->tol[0] * dsqrt(fro);
    KS_SYNC();                          // every lane has summed the unregularised matrix
    KS_PAR(i, k) ws->Minv[i * M + i] += reg;
    KS_SYNC();
}

// out = A v for a row-major k x k block (collective: one row per lane)
KS_FN inline void matvec(const double* A, const double* v, int k, double* out) {
    KS_PAR(i, k) out[i] = dots(A + i * M, 1, v, 1, k);
    KS_SYNC();
}

// =================================================================================
// The trust-region sub-problem.  mode 0: TR (tr.py:150-176), mode 1: LSSR1_TR
// (additionally clip to the radius and flip to a descent direction,
// lssr1_tr.py:563-587).  Reads gg, ginf, Sg, Yg and the L-SR1 blocks from the
// state; writes the step descriptor (cg, cS, cY) and the report fields.  Collective.
// Every step is p = cg g + lam Psi wv, so p^T p, g^T p and p^T B p are quadratic forms in
// (cg, lam) with coefficients g^T g, wv^T Psi^T g, wv^T G wv, ... computed once.
// =================================================================================
KS_BIG inline double newton_rt(bool f32, double x0, Scratch* ws, int m, double delta, double tol, int& it, int& nd) {
    if (f32) return (double)newton<float>((float)x0, ws->Lf, ws->Af, m, (float)delta, (float)tol, it, nd);
    return newton<double>(x0, ws->Ld, ws->Ad, m, delta, tol, it, nd);
}
KS_BIG inline double phibar_f_rt(bool f32, double sigma, Scratch* ws, int m, double delta, double tol) {
    if (f32) return (double)phibar_f<float>((float)sigma, ws->Lf, ws->Af, m, (float)delta, (float)tol);
    return phibar_f<double>(sigma, ws->Ld, ws->Ad, m, delta, tol);
}

// `f32`: the reference's working dtype T is float32 (else float64).  Quantities that are T values in the
// reference are held as doubles that are exactly representable in T (rT after every T operation).
KS_BIG inline void tr_solve(dd4ml_state* st, int mode, bool f32, Scratch* ws) {
    const double delta = st->delta[0], tol = st->tol[0], gg = st->gg[0];
    const int k = (st->second_order[0] != 0.0) ? (int)st->k[0] : 0;
    const double c_s0 = KS_CLOCK();
    const double gn = rT((st->norm_inf[0] != 0.0) ? st->ginf[0] : dsqrt(gg), f32);   // gnT
    const double gam = st->gamma[0];
    KS_SYNC();                          // all lanes hold the inputs before the master resets the report
    KS_PAR(i, MS) { st->cS[i] = 0.0; st->cY[i] = 0.0; }
    KS_MASTER {
        st->cyc_eigh[0] = st->cyc_newton[0] = st->jacobi_sweeps[0] = 0.0;
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
        st->gn[0] = gn;
    }
    if (gn <= tol) {  // tr.py:147 / lssr1_tr.py:509
        KS_MASTER {
            st->converged[0] = 1.0;
            st->pred[0] = st->pnorm[0] = st->gp[0] = st->psq[0] = st->dphi0[0] = 0.0;
            st->case_id[0] = DD4ML_CASE_FIRST_ORDER;
        }
        KS_SYNC();
        return;
    }
#define KS_FAIL(code) { KS_MASTER { st->err[0] = (code); st->solve_valid[0] = 0.0; } KS_SYNC(); return; }

    double pred, pp, gp, cg, lam = 0.0;   // the step: p = cg g + lam Psi wv
    if (k == 0) {  // optimizer_utils.py:197-215
        const double delT = rT(delta, f32);
        const double scale = rT(rT(ddiv(1.0, gn), f32) * delT, f32);  // python: delta / gn  ==  gn.reciprocal() * delta
        cg = -scale;
        pred = rT(delT * gn, f32);
        pp = cg * cg * gg;
        gp = cg * gg;
        KS_MASTER st->case_id[0] = DD4ML_CASE_FIRST_ORDER;
    } else {
        build_compact(st, ws);
        double* b = ws->b;
        KS_PAR(i, k) { const int pi = (int)st->order[i]; b[i] = st->Yg[pi] - gam * st->Sg[pi]; }  // Psi^T g
        KS_SYNC();
        bool ok = finite(gg) && finite(delta) && finite(gam);
        KS_LOOP
        for (int i = 0; i < k; ++i) {
            ok = ok && finite(b[i]);
            KS_LOOP
            for (int j = 0; j < k; ++j) ok = ok && finite(ws->G[i * M + j]) && finite(ws->Minv[i * M + j]);
        }
        if (!ok) KS_FAIL(DD4ML_ERR_NONFINITE)
        if (delta < 0) KS_FAIL(DD4ML_ERR_NEG_DELTA)

        // ---- OBS (obs.py:61-95) ----
        double *R = ws->R, *LU = ws->LU, *MR = ws->MR, *RMR = ws->RMR, *U = ws->U, *D = ws->D, *RinvU = ws->RinvU;
        if (!chol_upper(ws->G, k, R)) KS_FAIL(DD4ML_ERR_CHOLESKY)
        KS_PAR2(i, j, k) {
            LU[i * M + j] = ws->Minv[i * M + j];
            MR[i * M + j] = R[j * M + i];          // right-hand sides R^T
        } KS_END2
        KS_SYNC();
        if (!lu_factor(LU, k, ws->piv)) KS_FAIL(DD4ML_ERR_SINGULAR)
        lu_solve_cols(LU, ws->piv, k, MR, M, k);    // MR = Minv \ R^T
        KS_PAR2(i, j, k) { RMR[i * M + j] = dots(R + i * M, 1, MR + j, M, k); } KS_END2
        KS_SYNC();
        KS_PAR2(i, j, k) {
            if (i < j) {
                const double s = 0.5 * (RMR[i * M + j] + RMR[j * M + i]);
                RMR[i * M + j] = RMR[j * M + i] = s;
            }
        } KS_END2
        KS_SYNC();
        const double c_e0 = KS_CLOCK();
        const int sweeps = jacobi_eigh(RMR, k, D, U, ws);
        KS_MASTER { st->jacobi_sweeps[0] = sweeps; st->cyc_eigh[0] = KS_CLOCK() - c_e0; }
        double* Lam = ws->Ld;               // Lambda, a as T values held in doubles (+ float copies for T = float)
        double* a = ws->Ad;
        const double tolT = rT(OBS_TOL, f32), gamT = rT(gam, f32), delT = rT(delta, f32);
        KS_PAR(c, k) {  // RinvU[:, c] = R \ U[:, c];  a_c = (R^{-1} U)[:, c] . Psi^T g
            KS_LOOP
            for (int r = k - 1; r >= 0; --r) {
                double s = U[r * M + c];
                KS_LOOP
                for (int j = r + 1; j < k; ++j) s -= R[r * M + j] * RinvU[j * M + c];
                RinvU[r * M + c] = ddiv(s, R[r * M + r]);
            }
            const double s = dots(RinvU + c, M, b, 1, k);
            ws->ad[c] = s;
            a[c] = rT(s, f32);
            double L = rT(rT(D[c], f32) + gamT, f32);
            if (fabs(L) < tolT) L = 0.0;
            Lam[c] = L;
            ws->Af[c] = (float)a[c];
            ws->Lf[c] = (float)L;
        }
        KS_SYNC();
        const double gpgp = dotk(ws->ad, ws->ad, k);
        const double dd = gg - gpgp;
        const double a_last = rT(dsqrt(dd > 0.0 ? dd : 0.0), f32);
        const double lam_last = (fabs(gamT) < tolT) ? 0.0 : gamT;
        KS_MASTER { a[k] = a_last; Lam[k] = lam_last; ws->Af[k] = (float)a_last; ws->Lf[k] = (float)lam_last; }
        KS_SYNC();
        const double lam_min = (Lam[0] < gamT) ? Lam[0] : gamT;
        KS_MASTER st->lam_min[0] = lam_min;
        double tauT;
        double h2 = 0.0;
        KS_LOOP
        for (int i = 0; i <= k; ++i) { const double h = rT(ddiv(a[i], Lam[i]), f32); h2 = rT(h2 + rT(h * h, f32), f32); }
        if (lam_min > 0.0 && rT(dsqrt(h2), f32) <= delT) {  // case A, obs.py:97
            tauT = gamT;
            KS_MASTER st->case_id[0] = DD4ML_CASE_A;
        } else {
            if (lam_min <= 0.0 && phibar_f_rt(f32, -lam_min, ws, k + 1, delT, tolT) >= 0.0)
                KS_FAIL(DD4ML_ERR_HARD_CASE)   // obs.py:100-158, unreachable for gamma > 0
            double x0 = 0.0;
            if (!(lam_min > 0.0)) {  // obs.py:162-168
                double sh = rT(rT(ddiv(a[0], delT), f32) - Lam[0], f32);
                KS_LOOP
                for (int i = 1; i <= k; ++i) { const double v = rT(rT(ddiv(a[i], delT), f32) - Lam[i], f32); if (v > sh) sh = v; }
                x0 = (sh > -lam_min) ? sh : -lam_min;
            }
            int it, nd;
            const double c_n0 = KS_CLOCK();
            double sig = newton_rt(f32, x0, ws, k + 1, delT, tolT, it, nd);
            if (!finite(sig)) sig = newton_rt(f32, 0.0, ws, k + 1, delT, tolT, it, nd);
            KS_MASTER {
                st->cyc_newton[0] = KS_CLOCK() - c_n0;
                st->sigma0[0] = x0;
                st->sigma[0] = sig;
                st->newton_iters[0] = it;
                st->newton_deg[0] = nd;
                st->case_id[0] = DD4ML_CASE_C;
            }
            tauT = rT(gamT + sig, f32);
        }
        const double tau = tauT;
        KS_MASTER st->tau[0] = tau;
        // SMW step exactly as coded (obs.py:175-182): p_b = -g/tau + Psi (tau Minv + G)^{-1} Psi^T g
        double *W = ws->W, *wv = ws->wv, *x = ws->x, *z = ws->z, *Gw = ws->Gw;
        KS_PAR2(i, j, k) { W[i * M + j] = tau * ws->Minv[i * M + j] + ws->G[i * M + j]; } KS_END2
        KS_PAR(i, k) { wv[i] = b[i]; x[i] = b[i]; }
        KS_SYNC();
        if (!lu_factor(W, k, ws->pw)) KS_FAIL(DD4ML_ERR_SINGULAR)
        KS_PAR(c, 2) {
            if (c == 0) lu_solve_vec(W, ws->pw, k, wv, 1);     // wv = (tau Minv + G) \ Psi^T g
            else lu_solve_vec(LU, ws->piv, k, x, 1);           // x  = Minv \ Psi^T g
        }
        KS_SYNC();
        matvec(ws->G, wv, k, Gw);
        KS_PAR(i, k) z[i] = Gw[i];
        KS_SYNC();
        lu_solve_cols(LU, ws->piv, k, z, 1, 1);     // z  = Minv \ (G wv)
        const double bx = dotk(b, x, k), bw = dotk(b, wv, k), wGw = dotk(wv, Gw, k);
        const double bz = dotk(b, z, k), gwx = dotk(Gw, x, k), gwz = dotk(Gw, z, k);
        const double cgb = rT(ddiv(-1.0, tauT), f32);
#define KS_PP(c_, l_) ((c_) * (c_) * gg + 2.0 * (c_) * (l_) * bw + (l_) * (l_) * wGw)
#define KS_GP(c_, l_) ((c_) * gg + (l_) * bw)
        // ---- Cauchy point and dogleg (optimizer_utils.py:259-296) ----
        const double gBg = gam * gg + bx;
        KS_MASTER st->gBg[0] = gBg;
        cg = cgb;
        lam = 1.0;
        double branch = 0.0;
        if (st->dogleg[0] != 0.0 && gBg > 0.0) {
            const double alpha = ddiv(gn * gn, gBg);
            const double cc = (alpha * dsqrt(gg) >= delta) ? -ddiv(delta, gn) : -alpha;
            const double pb_norm = dsqrt(KS_PP(cgb, 1.0));
            const double pc_norm = fabs(cc) * dsqrt(gg);
            if (pb_norm <= delta) {
                branch = 1.0;
            } else if (pc_norm >= delta) {
                cg = cc; lam = 0.0; branch = 2.0;
            } else {
                const double dcg = cgb - cc;                   // d = p_b - p_c = dcg g + Psi wv
                const double qa = KS_PP(dcg, 1.0);
                const double qb = 2.0 * cc * KS_GP(dcg, 1.0);
                const double qc = cc * cc * gg - delta * delta;
                const double disc = qb * qb - 4.0 * qa * qc;
                if (disc < 0.0) {
                    cg = cc; lam = 0.0; branch = 3.0;
                } else {
                    const double t = ddiv(-qb + dsqrt(disc), 2.0 * qa);
                    cg = cc + t * dcg; lam = t; branch = 4.0;
                }
            }
        }
        KS_MASTER st->dogleg_branch[0] = branch;
        // ---- predicted reduction -(g^T p + 1/2 p^T B p) (optimizer_utils.py:298-305) ----
        // Psi^T p = q = cg b + lam G wv,  B p = gamma p + Psi Minv^{-1} q,  Minv^{-1} q = cg x + lam z
        pp = KS_PP(cg, lam);
        gp = KS_GP(cg, lam);
        const double qy = cg * cg * bx + cg * lam * (bz + gwx) + lam * lam * gwz;
        pred = -(gp + 0.5 * (gam * pp + qy));
#undef KS_PP
#undef KS_GP
    }

    if (mode == 1) {  // lssr1_tr.py:563-587 (the momentum term is identically zero)
        if (pp > tol) {
            const double nrm = dsqrt(rT(pp, f32));
            const double q = ddiv(delta, nrm);
            const double sc = (q < 1.0) ? q : 1.0;
            cg *= sc;
            lam *= sc;
            pp *= sc * sc;
            gp *= sc;
            KS_MASTER st->clip_scale[0] = sc;
        }
        if (gp > 0.0) {
            cg = -cg;
            lam = -lam;
            pred = -pred;
            gp = -gp;
            KS_MASTER st->flip[0] = 1.0;
        }
    }
    KS_PAR(i, k) {
        const int pi = (int)st->order[i];
        const double c = lam * ws->wv[i];
        st->cY[pi] = c;
        st->cS[pi] = -gam * c;
    }
    KS_MASTER {
        st->cg[0] = cg;
        st->pred[0] = pred;
        st->psq[0] = pp;
        st->pnorm[0] = dsqrt(pp > 0.0 ? pp : 0.0);
        st->gp[0] = gp;
        st->dphi0[0] = gp;
        st->cyc_solve[0] = KS_CLOCK() - c_s0;
    }
    KS_SYNC();
#undef KS_FAIL
}
// typed wrapper (tests/hostshim, readability)
template <class T>
KS_FN inline void tr_solve(dd4ml_state* st, int mode, Scratch* ws) { tr_solve(st, mode, sizeof(T) == 4, ws); }

// =================================================================================
// LSR1.update_memory acceptance chain (lsr1.py:66-144) on the candidate pair whose
// reductions sit in c_ss, c_sy, c_yy, c_Ss, c_Ys, c_Sy, c_Yy (PHYSICAL slot index)
// and whose vectors were written to the spare slot.  On acceptance the spare slot
// joins the ring (dropping the oldest pair at capacity), the Gram blocks grow by one
// row/column and gamma <- max(y^T y / y^T s, tol).  Collective; the result is uniform.
// =================================================================================
KS_BIG inline bool lsr1_try_update(dd4ml_state* st, Scratch* ws) {
    const double tol = st->tol[0];
    const double ss = st->c_ss[0], sy = st->c_sy[0], yy = st->c_yy[0];
    const int k = (int)st->k[0], m = (int)st->mem_len[0];
    const double gam = st->gamma[0];
    const int ns = (int)st->spare[0];
    const int oldest = (int)st->order[0];
    KS_SYNC();
    KS_MASTER st->pair_ok[0] = 0.0;
    const double sn = dsqrt(ss), yn = dsqrt(yy);
    if (!(sn > tol) || !(yn > tol)) return false;        // :72
    if (!(fabs(sy) > tol * (sn * yn))) return false;     // :78
    double us, uu;
    if (k == 0) {  // B = gamma I
        us = sy - gam * ss;
        uu = yy - 2.0 * gam * sy + gam * gam * ss;
    } else {
        build_compact(st, ws);
        double *bs = ws->bs, *by = ws->by, *x = ws->x, *LU = ws->LU, *Gx = ws->Gw;
        KS_PAR(i, k) {
            const int pi = (int)st->order[i];
            bs[i] = st->c_Ys[pi] - gam * st->c_Ss[pi];   // Psi^T s
            by[i] = st->c_Yy[pi] - gam * st->c_Sy[pi];   // Psi^T y
            x[i] = bs[i];
        }
        KS_PAR2(i, j, k) { LU[i * M + j] = ws->Minv[i * M + j]; } KS_END2
        KS_SYNC();
        if (!lu_factor(LU, k, ws->piv)) { KS_MASTER st->err[0] = DD4ML_ERR_SINGULAR; KS_SYNC(); return false; }
        lu_solve_cols(LU, ws->piv, k, x, 1, 1);          // x = Minv \ Psi^T s
        matvec(ws->G, x, k, Gx);
        const double bsx = dotk(bs, x, k);
        us = sy - gam * ss - bsx;
        uu = yy + gam * gam * ss + dotk(x, Gx, k) - 2.0 * gam * sy - 2.0 * dotk(by, x, k) + 2.0 * gam * bsx;
    }
    const double un = dsqrt(uu > 0.0 ? uu : 0.0);
    if (!(fabs(us) > tol * sn * un)) return false;       // :86
    double gc = ddiv(yy, sy);
    if (gc < tol) gc = tol;                              // :90
    const double pc2 = yy - 2.0 * gc * sy + gc * gc * ss;
    if (!(dsqrt(pc2 > 0.0 ? pc2 : 0.0) > tol * yn)) return false;  // :93
    // ---- accept ----
    int kk = k, newspare;
    if (k >= m) {  // drop the oldest pair; its slot becomes the spare
        newspare = oldest;
    } else {
        kk = k + 1;
        unsigned used = 1u << ns;       // first unused physical slot
        KS_LOOP
        for (int i = 0; i < k; ++i) used |= 1u << (int)st->order[i];
        newspare = 0;
        KS_LOOP
        while (newspare < MS && ((used >> newspare) & 1u)) ++newspare;
    }
    KS_SYNC();                          // every lane has read the old ring before it changes
    KS_PAR(i, k) {                      // Gram rows/columns of the new pair against the live pairs
        const int pj = (int)st->order[i];
        if (!(k >= m && i == 0)) {      // (the dropped pair's entries are dead)
            st->SS[ns * MS + pj] = st->SS[pj * MS + ns] = st->c_Ss[pj];
            st->YY[ns * MS + pj] = st->YY[pj * MS + ns] = st->c_Yy[pj];
            st->SY[pj * MS + ns] = st->c_Sy[pj];  // S_j^T y
            st->SY[ns * MS + pj] = st->c_Ys[pj];  // s^T Y_j
        }
    }
    KS_SYNC();
    KS_MASTER {
        if (k >= m) {
            KS_LOOP
            for (int i = 0; i + 1 < k; ++i) st->order[i] = st->order[i + 1];
            st->order[k - 1] = ns;
        } else {
            st->order[k] = ns;
        }
        st->SS[ns * MS + ns] = ss;
        st->SY[ns * MS + ns] = sy;
        st->YY[ns * MS + ns] = yy;
        st->k[0] = kk;
        st->spare[0] = newspare;
        st->gamma[0] = gc;
        st->gen[0] += 1.0;
        st->pair_ok[0] = 1.0;
    }
    KS_SYNC();
    return true;
}

// LSR1.B (lsr1.py:181-198): out = gamma v + Psi (Minv \ Psi^T v), expressed as the
// coefficients (cg, cS, cY) of the output pass.  S^T v, Y^T v are read from Sg, Yg.  Collective.
// With tau > 0 (OBS.ComputeSBySMW, obs.py:175-182, tau read from st->tau): p = -g/tau +
// Psi ((tau Minv + Psi^T Psi) \ Psi^T g), as coded in the reference (no 1/tau on the Psi term).
KS_BIG inline void coefs_B_or_smw(dd4ml_state* st, Scratch* ws, bool smw) {
    const int k = (int)st->k[0];
    const double gam = st->gamma[0], tau = st->tau[0];
    KS_SYNC();
    KS_PAR(i, MS) { st->cS[i] = 0.0; st->cY[i] = 0.0; }
    KS_MASTER {
        st->cg[0] = smw ? -1.0 / tau : gam;
        st->converged[0] = 0.0;
        st->solve_valid[0] = 0.0;
        st->err[0] = DD4ML_ERR_NONE;
    }
    KS_SYNC();
    if (k == 0) return;
    build_compact(st, ws);
    double *x = ws->x, *A = ws->W;
    KS_PAR(i, k) { const int pi = (int)st->order[i]; x[i] = st->Yg[pi] - gam * st->Sg[pi]; }
    KS_PAR2(i, j, k) {
        A[i * M + j] = smw ? tau * ws->Minv[i * M + j] + ws->G[i * M + j] : ws->Minv[i * M + j];
    } KS_END2
    KS_SYNC();
    if (!lu_factor(A, k, ws->pw)) { KS_MASTER st->err[0] = DD4ML_ERR_SINGULAR; KS_SYNC(); return; }
    lu_solve_cols(A, ws->pw, k, x, 1, 1);
    KS_PAR(i, k) {
        const int pi = (int)st->order[i];
        st->cY[pi] = x[i];
        st->cS[pi] = -gam * x[i];
    }
    KS_SYNC();
}
KS_FN inline void lsr1_B_coefs(dd4ml_state* st, Scratch* ws) { coefs_B_or_smw(st, ws, false); }
KS_FN inline void smw_coefs(dd4ml_state* st, Scratch* ws) { coefs_B_or_smw(st, ws, true); }

// rho and the accept decision of TR.step (tr.py:186-191), in the loss dtype T.  Pure.
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
// out_gg / out_ginf (norms of the trial gradient) have been filled by the kernel.  Collective.
KS_BIG inline void tr_finish(dd4ml_state* st, bool accept, double rho, bool f32, Scratch* ws) {
    const double delta = st->delta[0], tol = st->tol[0];
    const bool so = st->second_order[0] != 0.0;
    const double pnorm = st->pnorm[0];
    const bool try_pair = accept && so && dsqrt(st->c_ss[0]) > tol && dsqrt(st->c_yy[0]) > tol;  // tr.py:202
    KS_SYNC();
    KS_MASTER {
        st->rho[0] = rho;
        st->accept[0] = accept ? 1.0 : 0.0;
        st->pair_tried[0] = try_pair ? 1.0 : 0.0;
        st->pair_ok[0] = 0.0;
        st->steps[0] += 1.0;
    }
    KS_SYNC();
    if (try_pair) lsr1_try_update(st, ws);
    KS_MASTER {
        if (accept) {
            if (rT(rho, f32) > rT(st->nu_inc[0], f32) && rT(pnorm, f32) >= rT(0.9 * delta, f32)) {  // tr.py:207
                const double d = st->inc_factor[0] * delta;
                st->delta[0] = (st->max_delta[0] < d) ? st->max_delta[0] : d;
            }
            st->out_gn[0] = rT((st->norm_inf[0] != 0.0) ? st->out_ginf[0] : dsqrt(st->out_gg[0]), f32);
        } else {
            const double d = st->dec_factor[0] * delta;
            st->delta[0] = (st->min_delta[0] > d) ? st->min_delta[0] : d;
            st->out_gn[0] = st->gn[0];
            st->out_gg[0] = st->gg[0];
            st->out_ginf[0] = st->ginf[0];
        }
    }
    KS_SYNC();
}
template <class T>
KS_FN inline void tr_finish(dd4ml_state* st, bool accept, double rho, Scratch* ws) {
    tr_finish(st, accept, rho, sizeof(T) == 4, ws);
}

// APTS_Base.control_step decision (apts_base.py:483-501): reject when pred < tol or
// rho < nu_dec; enlarge when rho > nu_inc.  delta is the APTS-level radius.  Single thread.
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
