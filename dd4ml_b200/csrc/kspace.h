// k-space routines of the trust-region step: everything that happens between the streaming
// reductions over the n-vectors and the streaming output pass.
//
// All inputs are O(k^2) numbers (Gram blocks of the L-SR1 pairs, Psi^T g, g^T g) held in the device
// state block.  The routines run in the ONE-CTA k-space kernel (k_kspace.cu: one warp for memories of
// up to 5 pairs, four warps above) that follows every reduction kernel ("small m x m inner solve kept in
// one CTA").  The kernel is pure latency: it runs once per launch, right after the closure's kernels
// have evicted the instruction cache, on O(k^3) flops.  What sets its duration is therefore
//   (1) the number of DISTINCT instructions (cold fetch: ~7 cycles each) -- one shared copy of every
//       building block (Gauss-Jordan, small dot product, division, square root), no unrolled variants;
//   (2) the length of the dependent chain -- every O(k^2)/O(k^3) step is a *parallel region* over
//       matrix elements, and the algorithms are chosen for few, wide regions:
//         * Cholesky of the AUGMENTED block [Psi^T Psi | Psi^T g]: the extra column comes out as
//           R^-T Psi^T g, so that a = U^T (R^-T Psi^T g) needs no triangular solve per eigenvector;
//         * M^-1 by Gauss-Jordan with partial pivoting between two buffers (ONE region per pivot: no
//           row-swap / multiplier / update phases), after which every solve is a matrix-vector product;
//         * two-sided Jacobi with the round-robin ordering, rotations of a round applied to A and V in
//           ONE region between two buffers (A' = J^T A J element-wise from 4 old entries).
//
// Execution model (also what lets tests/hostshim compile this file for the CPU):
//   KS_PAR(i, n) { body }  KS_SYNC();      a parallel region: iterations are independent -- an iteration
//                                          never reads what another iteration of the same region writes
//   everything else                        uniform code, executed redundantly by all threads; shared state
//                                          is written by the master thread only (KS_MASTER), is never read
//                                          back before the next KS_SYNC(), and nothing that uniform code
//                                          has read is overwritten before the next KS_SYNC()
// On the device KS_PAR strides the threads of the CTA and KS_SYNC is a CTA barrier; on the host KS_PAR is
// a plain loop and KS_SYNC a no-op, which is the same program by the independence rule.
//
// Reference behaviour restated here (cruzas/DD4ML):
//   LSR1.precompute / B / update_memory      optimizers/lsr1.py:53-198
//   OBS.solve_tr_subproblem / Newton / SMW   solvers/obs.py:46-254
//   solve_tr_first_order / second_order      utility/optimizer_utils.py:197-307
//   TR.step ratio test and radius update     optimizers/tr.py:186-221
//   LSSR1_TR clip-to-radius / descent flip   optimizers/lssr1_tr.py:563-587
// Linear algebra (Cholesky, Gauss-Jordan, symmetric eigen-decomposition) runs in double; the quantities the
// reference thresholds or iterates on in its working dtype (Lambda, a_j, the Newton iteration, rho) are
// rounded to that dtype T first.
#pragma once
#include <math.h>

#include "state.h"

#ifdef __CUDACC__
#define KS_FN __host__ __device__
#define KS_BIG __host__ __device__ __noinline__    // one copy of every shared routine: code size is latency here
#define KS_LOOP _Pragma("unroll 1")
#else
#define KS_FN
#define KS_BIG
#define KS_LOOP
#endif
#if defined(__CUDA_ARCH__)
#define KS_CLOCK() ((double)clock64())
#define KS_TID ((int)threadIdx.x)
#define KS_NT ((int)blockDim.x)
#define KS_SYNC() __syncthreads()
#else
#define KS_CLOCK() 0.0
#define KS_TID 0
#define KS_NT 1
#define KS_SYNC() ((void)0)
#endif
#define KS_PAR(i, n) KS_LOOP for (int i = KS_TID; i < (n); i += KS_NT)
#define KS_MASTER if (KS_TID == 0)
// 2-D region over rows x cols elements, every slot live: KS_PAR2(i, j, rows, cols) { ... } KS_END2
#define KS_PAR2(i, j, rows, cols) { const ks::Div dv_(cols); const int ne_ = (rows) * (cols);                 \
    KS_LOOP for (int e_ = KS_TID; e_ < ne_; e_ += KS_NT) { const int i = dv_.quot(e_), j = e_ - i * (cols);
#define KS_END2 } }

namespace ks {

constexpr int M = DD4ML_MAXM;
constexpr int MS = DD4ML_MAXS;
constexpr int LD = M + 1;         // leading dimension of the k x k blocks (odd; column k can hold one more vector)
constexpr int LW = 2 * M + 1;     // leading dimension of the Gauss-Jordan blocks [A | B]
constexpr double OBS_TOL = 1e-6;  // obs.py:29

// e / c for 0 <= e < 4096, 1 <= c <= 32 without an integer division: (e + 1/2) / c is at least 1/(2c)
// away from an integer, far more than the float32 rounding error of the product
struct Div {
#if defined(__CUDA_ARCH__)
    float inv;
    __device__ explicit Div(int c) : inv(__fdividef(1.0f, (float)c)) {}
    __device__ int quot(int e) const { return __float2int_rz(((float)e + 0.5f) * inv); }
#else
    int c;
    explicit Div(int c_) : c(c_) {}
    int quot(int e) const { return e / c; }
#endif
};

KS_FN inline bool finite(double x) { return x == x && fabs(x) <= 1.79769313486231570e308; }
// IEEE double division / square root expand to ~30 instructions each: one out-of-line copy
KS_BIG inline double ddiv(double a, double b) { return a / b; }
KS_BIG inline double dsqrt(double x) { return sqrt(x); }
KS_BIG inline double drsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
// 2^-floor(log2 |x|) for a normal x (1 for 0 / subnormal / non-finite x): exact exponent normalisation
KS_FN inline double pow2_inv(double x) {
#if defined(__CUDA_ARCH__)
    const int e = (__double2hiint(x) >> 20) & 0x7ff;
    return (e == 0 || e >= 0x7fe) ? 1.0 : __hiloint2double((2046 - e) << 20, 0);
#else
    if (!(fabs(x) >= 2.2250738585072014e-308) || !(fabs(x) <= 8.9e307)) return 1.0;
    int ex;
    (void)frexp(x, &ex);
    return ldexp(1.0, 1 - ex);
#endif
}
// |tan| of the Jacobi rotation from the exponent-normalised (aqq - app, 2 apq): any value within ~1e-6
// relative of |a| / (|d| + sqrt(d^2 + a^2)) will do (see jacobi_core), so the device uses the approximate
// float32 reciprocal / square root units
KS_FN inline float rot_tan(float df, float af) {
#if defined(__CUDA_ARCH__)
    return __fdividef(fabsf(af), fabsf(df) + __fsqrt_rn(fmaf(df, df, af * af)));
#else
    return fabsf(af) / (fabsf(df) + sqrtf(df * df + af * af));
#endif
}
// 1 / sqrt(1 + t^2) to double precision (|t| <= 1): float32 seed + two Newton steps, in line
KS_FN inline double rot_cos(double t) {
    const double h = t * t + 1.0;
#if defined(__CUDA_ARCH__)
    double c = (double)rsqrtf((float)h);
    c = c * (1.5 - 0.5 * h * c * c);
    c = c * (1.5 - 0.5 * h * c * c);
    return c;
#else
    return 1.0 / sqrt(h);
#endif
}
// Round to the reference's working dtype (float32 or float64).  A float32 operation on float32
// operands equals the float64 operation rounded to float32 (53 >= 2*24 + 2 bits: no double-rounding
// error for + - * / sqrt), so `rT(a op b, f32)` IS the reference's float32 arithmetic.
KS_FN inline double rT(double x, bool f32) { return f32 ? (double)(float)x : x; }
KS_FN inline float ks_sqrt(float x) { return sqrtf(x); }
KS_FN inline double ks_sqrt(double x) { return sqrt(x); }

// Work arrays of the k-space routines: shared memory of the k-space kernel.  k x k blocks are row-major
// with leading dimension LD, the Gauss-Jordan blocks with LW.
struct Scratch {
    double Minv[M * LD];  // diag(S^T Y) + L + L^T - gamma S^T S + tol*||.||_F I   (lsr1.py:168-179)
    double G[M * LD];     // Psi^T Psi
    double Mi[M * LD];    // Minv^{-1}
    double R[M * LD];     // Cholesky factor of G (upper); column k: R^-T Psi^T g
    double T[M * LD];     // Minv^{-1} R^T, then R Minv^{-1} R^T
    double J0[M * LD], J1[M * LD], V0[M * LD], V1[M * LD];   // Jacobi: matrix and eigenvectors, two buffers each
    double X0[M * LW], X1[M * LW];                           // Gauss-Jordan: two buffers
    double vec[6 * M];    // b = Psi^T g, x = Minv\b, wv, G wv, z = Minv\(G wv), spare
    double Rd[M], rowa[M], rowb[M], ad[M + 1], dot6[6];
    double pacc[2 * M];   // per rotation slot: off-diagonal mass annihilated in this sweep (two sweeps alternate)
    double rc[M], rt[M];  // Jacobi rotations of one round, per INDEX: new col j = rc[j] col j + rt[j] col mate[j]
    int mate[M];
    double Ld[M + 1], Ad[M + 1];   // Lambda, a in the working dtype (T = double)
    float Lf[M + 1], Af[M + 1];    //                                 (T = float)
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

// sum_{i<k} a[i*sa] * b[i*sb] (per thread).  ONE out-of-line copy.  All operands are fetched before the
// first multiply (the shared-memory latency is paid once, not once per term) in fixed-size batches of 4:
// terms beyond k are predicated off.
KS_BIG inline double dots(const double* a, int sa, const double* b, int sb, int k) {
    double s = 0.0;
#if defined(__CUDA_ARCH__)
    KS_LOOP
    for (int i0 = 0; i0 < k; i0 += 4, a += 4 * sa, b += 4 * sb) {
        double x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool on = i0 + u < k;
            x[u] = on ? a[u * sa] : 0.0;
            y[u] = on ? b[u * sb] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) s += x[u] * y[u];
    }
#else
    for (int i = 0; i < k; ++i) s += a[i * sa] * b[i * sb];
#endif
    return s;
}
KS_FN inline double dotk(const double* a, const double* b, int k) { return dots(a, 1, b, 1, k); }

// C[i*ldc + j] = sum_{l<nl} A[i*sai + l*sal] * B[l*sbl + j*sbj], i < ni, j < nj.  Collective.
KS_BIG inline void mm(double* C, int ldc, const double* A, int sai, int sal, const double* B, int sbl, int sbj,
                      int ni, int nj, int nl) {
    KS_PAR2(i, j, ni, nj) { C[i * ldc + j] = dots(A + i * sai, sal, B + j * sbj, sbl, nl); } KS_END2
    KS_SYNC();
}

// Gauss-Jordan elimination with partial pivoting (the pivot rule of torch.linalg.solve's getrf) of the
// k x w block [A | B] in ws->X0 (w > k, leading dimension LW).  ONE region per pivot: element (i, j) of
// the next block is formed from the current one -- row swap, normalisation of the pivot row and
// elimination of column c from every other row at once -- in the other buffer.  Columns <= c are dead
// (unit vectors) and skipped.  Returns the buffer whose columns k..w-1 hold A^{-1} B, nullptr when a pivot
// is exactly zero or non-finite.  Collective.
KS_BIG inline double* gauss_jordan(Scratch* ws, int k, int w) {
    double* src = ws->X0;
    double* dst = ws->X1;
    KS_LOOP
    for (int c = 0; c < k; ++c) {
        int p = c;
        double best = fabs(src[c * LW + c]);
        KS_LOOP
        for (int r = c + 1; r < k; ++r) {
            const double v = fabs(src[r * LW + c]);
            if (v > best) { best = v; p = r; }
        }
        if (best == 0.0 || !finite(best)) return nullptr;
        const double inv = ddiv(1.0, src[p * LW + c]);
        const int nc = w - c - 1;
        KS_PAR2(i, jj, k, nc) {
            const int j = c + 1 + jj;
            const int ri = (i == c) ? p : ((i == p) ? c : i);      // row i after the swap c <-> p
            const double pj = src[p * LW + j] * inv;              // normalised pivot row
            dst[i * LW + j] = (i == c) ? pj : src[ri * LW + j] - src[ri * LW + c] * pj;
        } KS_END2
        KS_SYNC();
        double* t = src; src = dst; dst = t;
    }
    return src;
}

// R upper triangular with R^T R = G (torch.linalg.cholesky(upper=True)), right-looking, on the block
// [G | b] (b in column k): one region scales the pivot row, one updates the trailing block, so that
// on exit R[:, k] = R^-T b.  Strictly lower part zero.  Collective.
KS_BIG inline bool chol_aug(Scratch* ws, int k) {
    double* R = ws->R;
    const double* b = ws->vec;
    KS_PAR2(i, j, k, k + 1) { R[i * LD + j] = (j == k) ? b[i] : ((j >= i) ? ws->G[i * LD + j] : 0.0); } KS_END2
    KS_SYNC();
    KS_LOOP
    for (int j = 0; j < k; ++j) {
        const double s = R[j * LD + j];
        if (!(s > 0.0) || !finite(s)) return false;
        const double inv = drsqrt(s);
        const int w = k - j;               // columns j+1 .. k
        KS_PAR(c, w + 1) {
            if (c == w) ws->Rd[j] = s * inv;
            else R[j * LD + j + 1 + c] *= inv;
        }
        KS_SYNC();
        KS_PAR2(a, c, k - 1 - j, w) {      // rows j+1 .. k-1
            const int i = j + 1 + a, cc = j + 1 + c;
            if (cc >= i) R[i * LD + cc] -= R[j * LD + i] * R[j * LD + cc];
        } KS_END2
        KS_SYNC();
    }
    KS_PAR(j, k) R[j * LD + j] = ws->Rd[j];
    KS_SYNC();
    return true;
}

// Jacobi eigen-decomposition of the symmetric k x k matrix in ws->J0 (destroyed) with the round-robin
// (tournament) ordering: the floor(k/2) rotations of a round touch disjoint index pairs and are applied
// together -- k-1 (k) rounds per sweep instead of k(k-1)/2 dependent rotations.  A round is two regions:
// the rotation angles (per index j: partner mate[j] and the coefficients of new col j = rc[j] col j +
// rt[j] col mate[j]), then A' = J^T A J and V' = V J element by element into the other buffers.
// The sweeps stop when the off-diagonal mass annihilated by a whole sweep is below 1e-17 of ||A||_F^2
// (Jacobi converges quadratically: what is left after that sweep is far smaller; measured on random and
// clustered spectra, k <= 10: eigen-pairs to <= 1.4e-13 ||A||, 3.6 sweeps at k = 3, 5.7 at k = 10).
// On exit *Aout / *Vout point at the buffers with the diagonalised matrix / the eigenvectors (columns,
// unsorted, unnormalised signs).  Returns the number of sweeps.  Collective.
// The rotation angle is computed in float32 on exponent-normalised inputs (any angle keeps the
// transformation orthogonal as long as c = 1/sqrt(1 + t^2) and s = t c are formed in double; a 1e-7
// relative error of t only leaves 1e-7 of the pivot behind for the next sweep), which takes the
// double-precision division and square root off the dependent chain of every round.
KS_BIG inline int jacobi_core(Scratch* ws, int k, double** Aout, double** Vout) {
    double *A = ws->J0, *A2 = ws->J1, *V = ws->V0, *V2 = ws->V1;
    KS_PAR2(i, j, k, k) { V[i * LD + j] = (i == j) ? 1.0 : 0.0; } KS_END2
    KS_PAR(i, k) ws->rowa[i] = dots(A + i * LD, 1, A + i * LD, 1, k);
    KS_SYNC();
    double fro = 0.0;
    KS_LOOP
    for (int i = 0; i < k; ++i) fro += ws->rowa[i];
    const double thr = 1e-17 * fro;
    const int n = k + (k & 1);          // players of the tournament (one dummy when k is odd)
    const int half = n >> 1;
    int sweeps = 0;
    if (k <= 3) {
        // k = 2, 3: ONE real rotation per round (k = 3: the fourth player is the dummy), so every thread forms it
        // itself from three matrix entries -- no rotation region, no per-index tables, one barrier per round.
        // Same pairs in the same order as the general schedule below: bit-identical results.
        KS_LOOP
        for (int sweep = 0; sweep < 40 && k > 1; ++sweep) {
            double tot = 0.0;
            KS_LOOP
            for (int round = 0; round < n - 1; ++round) {
                int p = 0, q = 1;
                if (k == 3) {
                    p = round + 1; if (p >= 3) p -= 3;
                    q = round + 2; if (q >= 3) q -= 3;
                    if (p > q) { const int u = p; p = q; q = u; }
                }
                double c = 1.0, s = 0.0;
                const double apq = A[p * LD + q];
                if (apq != 0.0) {
                    const double d = A[q * LD + q] - A[p * LD + p];
                    const double a2 = 2.0 * apq;
                    const double sc = pow2_inv(fmax(fabs(d), fabs(a2)));
                    const float tf = rot_tan((float)(d * sc), (float)(a2 * sc));
                    const double tt = ((d >= 0.0) == (a2 >= 0.0)) ? (double)tf : -(double)tf;
                    c = rot_cos(tt);
                    s = tt * c;
                    tot += apq * apq;
                }
                KS_PAR2(i, j, k, k) {
                    const int mi = (i == p) ? q : ((i == q) ? p : i), mj = (j == p) ? q : ((j == q) ? p : j);
                    const double ci = (mi != i) ? c : 1.0, ti = (i == p) ? -s : ((i == q) ? s : 0.0);
                    const double cj = (mj != j) ? c : 1.0, tj = (j == p) ? -s : ((j == q) ? s : 0.0);
                    const double u = cj * A[i * LD + j] + tj * A[i * LD + mj];
                    const double v = cj * A[mi * LD + j] + tj * A[mi * LD + mj];
                    A2[i * LD + j] = ci * u + ti * v;
                    V2[i * LD + j] = cj * V[i * LD + j] + tj * V[i * LD + mj];
                } KS_END2
                KS_SYNC();
                double* x = A; A = A2; A2 = x;
                x = V; V = V2; V2 = x;
            }
            ++sweeps;
            if (!(tot > thr)) break;
        }
        *Aout = A;
        *Vout = V;
        return sweeps;
    }
    KS_LOOP
    for (int sweep = 0; sweep < 40 && k > 1; ++sweep) {
        double* pacc = ws->pacc + (sweep & 1) * M;
        KS_PAR(t, half) pacc[t] = 0.0;  // (slot t is only touched by its own thread until the sweep's last barrier)
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
                int mp = p;
                if (q < k) {
                    mp = q;
                    const double apq = A[p * LD + q];
                    if (apq != 0.0) {
                        // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (aqq - app) / (2 apq)
                        const double d = A[q * LD + q] - A[p * LD + p];
                        const double a2 = 2.0 * apq;
                        const double sc = pow2_inv(fmax(fabs(d), fabs(a2)));
                        const float df = (float)(d * sc), af = (float)(a2 * sc);
                        const float tf = rot_tan(df, af);
                        const double tt = ((d >= 0.0) == (a2 >= 0.0)) ? (double)tf : -(double)tf;
                        c = rot_cos(tt);
                        s = tt * c;
                        pacc[t] += apq * apq;
                    }
                    ws->mate[q] = p; ws->rc[q] = c; ws->rt[q] = s;
                }
                ws->mate[p] = mp; ws->rc[p] = c; ws->rt[p] = -s;      // (q >= k: p sits this round out)
            }
            KS_SYNC();
            KS_PAR2(i, j, k, k) {
                const int mi = ws->mate[i], mj = ws->mate[j];
                const double ci = ws->rc[i], ti = ws->rt[i], cj = ws->rc[j], tj = ws->rt[j];
                const double u = cj * A[i * LD + j] + tj * A[i * LD + mj];       // (A J)[i][j]
                const double v = cj * A[mi * LD + j] + tj * A[mi * LD + mj];     // (A J)[mate i][j]
                A2[i * LD + j] = ci * u + ti * v;   // (the pivot keeps its ~1e-7 residue of the float32 angle: next sweep)
                V2[i * LD + j] = cj * V[i * LD + j] + tj * V[i * LD + mj];
            } KS_END2
            KS_SYNC();
            double* x = A; A = A2; A2 = x;
            x = V; V = V2; V2 = x;
        }
        ++sweeps;
        double tot = 0.0;
        KS_LOOP
        for (int t = 0; t < half; ++t) tot += pacc[t];
        if (!(tot > thr)) break;        // (also leaves on NaN)
    }
    *Aout = A;
    *Vout = V;
    return sweeps;
}

// Stand-alone form (tests): w ascending, U[:, j] the eigenvector of w[j], each with its largest-|.|
// component positive (lowest index on ties) -- the documented sign convention of this implementation
// (DESIGN.md section 2.1).  Ain / U: row-major, leading dimension LD.
KS_FN inline int jacobi_eigh(const double* Ain, int k, double* w, double* U, Scratch* ws) {
    KS_PAR2(i, j, k, k) { ws->J0[i * LD + j] = Ain[i * LD + j]; } KS_END2
    KS_SYNC();
    double *A, *V;
    const int sweeps = jacobi_core(ws, k, &A, &V);
    KS_PAR(j, k) {
        const double wj = A[j * LD + j];
        int r = 0, im = 0;
        double best = -1.0;
        KS_LOOP
        for (int i = 0; i < k; ++i) {
            const double wi = A[i * LD + i];
            r += (wi < wj || (wi == wj && i < j)) ? 1 : 0;
            if (fabs(V[i * LD + j]) > best) { best = fabs(V[i * LD + j]); im = i; }
        }
        const double sg = (V[im * LD + j] < 0.0) ? -1.0 : 1.0;
        w[r] = wj;
        KS_LOOP
        for (int i = 0; i < k; ++i) U[i * LD + r] = sg * V[i * LD + j];
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
    iters = 0;
    ndeg = 0;
    KS_LOOP
    for (;;) {  // obs.py:217: evaluate; while |f| > tol and iters < 200: step, evaluate
        sec.fg(x, f, g, deg);
        ndeg += deg ? 1 : 0;
        if (!(fabs(f) > tol) || iters >= 200) break;
        x = x - f / g;
        ++iters;
    }
    return x;
}

// ---- logical (oldest..newest) views of the Gram blocks: Minv and G = Psi^T Psi (collective) ------
// Returns false when a Gram entry is not finite.
KS_BIG inline bool build_compact(const dd4ml_state* st, Scratch* ws) {
    const int k = (int)st->k[0];
    const double gam = st->gamma[0];
    const bool mx = st->use_mx[0] != 0.0;
    KS_PAR2(i, j, k, k) {
        const int pi = (int)st->order[i], pj = (int)st->order[j];
        const double sy_ij = st->SY[pi * MS + pj], sy_ji = st->SY[pj * MS + pi];
        const double ss = st->SS[pi * MS + pj], yy = st->YY[pi * MS + pj];
        ws->Minv[i * LD + j] = mx ? st->Mx[i * M + j]       // explicit Minv (already regularised)
                                  : ((i >= j) ? sy_ij : sy_ji) - gam * ss;
        ws->G[i * LD + j] = yy - gam * (sy_ij + sy_ji) + gam * gam * ss;
    } KS_END2
    KS_SYNC();
    KS_PAR(i, 2 * k) {                  // row sums of squares of Minv (i < k) and of G
        const double* row = (i < k) ? ws->Minv + i * LD : ws->G + (i - k) * LD;
        const double s = dots(row, 1, row, 1, k);
        if (i < k) ws->rowa[i] = s; else ws->rowb[i - k] = s;
    }
    KS_SYNC();
    double fro = 0.0, frog = 0.0;
    KS_LOOP
    for (int i = 0; i < k; ++i) { fro += ws->rowa[i]; frog += ws->rowb[i]; }
    if (!mx) {
        const double reg = st->tol[0] * dsqrt(fro);
        KS_PAR(i, k) ws->Minv[i * LD + i] += reg;
        KS_SYNC();
    }
    return finite(fro) && finite(frog);
}

// Mi = Minv^{-1} (Gauss-Jordan on [Minv | I]).  Collective; false when singular.
KS_BIG inline bool invert_minv(Scratch* ws, int k) {
    KS_PAR2(i, j, k, 2 * k) { ws->X0[i * LW + j] = (j < k) ? ws->Minv[i * LD + j] : ((j - k == i) ? 1.0 : 0.0); } KS_END2
    KS_SYNC();
    const double* X = gauss_jordan(ws, k, 2 * k);
    if (!X) return false;
    KS_PAR2(i, j, k, k) { ws->Mi[i * LD + j] = X[i * LW + k + j]; } KS_END2
    KS_SYNC();
    return true;
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
    const bool dogleg = st->dogleg[0] != 0.0;
    KS_SYNC();                          // all threads hold the inputs before the master resets the report
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
    double* wv = ws->vec + 2 * M;
    if (k == 0) {  // optimizer_utils.py:197-215
        const double delT = rT(delta, f32);
        const double scale = rT(rT(ddiv(1.0, gn), f32) * delT, f32);  // python: delta / gn  ==  gn.reciprocal() * delta
        cg = -scale;
        pred = rT(delT * gn, f32);
        pp = cg * cg * gg;
        gp = cg * gg;
        KS_MASTER st->case_id[0] = DD4ML_CASE_FIRST_ORDER;
    } else {
        double *b = ws->vec, *x = ws->vec + M, *Gw = ws->vec + 3 * M, *z = ws->vec + 4 * M;
        KS_PAR(i, k) { const int pi = (int)st->order[i]; b[i] = st->Yg[pi] - gam * st->Sg[pi]; }  // Psi^T g
        const bool okm = build_compact(st, ws);                 // (its barriers publish b)
        if (!(okm && finite(gg) && finite(delta) && finite(gam) && finite(dotk(b, b, k)))) KS_FAIL(DD4ML_ERR_NONFINITE)
        if (delta < 0) KS_FAIL(DD4ML_ERR_NEG_DELTA)

        // ---- OBS (obs.py:61-95): eigen-decomposition of R Minv^{-1} R^T, R^T R = Psi^T Psi ----
        if (!chol_aug(ws, k)) KS_FAIL(DD4ML_ERR_CHOLESKY)
        if (!invert_minv(ws, k)) KS_FAIL(DD4ML_ERR_SINGULAR)
        mm(ws->T, LD, ws->Mi, LD, 1, ws->R, 1, LD, k, k, k);   // T = Minv^{-1} R^T
        mm(ws->J1, LD, ws->R, LD, 1, ws->T, LD, 1, k, k, k);   // R Minv^{-1} R^T
        KS_PAR2(i, j, k, k) { ws->J0[i * LD + j] = 0.5 * (ws->J1[i * LD + j] + ws->J1[j * LD + i]); } KS_END2
        KS_SYNC();
        const double c_e0 = KS_CLOCK();
        double *A, *V;
        const int sweeps = jacobi_core(ws, k, &A, &V);
        KS_MASTER { st->jacobi_sweeps[0] = sweeps; st->cyc_eigh[0] = KS_CLOCK() - c_e0; }
        double* Lam = ws->Ld;               // Lambda, a as T values held in doubles (+ float copies for T = float)
        double* a = ws->Ad;
        const double tolT = rT(OBS_TOL, f32), gamT = rT(gam, f32), delT = rT(delta, f32);
        KS_PAR(c, k) {  // eigenvalue c -> ascending rank r, sign convention, a_r = u_r . (R^-T Psi^T g)
            const double wc = A[c * LD + c];
            int r = 0, im = 0;
            double best = -1.0;
            KS_LOOP
            for (int i = 0; i < k; ++i) {
                const double wi = A[i * LD + i], vi = fabs(V[i * LD + c]);
                r += (wi < wc || (wi == wc && i < c)) ? 1 : 0;
                if (vi > best) { best = vi; im = i; }
            }
            double s = dots(V + c, LD, ws->R + k, LD, k);
            if (V[im * LD + c] < 0.0) s = -s;
            ws->ad[r] = s;
            a[r] = rT(s, f32);
            double L = rT(rT(wc, f32) + gamT, f32);
            if (fabs(L) < tolT) L = 0.0;
            Lam[r] = L;
            ws->Af[r] = (float)a[r];
            ws->Lf[r] = (float)L;
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
        KS_PAR2(i, j, k, k + 1) {
            ws->X0[i * LW + j] = (j < k) ? tau * ws->Minv[i * LD + j] + ws->G[i * LD + j] : b[i];
        } KS_END2
        KS_SYNC();
        const double* X = gauss_jordan(ws, k, k + 1);       // wv = (tau Minv + G) \ Psi^T g
        if (!X) KS_FAIL(DD4ML_ERR_SINGULAR)
        KS_PAR(i, 2 * k) {
            if (i < k) wv[i] = X[i * LW + k];
            else x[i - k] = dots(ws->Mi + (i - k) * LD, 1, b, 1, k);      // x = Minv \ Psi^T g
        }
        KS_SYNC();
        mm(Gw, 1, ws->G, LD, 1, wv, 1, 0, k, 1, k);         // G wv
        mm(z, 1, ws->Mi, LD, 1, Gw, 1, 0, k, 1, k);         // z = Minv \ (G wv)
        KS_PAR(q, 6) {  // b.x  b.wv  wv.Gw  b.z  Gw.x  Gw.z   (vec blocks: 0 b, 1 x, 2 wv, 3 Gw, 4 z)
            const int ia = (0x330200 >> (4 * q)) & 15, ib = (0x414321 >> (4 * q)) & 15;
            ws->dot6[q] = dots(ws->vec + ia * M, 1, ws->vec + ib * M, 1, k);
        }
        KS_SYNC();
        const double bx = ws->dot6[0], bw = ws->dot6[1], wGw = ws->dot6[2];
        const double bz = ws->dot6[3], gwx = ws->dot6[4], gwz = ws->dot6[5];
        const double cgb = rT(ddiv(-1.0, tauT), f32);
#define KS_PP(c_, l_) ((c_) * (c_) * gg + 2.0 * (c_) * (l_) * bw + (l_) * (l_) * wGw)
#define KS_GP(c_, l_) ((c_) * gg + (l_) * bw)
        // ---- Cauchy point and dogleg (optimizer_utils.py:259-296) ----
        const double gBg = gam * gg + bx;
        KS_MASTER st->gBg[0] = gBg;
        cg = cgb;
        lam = 1.0;
        double branch = 0.0;
        if (dogleg && gBg > 0.0) {
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
        const double c = lam * wv[i];
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
        double *bs = ws->vec, *by = ws->vec + M, *x = ws->vec + 2 * M, *Gx = ws->vec + 3 * M;
        KS_PAR(i, k) {
            const int pi = (int)st->order[i];
            bs[i] = st->c_Ys[pi] - gam * st->c_Ss[pi];   // Psi^T s
            by[i] = st->c_Yy[pi] - gam * st->c_Sy[pi];   // Psi^T y
        }
        build_compact(st, ws);
        KS_PAR2(i, j, k, k + 1) { ws->X0[i * LW + j] = (j < k) ? ws->Minv[i * LD + j] : bs[i]; } KS_END2
        KS_SYNC();
        const double* X = gauss_jordan(ws, k, k + 1);    // x = Minv \ Psi^T s
        if (!X) { KS_MASTER st->err[0] = DD4ML_ERR_SINGULAR; KS_SYNC(); return false; }
        KS_PAR(i, k) x[i] = X[i * LW + k];
        KS_SYNC();
        mm(Gx, 1, ws->G, LD, 1, x, 1, 0, k, 1, k);
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
    KS_SYNC();                          // every thread has read the old ring before it changes
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
    KS_PAR2(i, j, k, k + 1) {
        const int pi = (int)st->order[i];
        ws->X0[i * LW + j] = (j == k) ? st->Yg[pi] - gam * st->Sg[pi]
                                      : (smw ? tau * ws->Minv[i * LD + j] + ws->G[i * LD + j] : ws->Minv[i * LD + j]);
    } KS_END2
    KS_SYNC();
    const double* X = gauss_jordan(ws, k, k + 1);
    if (!X) { KS_MASTER st->err[0] = DD4ML_ERR_SINGULAR; KS_SYNC(); return; }
    KS_PAR(i, k) {
        const int pi = (int)st->order[i];
        st->cY[pi] = X[i * LW + k];
        st->cS[pi] = -gam * X[i * LW + k];
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
