/* dd4ml_b200 -- C ABI of the B200-native trust-region hot path of cruzas/DD4ML.
 *
 * The reference is pure Python/PyTorch, so there is no FFI to replace; each entry
 * point below is the fused device-side replacement of the ATen op sequence of one
 * reference function on the hot path (file:line cited per function, paths relative
 * to the reference's src/dd4ml/).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer borrowed from
 *     the caller (PyTorch storages); the library allocates nothing.
 *   - dtype codes: 0 = float32, 1 = float64.  vt = dtype of gradient/step vectors,
 *     ht = dtype of the L-SR1 pairs, wt = dtype of the parameters, lt = dtype of the
 *     0-d loss tensors.  Supported (ht,vt,wt): (0,0,0) (0,0,1) (0,1,1) (1,1,1).
 *   - `state`  : device block of dd4ml_state_doubles() doubles (csrc/state.h), one per
 *     optimizer instance.  `ws`: device workspace of dd4ml_workspace_doubles() doubles.
 *   - S, Y     : L-SR1 pair-major ring buffers, DD4ML_MAXS slots of `ld` elements
 *     (ld >= n, ld % 4 == 0), slot j at S + j*ld.
 *   - kcap    : the optimizer's memory length (<= DD4ML_MAXM); selects the kernel variant
 *     unrolled for 4 or 8 live pairs.  The number of pairs actually stored is device state.
 *   - all work is enqueued on `stream` (a cudaStream_t); nothing synchronises.
 *   - return value: 0 on success, otherwise a cudaError_t or a DD4ML_E_* code
 *     (dd4ml_error_string).  Numerical failures (Cholesky, singular solve, NaN) are
 *     reported through state.err, read by the host with the report block.
 */
#ifndef DD4ML_B200_H
#define DD4ML_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD4ML_E_DTYPE 10001      /* unsupported dtype combination */
#define DD4ML_E_ARG 10002        /* bad argument (n < 0, ld, null pointer, mem_length > DD4ML_MAXM) */
#define DD4ML_E_ALIGN 10003      /* reserved */

int dd4ml_version(void);
const char* dd4ml_error_string(int code);

/* layout of the state block (name/offset/count in doubles) and sizes */
int dd4ml_state_doubles(void);
int dd4ml_workspace_doubles(void);
int dd4ml_num_fields(void);
int dd4ml_field(int i, const char** name, int* offset, int* count);
int dd4ml_max_mem_length(void);

/* zero the state + workspace and install defaults (gamma, spare slot) */
int dd4ml_state_init(double* state, double* ws, double gamma0, void* stream);
/* state[offsets[i]] = values[i], i < count <= 32 (hyper-parameters, delta written by the host) */
int dd4ml_state_set(double* state, int count, const int* offsets, const double* values, void* stream);
/* dst[dst_off] = clamp(src[src_off] * scale, lo, hi): radius hand-over between the APTS, global and local
 * optimizer state blocks on the device (apts_base.py:228-237, apts_d.py:202, apts_p.py:169,236) */
int dd4ml_state_link(double* dst, int dst_off, const double* src, int src_off, double scale, double lo, double hi,
                     void* stream);

/* ---- TR.step, optimizers/tr.py:138-222 -------------------------------------------------
 * propose: ||g|| (2 or inf), S^T g, Y^T g in one pass; the k-space kernel enqueued behind it then solves the
 *   sub-problem in k-space (utility/optimizer_utils.py:197-307, optimizers/lsr1.py:146-198,
 *   solvers/obs.py:46-254) and leaves the step descriptor in `state`.
 *   mode 0 = TR, 1 = LSSR1_TR (clip + descent flip, lssr1_tr.py:563-587), 2 = coefficients of
 *   LSR1.B(v) (g = v), 3 = OBS.ComputeSBySMW for tau = state.tau (obs.py:175-182).
 * apply:   p = cg*g + sum_j cS_j S_j + cY_j Y_j;  optionally w += p  (tr.py:106-131,179).
 * decide:  rho test (tr.py:186-191); accepted: y = g_trial - g_old, candidate pair
 *   reductions, g_out <- g_trial, L-SR1 update chain + radius update in the k-space kernel
 *   (lsr1.py:53-144, tr.py:193-214); rejected: w -= p, radius shrink (tr.py:216-222).
 *   loss_out (nullable): device-side select of what TR.step returns -- loss_out <- accepted ?
 *   trial_loss : loss and, when g_out != g_old, g_out <- g_old on rejection -- so the host needs no
 *   read-back per inner iteration. */
int dd4ml_tr_propose(double* state, double* ws, const void* g, const void* S, const void* Y,
                     int64_t n, int64_t ld, int vt, int ht, int mode, int kcap, void* stream);
int dd4ml_tr_apply(double* state, double* ws, const void* g, const void* S, const void* Y, void* p,
                   void* w, int64_t n, int64_t ld, int vt, int ht, int wt, int kcap, void* stream);
int dd4ml_tr_decide(double* state, double* ws, const void* loss, const void* trial_loss, int lt,
                    const void* g_trial, const void* g_old, void* g_out, const void* p, void* w,
                    void* S, void* Y, void* loss_out, int64_t n, int64_t ld, int vt, int ht, int wt, int kcap,
                    void* stream);

/* ---- LSSR1_TR.step, optimizers/lssr1_tr.py:493-654 --------------------------------------
 * begin: s = w - old_w, y = g - prev_g (lssr1_tr.py:518-526) with every reduction of the
 *   L-SR1 update chain and of the sub-problem in one pass; old_w_out <- w, prev_g_out <- g (separate
 *   buffers: the reference keeps old_wk / prev_grad when ||g|| <= tol, lssr1_tr.py:509-526, which is known
 *   only after the pass -- the caller swaps the buffer pairs unless state.converged); the k-space
 *   kernel runs update_memory, precompute and the sub-problem solve; `loss` (phi_0) is copied into
 *   the report block so the host needs no separate read.  delta/min_delta/max_delta: the radius
 *   LSSR1_TR updated on the host after the previous step (lssr1_tr.py:621-652), stored into the
 *   state block before the solve; delta < 0 keeps the block's values.
 * trial: w = base + alpha * p (lssr1_tr.py:262).
 * eval:  deriv = g . p, g_save <- g, loss copied into the report (lssr1_tr.py:265-271). */
int dd4ml_lssr1_begin(double* state, double* ws, const void* w, const void* old_w, void* old_w_out,
                      const void* g, const void* prev_g, void* prev_g_out, int has_prev, void* S, void* Y,
                      const void* loss, int lt, int64_t n, int64_t ld, int vt, int ht, int wt, int kcap,
                      double delta, double min_delta, double max_delta, void* stream);
int dd4ml_lssr1_trial(void* w, const void* base, const void* p, double alpha, int64_t n, int vt,
                      int wt, void* stream);
int dd4ml_lssr1_eval(double* state, double* ws, const void* g, const void* p, void* g_save,
                     const void* loss, int lt, int64_t n, int vt, void* stream);

/* ---- APTS, optimizers/apts_base.py / apts_d.py / apts_p.py -------------------------------
 * residual: g0 <- Gg, gl0 <- Gl, resid <- Gg - Gl  (apts_d.py:179-186).
 * foc:      loss_out = loss_in + resid.(w_loc - w_glob) if ||w_loc - w_glob|| > tol
 *           (apts_base.py:262-292).
 * pack:     buf[0:n] = w_loc - w0, buf[n] = (loss0 - loss1)*weight, buf[n+1] = evals
 *           (apts_d.py:195-196,139-141; apts_base.py:411-416) -- the all-reduce payload.
 * trial:    mode 0: w = w0 + buf*scale (apts_base.py:445, apts_d.py:149-150);
 *           mode 1: w = w0 + clamp(buf - w0, -clampv, clampv) (apts_p.py:221-224).
 *           `bt` = dtype of the reference's pre-allocated `_step_buffer` / coalesced buffer
 *           (apts_d.py:112,129-141, apts_p.py:108), i.e. the dtype the model had when the optimizer
 *           was constructed: pack writes buf[0:n+2] in bt, trial mode 0 reads the step in bt and
 *           mode 1 (buf = merged parameters, dtype wt) rounds the step and the clamp bound to bt.
 *           bt == wt unless the model was converted after construction (trainer.py:1203-1211).
 * control:  APTS_Base.control_step decision (apts_base.py:483-501): accepted: g0 <- G
 *           (trial gradient), loss_out <- f_trial; rejected: w <- w0, loss_out <- f0;
 *           radius update on `state`; `aux` (nullable, dtype pt) is copied into state.aux so the
 *           summed evaluation counter of the same all-reduce rides along with the report. */
int dd4ml_apts_residual(const void* Gg, const void* Gl, void* g0, void* gl0, void* resid,
                        int64_t n, int vt, void* stream);
int dd4ml_apts_foc(double* state, double* ws, const void* w_loc, const void* w_glob,
                   const void* resid, const void* loss_in, void* loss_out, int lt, double tol,
                   int64_t n, int vt, int wt, void* stream);
int dd4ml_apts_pack(void* buf, const void* w_loc, const void* w0, const void* loss0,
                    const void* loss1, int lt, double weight, double evals, int64_t n, int wt,
                    int bt, void* stream);
int dd4ml_apts_trial(void* w, const void* w0, const void* buf, int mode, double scale,
                     double clampv, int64_t n, int wt, int bt, void* stream);
/* combine: the additive-Schwarz combination fused with its exchange over NVLink peer memory
 *           (apts_d.py:139-150 all-reduce of [step || pred] + apts_base.py:445 w = w0 + step): `peer_bufs`
 *           is a HOST array of `nranks` DEVICE pointers to the packed buffers (dtype bt, n+2 elements,
 *           written by dd4ml_apts_pack) of all ranks in rank order -- symmetric-memory / P2P-mapped
 *           addresses.  w = w0 + scale * sum_r buf_r[0:n] (rank-ordered float64 accumulation: identical
 *           on every rank), tail_out[0:2] (dtype bt) = sum_r buf_r[n:n+2].  The caller orders it against
 *           the peers' pack kernels with a device-side barrier.  Replaces pack -> ncclAllReduce -> trial. */
int dd4ml_apts_combine(void* w, const void* w0, const void* const* peer_bufs, int nranks, double scale,
                       void* tail_out, int64_t n, int wt, int bt, void* stream);
int dd4ml_apts_control(double* state, double* ws, const void* f0, const void* ftrial, int lt,
                       const void* pred, int pt, void* w, const void* w0, void* g0, const void* G,
                       void* loss_out, const void* aux, int64_t n, int vt, int wt, void* stream);

/* ---- TRAdam (optimizers/tradam.py:80-111): Adam direction clipped to the radius `lr` ---------
 * moments: m <- b1 m + (1-b1) g, v <- b2 v + (1-b2) g^2 in place; s = (m/bc1)/(sqrt(v/bc2)+eps)
 *          is reduced to its length (max|s| if norm_inf, else s.s -- not its root, tradam.py:101);
 *          state.pnorm <- length, state.aux <- scale = lr/length if length > lr else lr.
 *          bc1 = 1-b1^t, bc2 = 1-b2^t.  m, v, g of dtype vt.
 * apply:   step = scale * s (recomputed from m, v), w <- w - step; `step` (nullable, dtype vt)
 *          receives the applied step (the reference's _step_buf). */
int dd4ml_tradam_moments(double* state, double* ws, const void* g, void* m, void* v, double beta1,
                         double beta2, double bc1, double bc2, double eps, double lr, int norm_inf,
                         int64_t n, int vt, void* stream);
int dd4ml_tradam_apply(double* state, double* ws, const void* m, const void* v, void* w, void* step,
                       double bc1, double bc2, double eps, int64_t n, int vt, int wt, void* stream);

/* ---- flat-buffer plumbing (utility/apts_utils.py:17-62, tr.py:88-104) ---------------------
 * gather:   dst[off_i : off_i+len_i] = src_i  for up to 96 tensors per call (multi-tensor).
 * seg_copy: dst[dst_off_i : +len_i] = src[src_off_i : +len_i]; optional zero fill of dst
 *           first (apts_p.py:145-150). */
int dd4ml_gather(void* dst, int count, const void* const* srcs, const int64_t* offsets,
                 const int64_t* lens, int dtype, void* stream);
int dd4ml_seg_copy(void* dst, const void* src, int count, const int64_t* dst_off,
                   const int64_t* src_off, const int64_t* lens, int64_t n_dst, int zero_fill,
                   int dtype, void* stream);

/* ---- OBS helpers as stand-alone calls (solvers/obs.py:184-254) ------------------------------
 * mode 0/1: phiBar_f / phiBar_fg at sigma = x0 -> out = {phiBar, phiBar', degenerate flag};
 * mode 2:   Newton from x0 (<= 200 iterations, |phiBar| <= tol) -> out = {sigma*, iterations,
 *           degenerate evaluations}.  Lam, a: m = k+1 DEVICE values of dtype `dtype`; out: 3 device
 *           doubles.  On the hot path these run inside tr_propose / lssr1_begin.
 * ComputeSBySMW (obs.py:175-182) is tr_propose(mode = 3) with state.tau set, followed by tr_apply. */
int dd4ml_obs_secular(int mode, double x0, const void* Lam, const void* a, int m, double delta,
                      double tol, int dtype, double* out, void* stream);

/* ---- GPU-resident batch feed (dataloaders.py:18-229 samplers, trainer.py:356-382) ----------
 * gather_rows: dst[r, :] = src[idx[r], :] for r < nrows; rows are row_bytes bytes, idx is a
 *              DEVICE int64 vector (one epoch of sampler batches is uploaded once); rows whose
 *              index falls outside [0, src_rows) are left untouched. */
int dd4ml_gather_rows(void* dst, const void* src, const int64_t* idx, int64_t nrows,
                      int64_t row_bytes, int64_t src_rows, void* stream);

/* ---- LSR1.update_memory, optimizers/lsr1.py:53-144: candidate reductions in one pass (the
 * pair is written to the spare ring slot), acceptance chain + Gram/gamma update in the k-space kernel */
int dd4ml_lsr1_update(double* state, double* ws, const void* s, const void* y, void* S, void* Y,
                      int64_t n, int64_t ld, int vt, int ht, int kcap, void* stream);

/* ---- LSR1.B, optimizers/lsr1.py:181-198: out = gamma v + Psi (Minv \ Psi^T v) -------------
 * two launches: reduce (S^T v, Y^T v) + k-space kernel, then the output pass. */
int dd4ml_lsr1_B(double* state, double* ws, const void* v, const void* S, const void* Y, void* out,
                 int64_t n, int64_t ld, int vt, int ht, int kcap, void* stream);

/* ---- measurement hook: switch the one-CTA k-space kernel that follows tr_propose / lssr1_begin /
 * tr_decide / lsr1_update off (0) or on (1) so that the streaming reduction kernel can be timed alone;
 * returns the previous setting.  Results are not post-processed while it is off. */
int dd4ml_debug_kspace(int on);

#ifdef __cplusplus
}
#endif
#endif
