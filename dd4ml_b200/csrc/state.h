// Device-resident optimizer state block shared by all dd4ml_b200 kernels.
//
// One block per optimizer instance (TR or LSSR1_TR).  Every field is a double so
// that the host can mirror the block as a float64 tensor / numpy array; integer
// quantities (k, flags, slot indices) are stored as exact small doubles.
//
// The L-SR1 memory itself (S, Y) lives in two pair-major ring buffers of
// DD4ML_MAXS physical slots x ld elements each; this block
// holds only k-space data: the slot order, gamma and the Gram blocks
// S^T S, S^T Y, Y^T Y indexed by PHYSICAL slot.
#pragma once

#define DD4ML_MAXM 10                   // largest supported memory length
#define DD4ML_MAXS (DD4ML_MAXM + 1)     // physical slots = live pairs + 1 spare (candidate)
#define DD4ML_G (DD4ML_MAXS * DD4ML_MAXS)

// X(name, count)
#define DD4ML_STATE_FIELDS(X)                                                         \
    /* ---- hyper-parameters (host-written) ---- */                                   \
    X(delta, 1) X(min_delta, 1) X(max_delta, 1) X(nu_dec, 1) X(nu_inc, 1)             \
    X(inc_factor, 1) X(dec_factor, 1) X(tol, 1) X(norm_inf, 1) X(second_order, 1)     \
    X(dogleg, 1) X(mem_len, 1)                                                        \
    /* ---- report block (device-written, host reads [0, report_end)) ---- */         \
    X(gn, 1) X(gg, 1) X(ginf, 1) X(pred, 1) X(pnorm, 1) X(gp, 1) X(case_id, 1)        \
    X(sigma, 1) X(sigma0, 1) X(tau, 1) X(newton_iters, 1) X(newton_deg, 1) X(err, 1)  \
    X(converged, 1) X(k, 1) X(gamma, 1) X(rho, 1) X(accept, 1) X(pair_tried, 1)       \
    X(pair_ok, 1) X(loss, 1) X(trial_loss, 1) X(out_gn, 1) X(out_gg, 1)               \
    X(out_ginf, 1) X(dphi0, 1) X(flip, 1) X(clip_scale, 1) X(psq, 1) X(lam_min, 1)    \
    X(gBg, 1) X(dogleg_branch, 1) X(deriv, 1) X(eval_loss, 1) X(pmax, 1)              \
    X(foc_dd, 1) X(foc_rd, 1) X(steps, 1) X(solve_valid, 1) X(gen, 1) X(aux, 1)      \
    X(cyc_solve, 1) X(cyc_eigh, 1) X(cyc_newton, 1) X(cyc_update, 1) X(jacobi_sweeps, 1) \
    /* err_seen: last non-zero err since the host cleared it (sync-free modes read late);       */ \
    /* inner_err: err_seen of the global / local optimizer, linked into the APTS control block  */ \
    X(err_seen, 1) X(inner_err, 2)                                                    \
    X(report_end, 1)                                                                  \
    /* ---- L-SR1 ring bookkeeping and Gram blocks ---- */                            \
    X(spare, 1) X(order, DD4ML_MAXM)                                                  \
    X(SS, DD4ML_G) X(SY, DD4ML_G) X(YY, DD4ML_G)                                      \
    /* ---- current sub-problem: reductions and step descriptor ---- */               \
    X(Sg, DD4ML_MAXS) X(Yg, DD4ML_MAXS)                                               \
    X(cg, 1) X(cS, DD4ML_MAXS) X(cY, DD4ML_MAXS)                                      \
    /* ---- candidate pair reductions (decide / begin kernels) ---- */                \
    X(c_ss, 1) X(c_sy, 1) X(c_yy, 1) X(c_sg, 1) X(c_yg, 1)                            \
    X(c_Ss, DD4ML_MAXS) X(c_Ys, DD4ML_MAXS) X(c_Sy, DD4ML_MAXS) X(c_Yy, DD4ML_MAXS)                  \
    /* ---- caller-supplied M^{-1} (OBS.solve_tr_subproblem / ComputeSBySMW with explicit Psi, Minv:  */ \
    /* obs.py:46,175): logical order, leading dimension DD4ML_MAXM; used instead of the Gram-derived  */ \
    /* matrix while use_mx != 0                                                                       */ \
    X(use_mx, 1) X(Mx, DD4ML_MAXM * DD4ML_MAXM)

struct dd4ml_state {
#define X(name, count) double name[count];
    DD4ML_STATE_FIELDS(X)
#undef X
};

// error codes stored in state.err (0 = ok)
#define DD4ML_ERR_NONE 0
#define DD4ML_ERR_CHOLESKY 1        // Psi^T Psi not positive definite (torch.linalg.cholesky would raise)
#define DD4ML_ERR_SINGULAR 2        // LU pivot exactly zero (torch.linalg.solve would raise)
#define DD4ML_ERR_NONFINITE 3       // NaN/Inf in g, delta, gamma, Gram blocks (obs.py:48-57 ValueError)
#define DD4ML_ERR_HARD_CASE 4       // OBS hard case (obs.py:100-158); unreachable for gamma > 0
#define DD4ML_ERR_NEG_DELTA 5       // delta < 0 (obs.py:58)

// OBS case ids stored in state.case_id
#define DD4ML_CASE_FIRST_ORDER 0
#define DD4ML_CASE_A 1
#define DD4ML_CASE_C 3
