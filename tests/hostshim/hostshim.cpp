// TEST INFRASTRUCTURE ONLY: compiles dd4ml_b200/csrc/kspace.h for the host so that the
// k-space routines (the exact source the one-warp k-space kernel runs; its parallel regions
// become plain loops here, see the execution model in kspace.h) can be checked against the
// oracle without a GPU.  Never loaded by the product package.
#include "../../dd4ml_b200/csrc/kspace.h"
#include "../../dd4ml_b200/csrc/state_table.h"

extern "C" {
int hostshim_state_doubles() { return (int)(sizeof(dd4ml_state) / sizeof(double)); }
int hostshim_num_fields() { return DD4ML_NUM_FIELDS; }
int hostshim_field(int i, const char** name, int* offset, int* count) {
    if (i < 0 || i >= DD4ML_NUM_FIELDS) return 1;
    *name = DD4ML_FIELD_TABLE[i].name;
    *offset = DD4ML_FIELD_TABLE[i].offset;
    *count = DD4ML_FIELD_TABLE[i].count;
    return 0;
}
void hostshim_tr_solve(double* st, int mode, int is_float) {
    ks::Scratch ws;
    if (is_float) ks::tr_solve<float>((dd4ml_state*)st, mode, &ws);
    else ks::tr_solve<double>((dd4ml_state*)st, mode, &ws);
}
int hostshim_try_update(double* st) { ks::Scratch ws; return ks::lsr1_try_update((dd4ml_state*)st, &ws) ? 1 : 0; }
void hostshim_tr_finish(double* st, double loss, double trial, int is_float) {
    dd4ml_state* s = (dd4ml_state*)st;
    double rho;
    bool acc = is_float ? ks::tr_accept<float>(loss, trial, s->pred[0], s->tol[0], s->nu_dec[0], &rho)
                        : ks::tr_accept<double>(loss, trial, s->pred[0], s->tol[0], s->nu_dec[0], &rho);
    ks::Scratch ws;
    if (is_float) ks::tr_finish<float>(s, acc, rho, &ws); else ks::tr_finish<double>(s, acc, rho, &ws);
}
int hostshim_apts_control(double* st, double f0, double ft, double pred, int is_float) {
    return (is_float ? ks::apts_control<float>((dd4ml_state*)st, f0, ft, pred)
                     : ks::apts_control<double>((dd4ml_state*)st, f0, ft, pred)) ? 1 : 0;
}
void hostshim_smw_coefs(double* st) { ks::Scratch ws; ks::smw_coefs((dd4ml_state*)st, &ws); }
void hostshim_B_coefs(double* st) { ks::Scratch ws; ks::lsr1_B_coefs((dd4ml_state*)st, &ws); }
// OBS.phiBar_fg (mode 1) / Newton (mode 2) in float or double; out = 3 doubles
void hostshim_secular(int mode, double x0, const double* Lam, const double* a, int m, double delta, double tol,
                      int is_float, double* out) {
    if (is_float) {
        float L[ks::M + 1], A[ks::M + 1];
        for (int j = 0; j < m; ++j) { L[j] = (float)Lam[j]; A[j] = (float)a[j]; }
        if (mode == 2) { int it = 0, nd = 0; out[0] = ks::newton<float>((float)x0, L, A, m, (float)delta, (float)tol, it, nd); out[1] = it; out[2] = nd; }
        else { float f, g; bool d; ks::phibar_fg<float>((float)x0, L, A, m, (float)delta, (float)tol, f, g, d); out[0] = f; out[1] = g; out[2] = d; }
    } else {
        if (mode == 2) { int it = 0, nd = 0; out[0] = ks::newton<double>(x0, Lam, a, m, delta, tol, it, nd); out[1] = it; out[2] = nd; }
        else { double f, g; bool d; ks::phibar_fg<double>(x0, Lam, a, m, delta, tol, f, g, d); out[0] = f; out[1] = g; out[2] = d; }
    }
}
int hostshim_eigh(const double* A, int k, double* w, double* V) {
    double a[ks::M * ks::LD], v[ks::M * ks::LD], ww[ks::M];
    ks::Scratch ws;
    for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) a[i * ks::LD + j] = A[i * k + j];
    const int sweeps = ks::jacobi_eigh(a, k, ww, v, &ws);
    for (int i = 0; i < k; ++i) { w[i] = ww[i]; for (int j = 0; j < k; ++j) V[i * k + j] = v[i * ks::LD + j]; }
    return sweeps;
}
}
