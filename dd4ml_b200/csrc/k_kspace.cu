// The one-CTA k-space kernel (1 warp for memories of up to 5 pairs, 4 warps above): see kspace_task.h and kspace.h.
#include <stdint.h>

#include "kspace.h"
#include "kspace_task.h"

namespace {

// OBS works in the dtype of Psi when pairs are stored (obs.py:88), the first-order step in the dtype
// of the gradient.  ONE call site per instantiation: the routine is large and must exist once.
__device__ __noinline__ void solve_any(dd4ml_state* st, int mode, int ht, int vt, ks::Scratch* scr) {
    const int tt = (st->second_order[0] != 0.0 && st->k[0] > 0.0) ? ht : vt;
    ks::tr_solve(st, mode, tt == 0, scr);
}

__device__ __noinline__ void finish_any(dd4ml_state* st, int lt, bool accept, double rho, ks::Scratch* scr) {
    ks::tr_finish(st, accept, rho, lt == 0, scr);
}

__global__ void __launch_bounds__(128) kspace_kernel(dd4ml_state* gst, KTask t) {
    __shared__ dd4ml_state st;
    __shared__ ks::Scratch scr;
    constexpr int NST = (int)(sizeof(dd4ml_state) / sizeof(double));
    // programmatic dependent launch: everything above ran while the reduction kernel was still
    // streaming; its results (the reduction fields of the state block) are visible after this
    asm volatile("griddepcontrol.wait;" ::: "memory");
    double* g = reinterpret_cast<double*>(gst);
    double* s = reinterpret_cast<double*>(&st);
    for (int i = threadIdx.x; i < NST; i += blockDim.x) s[i] = __ldcg(g + i);
    __syncthreads();
    int mode = 0;
    bool do_solve = false, restore_ninf = false;
    double ninf_saved = 0.0;
    switch (t.kind) {
        case KT_SOLVE:
            mode = t.mode;
            do_solve = true;
            break;
        case KT_B:
            ks::lsr1_B_coefs(&st, &scr);
            break;
        case KT_SMW:
            ks::smw_coefs(&st, &scr);
            break;
        case KT_BEGIN: {
            // lssr1_tr.py:509: the convergence test precedes the memory update
            const double gg = st.gg[0], tol = st.tol[0];
            const double gn = t.vt == 0 ? (double)(float)sqrt(gg) : sqrt(gg);
            const bool pair = t.has_prev && st.second_order[0] != 0.0;
            const int spare = (int)st.spare[0];
            const double ninf = st.norm_inf[0];
            __syncthreads();
            if (pair && gn > tol && sqrt(st.c_ss[0]) > tol && sqrt(st.c_yy[0]) > tol) {  // lssr1_tr.py:521
                if (threadIdx.x == 0) st.pair_tried[0] = 1.0;
                if (ks::lsr1_try_update(&st, &scr)) {
                    if (threadIdx.x == 0) { st.Sg[spare] = st.c_sg[0]; st.Yg[spare] = st.c_yg[0]; }
                }
            }
            if (threadIdx.x == 0) st.norm_inf[0] = 0.0;      // LSSR1_TR forces the 2-norm (lssr1_tr.py:94)
            __syncthreads();
            mode = 1;
            do_solve = true;
            restore_ninf = true;
            ninf_saved = ninf;
            break;
        }
        case KT_DECIDE: {
            if (st.converged[0] != 0.0) break;               // tr.py:147: nothing happened (flags set by the kernel)
            double rho = 0.0;
            const bool accept = (t.lt == 0)
                ? ks::tr_accept<float>(st.loss[0], st.trial_loss[0], st.pred[0], st.tol[0], st.nu_dec[0], &rho)
                : ks::tr_accept<double>(st.loss[0], st.trial_loss[0], st.pred[0], st.tol[0], st.nu_dec[0], &rho);
            const double so = st.second_order[0];
            __syncthreads();
            if (!t.mem && threadIdx.x == 0) st.second_order[0] = 0.0;
            __syncthreads();
            finish_any(&st, t.lt, accept, rho, &scr);
            if (threadIdx.x == 0) { st.second_order[0] = so; st.solve_valid[0] = 0.0; }
            break;
        }
        case KT_UPDATE:
            if (threadIdx.x == 0) { st.err[0] = 0.0; st.pair_tried[0] = 1.0; }
            __syncthreads();
            ks::lsr1_try_update(&st, &scr);
            if (threadIdx.x == 0) st.solve_valid[0] = 0.0;
            break;
        default:
            break;
    }
    if (do_solve) {                     // the one call site of the sub-problem solve
        solve_any(&st, mode, t.ht, t.vt, &scr);
        if (restore_ninf && threadIdx.x == 0) st.norm_inf[0] = ninf_saved;
    }
    __syncthreads();
    if (threadIdx.x == 0 && st.err[0] != 0.0) st.err_seen[0] = st.err[0];
    __syncthreads();
    for (int i = threadIdx.x; i < NST; i += blockDim.x) g[i] = s[i];
}

}  // namespace

static bool g_kspace_enabled = true;
// Measurement hook (bench.py / tools): with the k-space kernel switched off an entry point launches only its
// streaming reduction kernel, so that kernel can be timed alone with CUDA events (its results are then NOT
// post-processed: never switch it off outside a micro-benchmark).
extern "C" int dd4ml_debug_kspace(int on) {
    const int was = g_kspace_enabled ? 1 : 0;
    g_kspace_enabled = on != 0;
    return was;
}

int dd4ml_ks_launch(double* state, const KTask& task, cudaStream_t stream, int kcap) {
    if (!g_kspace_enabled) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(kcap > 5 ? 128 : 32);   // parallel regions are k x k .. k x 2k elements wide
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kspace_kernel, reinterpret_cast<dd4ml_state*>(state), task);
}
