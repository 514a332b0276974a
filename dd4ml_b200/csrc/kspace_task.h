// The k-space kernel (k_kspace.cu) and its launch: ONE CTA (1 or 4 warps) that runs the O(k^2)-input part of a
// trust-region step (L-SR1 update chain, OBS / dogleg solve, ratio test bookkeeping) on the device
// state block, right behind the streaming reduction kernel that filled the block's reduction fields.
// It is launched with programmatic dependent launch: the reduction kernels signal
// `griddepcontrol.launch_dependents` at their start, so this kernel is resident (launch latency paid,
// first instructions fetched) while they still stream, and `griddepcontrol.wait`s for their results.
#pragma once
#include <cuda_runtime.h>

enum {
    KT_SOLVE = 0,    // tr_propose mode 0 / 1: sub-problem solve (tr.py:150-176 / lssr1_tr.py:540-587)
    KT_B = 1,        // LSR1.B coefficients (lsr1.py:181-198)
    KT_SMW = 2,      // OBS.ComputeSBySMW coefficients for state.tau (obs.py:175-182)
    KT_BEGIN = 3,    // lssr1_begin: update_memory chain + solve (lssr1_tr.py:509-587)
    KT_DECIDE = 4,   // tr_decide: radius + L-SR1 update after the trial evaluation (tr.py:186-221)
    KT_UPDATE = 5,   // standalone LSR1.update_memory (lsr1.py:53-144)
};

struct KTask {
    int kind;
    int mode;        // KT_SOLVE: 0 TR, 1 LSSR1_TR
    int ht, vt, lt;  // dtypes (0 float, 1 double): L-SR1 pairs, vectors, loss scalars
    int has_prev;    // KT_BEGIN: a previous iterate exists (candidate pair was reduced)
    int mem;         // KT_DECIDE: the optimizer keeps an L-SR1 memory
};

// enqueue the k-space kernel for `state` behind the work already in `stream`; returns a cudaError_t
int dd4ml_ks_launch(double* state, const KTask& task, cudaStream_t stream, int kcap);
