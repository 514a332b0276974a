"""CPU oracle for the DD4ML trust-region hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain torch-CPU restatement of the reference algorithms
(cruzas/DD4ML ``src/dd4ml/optimizers/{tr,lsr1,lssr1_tr,apts_base,apts_d,apts_p,
apts_pinn}.py``, ``solvers/obs.py``, ``utility/optimizer_utils.py``) written over
*flat vectors and function closures* ``fg(w) -> (loss, grad)`` instead of
``nn.Module`` parameter lists.  Every function cites the reference file:line it
follows.

It may be imported only from ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and only as the
checker / reported CPU baseline.  The product package ``dd4ml_b200`` never
imports it and has no CPU fallback.

Parity pinning: the reference ships no golden vectors for this path
(SURVEY.md §8c).  The oracle is pinned against outputs of the *unmodified
reference itself*, generated in the authoring container by
``tests/golden/gen_golden.py`` and committed under ``tests/golden/*.npz``
(checked by ``tests/test_oracle_golden.py``).

One reference behaviour is LAPACK-implementation-defined: OBS uses
``max(a_j/delta - Lambda_j)`` with *signed* ``a_j`` whose signs are the signs of
the eigenvectors returned by ``torch.linalg.eigh`` (MKL ``ssyevd``).  The oracle
therefore has two modes, ``eig_sign="lapack"`` (literal; what the goldens pin)
and ``eig_sign="canonical"`` (largest-magnitude component of every eigenvector
made positive; the convention the CUDA path implements).  See DESIGN.md §5.
"""

from .lsr1 import LSR1Memory  # noqa: F401
from .obs import obs_solve  # noqa: F401
from .subproblem import solve_first_order, solve_second_order  # noqa: F401
from .tr import TROracle  # noqa: F401
from .lssr1_tr import LSSR1TROracle  # noqa: F401
from .apts import APTSDOracle, APTSPOracle, apts_hparams, tr_hparams, loc_tr_hparams, lssr1_hparams, lssr1_loc_hparams  # noqa: F401
from .tradam import TRAdamOracle  # noqa: F401
