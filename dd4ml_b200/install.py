"""Make DD4ML's own driver pick up the B200 optimizers.

``dd4ml.utility.trainer_setup.get_config_model_and_trainer`` selects the TR-family optimizers
with late imports such as ``from dd4ml.optimizers.apts_d import APTS_D``
(trainer_setup.py:231,274,298,321,345) and ``APTS_Base.setup_APTS_hparams`` maps the
``glob_opt`` / ``loc_opt`` strings to classes (apts_base.py:46-55).  ``install()`` places
the dd4ml_b200 modules under those module paths in ``sys.modules`` so every one of these
imports resolves to the B200 classes; nothing in DD4ML's source tree is modified.
"""
from __future__ import annotations

import importlib
import sys

_MODULES = {
    "dd4ml.optimizers.tr": "dd4ml_b200.optimizers.tr",
    "dd4ml.optimizers.lsr1": "dd4ml_b200.optimizers.lsr1",
    "dd4ml.optimizers.lssr1_tr": "dd4ml_b200.optimizers.lssr1_tr",
    "dd4ml.optimizers.tradam": "dd4ml_b200.optimizers.tradam",
    "dd4ml.optimizers.apts_base": "dd4ml_b200.optimizers.apts_base",
    "dd4ml.optimizers.apts_d": "dd4ml_b200.optimizers.apts_d",
    "dd4ml.optimizers.apts_p": "dd4ml_b200.optimizers.apts_p",
    "dd4ml.optimizers.apts_pinn": "dd4ml_b200.optimizers.apts_pinn",
    "dd4ml.solvers.obs": "dd4ml_b200.solvers.obs",
}


def install(register_factory_keys: bool = True) -> dict[str, str]:
    """Route ``dd4ml.optimizers.{tr,lsr1,lssr1_tr,apts_base,apts_d,apts_p,apts_pinn}`` and
    ``dd4ml.solvers.obs`` to dd4ml_b200.  Call before ``dd4ml.utility.trainer_setup`` builds
    the optimizer.  Returns the mapping that was installed."""
    done = {}
    for ref_name, ours in _MODULES.items():
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        parent_name, _, leaf = ref_name.rpartition(".")
        parent = sys.modules.get(parent_name)
        if parent is not None:
            setattr(parent, leaf, mod)
        done[ref_name] = ours
    if register_factory_keys and "dd4ml.utility.factory" in sys.modules:
        fac = sys.modules["dd4ml.utility.factory"]
        reg = getattr(fac, "optimizer_factory", None)
        if reg is not None and hasattr(reg, "register"):
            for key, path, cls in (("tr_b200", "dd4ml_b200.optimizers.tr", "TR"),
                                   ("lssr1_tr_b200", "dd4ml_b200.optimizers.lssr1_tr", "LSSR1_TR")):
                try:
                    reg.register(key, path, cls)
                except Exception:  # noqa: BLE001  (duplicate key on a second install())
                    pass
    return done
