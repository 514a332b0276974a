"""CPU: the C-ABI library loads and exports every symbol include/dd4ml_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dd4ml_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dd4ml_[a-zA-Z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from dd4ml_b200 import _lib
    lib = _lib.get()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib.cdll, n), f"{n} declared in include/dd4ml_b200.h but not exported"
    assert set(_lib.EXPORTS) <= set(names)
    assert lib.cdll.dd4ml_version() >= 100


def test_state_layout_is_consistent():
    from dd4ml_b200 import _lib
    lay = _lib.get().layout
    assert lay.size == _lib.get().state_doubles
    assert lay.report_end < lay.size
    offs = sorted((o, c) for o, c in lay.fields.values())
    for (o, c), (o2, _) in zip(offs, offs[1:]):
        assert o + c == o2
    assert lay.maxs == lay.maxm + 1


def test_no_cpu_path():
    import pytest
    import torch
    import dd4ml_b200
    from dd4ml_b200.utility.optimizer_utils import get_tr_hparams
    from dd4ml_b200.workloads import Cfg
    m = torch.nn.Linear(3, 2)
    hp = get_tr_hparams(Cfg(delta=0.1, min_delta=1e-3, max_delta=2.0, norm_type=2, glob_second_order=False,
                            glob_dogleg=False, tol=1e-6))
    with pytest.raises(dd4ml_b200.DD4MLError):
        dd4ml_b200.TR(m.parameters(), **hp)
    with pytest.raises(dd4ml_b200.DD4MLError):
        dd4ml_b200.LSR1(device="cpu")


def test_hparam_tables_match_reference_values():
    """optimizer_utils.py:27-151 constants."""
    import math
    from dd4ml_b200.utility.optimizer_utils import (get_apts_hparams, get_loc_tr_hparams, get_lssr1_loc_tr_hparams,
                                                    get_lssr1_tr_hparams, get_tr_hparams)
    from dd4ml_b200.workloads import apts_d_yaml_config
    c = apts_d_yaml_config()
    a = get_apts_hparams(c)
    assert (a["nu_dec"], a["nu_inc"], a["inc_factor"], a["dec_factor"]) == (0.25, 0.75, 1.2, 0.9)
    t = get_tr_hparams(c)
    assert t["mem_length"] == 5 and t["nu_dec"] == 0.25 and t["inc_factor"] == 1.2 and t["second_order"] is True
    lt = get_loc_tr_hparams(c)
    assert (lt["nu_dec"], lt["nu_inc"], lt["inc_factor"], lt["dec_factor"]) == (0.3, 0.7, 1.5, 0.6)
    g = get_lssr1_tr_hparams(c)
    assert g["gamma"] == 1e-3 and g["alpha_max"] == 2.0 and g["sync"] is True and g["nu_3"] == 0.9 and g["c_2"] == 0.9
    l = get_lssr1_loc_tr_hparams(c)
    assert l["sync"] is False and l["delta"] == c.delta
    c.norm_type = math.inf
    assert get_loc_tr_hparams(c)["delta"] == c.delta
    from dd4ml_b200.utility.optimizer_utils import get_loc_tradam_hparams
    c.learning_rate = 0.02
    assert get_loc_tradam_hparams(c) == {"lr": 0.02, "betas": (0.9, 0.999), "eps": 1e-8, "norm_type": math.inf}


def test_install_overlays_reference_module_paths(monkeypatch):
    """dd4ml_b200.install(): DD4ML's late imports (trainer_setup.py:231-345) resolve to the B200 classes."""
    import sys
    import types
    import dd4ml_b200
    for name in ("dd4ml", "dd4ml.optimizers", "dd4ml.solvers"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []
            monkeypatch.setitem(sys.modules, name, mod)
    done = dd4ml_b200.install()
    try:
        from dd4ml.optimizers.apts_d import APTS_D
        from dd4ml.optimizers.apts_p import APTS_P
        from dd4ml.optimizers.apts_pinn import APTS_PINN
        from dd4ml.optimizers.lssr1_tr import LSSR1_TR
        from dd4ml.optimizers.tr import TR
        from dd4ml.optimizers.lsr1 import LSR1
        from dd4ml.solvers.obs import OBS
        assert APTS_D is dd4ml_b200.APTS_D and APTS_P is dd4ml_b200.APTS_P and APTS_PINN is dd4ml_b200.APTS_PINN
        assert TR is dd4ml_b200.TR and LSSR1_TR is dd4ml_b200.LSSR1_TR and LSR1 is dd4ml_b200.LSR1 and OBS is dd4ml_b200.OBS
        # class names keep the substrings the reference Trainer dispatches on (trainer.py:163-168,924-927)
        assert "apts_d" in APTS_D.__name__.lower() and "apts_pinn" in APTS_PINN.__name__.lower()
        from dd4ml_b200.workloads import apts_d_yaml_config
        cfg = APTS_D.setup_APTS_hparams(apts_d_yaml_config())
        assert cfg.glob_opt is LSSR1_TR and cfg.loc_opt is LSSR1_TR and cfg.glob_opt_hparams["sync"] is True
    finally:
        for k in done:
            sys.modules.pop(k, None)
