"""CPU: the C-ABI library loads and exports every symbol include/dd4ml_b200.h declares."""
import ctypes
import os
import re

import pytest

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


def _reference_src():
    for root in ("/root/reference/src", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(root, "dd4ml")):
            return root
    return None


@pytest.mark.skipif(_reference_src() is None, reason="reference package not available on this box")
def test_hparam_tables_equal_the_reference_functions():
    """Every hyper-parameter getter (optimizer_utils.py:27-164) and `setup_APTS_hparams` (apts_base.py:40-83)
    diffed against the reference's own functions over a grid of configs, in a subprocess (the reference's
    modules and the install() overlay must not leak into this interpreter)."""
    import subprocess
    import sys
    code = r'''
import itertools, math, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import dd4ml.utility as RU
from dd4ml.utility import CfgNode
from dd4ml.optimizers.apts_base import APTS_Base as RefBase
import dd4ml_b200.utility.optimizer_utils as OU
from dd4ml_b200.optimizers.apts_base import APTS_Base as OurBase
import torch.distributed as dist
names = ["get_apts_hparams", "get_lssr1_tr_hparams", "get_lssr1_loc_tr_hparams", "get_tr_hparams",
         "get_loc_tr_hparams", "get_loc_tradam_hparams"]
n = 0
for ws in (1, 2):
    if ws == 2:
        import os
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", "29921"
        # the local tables scale radii by 1 / world_size (optimizer_utils.py:11-15): fake a 2-rank world
        import unittest.mock as um
        p1 = um.patch("torch.distributed.is_initialized", return_value=True); p1.start()
        p2 = um.patch("torch.distributed.get_world_size", return_value=2); p2.start()
    for norm, so, dog, m, paper in itertools.product((2, math.inf), (False, True), (False, True), (3, 5, 10), (False, True)):
        if dog and not so:
            continue
        kw = dict(delta=0.07, min_delta=0.003, max_delta=1.7, norm_type=norm, glob_second_order=so, glob_dogleg=dog,
                  loc_second_order=so, loc_dogleg=dog, mem_length=m, max_wolfe_iters=4, max_zoom_iters=6, tol=1e-7,
                  paper_tr_update=paper, learning_rate=0.02, glob_opt="lssr1_tr", loc_opt="tr")
        for name in names:
            a, b = getattr(RU, name)(CfgNode(**kw)), getattr(OU, name)(CfgNode(**kw))
            assert a == b, (name, kw, a, b)
            n += 1
        for glob, loc in (("tr", "tr"), ("lssr1_tr", "lssr1_tr"), ("tr", "tradam"), ("lssr1_tr", "sgd")):
            ra = RefBase.setup_APTS_hparams(CfgNode(**dict(kw, glob_opt=glob, loc_opt=loc)))
            oa = OurBase.setup_APTS_hparams(CfgNode(**dict(kw, glob_opt=glob, loc_opt=loc)))
            assert ra.glob_opt.__name__ == oa.glob_opt.__name__ and ra.loc_opt.__name__ == oa.loc_opt.__name__
            assert ra.glob_opt_hparams == oa.glob_opt_hparams and ra.loc_opt_hparams == oa.loc_opt_hparams
            assert ra.apts_params == oa.apts_params
            n += 1
print("compared", n)
''' % (_reference_src(), ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "compared" in r.stdout and int(r.stdout.split()[-1]) > 300
