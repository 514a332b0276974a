"""The reference's OWN driver path with the optimizers swapped by `dd4ml_b200.install()`:

    dd4ml.utility.trainer_setup.get_config_model_and_trainer(args, None)   (trainer_setup.py:37-432: config merge,
        dataset / model / criterion factories, late `from dd4ml.optimizers.X import Y`, setup_APTS_hparams)
    -> dd4ml.trainer.Trainer.run()                                         (trainer.py:606-613 -> run_by_epoch_PINN
        :1198-1290: model.double(), optimizer.loc_model.double(), _train_one_batch_PINN :985-1100 with its
        signature / class-name dispatch of optimizer.step)

on BASELINE config 5's data set, model and loss (allencahn1d / pinn_ffnn / pinn_allencahn).  Two subprocesses
run tests/wiring_driver.py: the unmodified reference on the CPU, and the same script with one extra line,
`dd4ml_b200.install()`, on cuda:0.  What the Trainer reports per epoch (loss, optimizer.delta, grad_evals)
must agree.  Needs the pip-installed reference in baseline/_ref (it travels to the GPU box)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HAVE_REF = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "dd4ml"))


def _run(impl, optimizer, inner, port):
    r = subprocess.run([sys.executable, os.path.join(HERE, "wiring_driver.py"), "--impl", impl, "--optimizer", optimizer,
                        "--inner", inner, "--port", str(port)], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1])


@pytest.mark.skipif(not torch.cuda.is_available() or not HAVE_REF, reason="needs a GPU and baseline/_ref")
@pytest.mark.parametrize("optimizer,inner,rtol", [
    ("apts_p", "tr", 1e-7),            # BASELINE config 5 (TR / TR)
    ("apts_p", "lssr1_tr", 1e-5),      # config_apts_p.yaml's inner optimizers
    ("apts_pinn", "tr", 1e-7),         # the optimizer tests/test_pinn_subdomains.py drives
    ("tr", "tr", 1e-7),                # plain TR through the Trainer's generic closure
])
def test_reference_trainer_runs_the_b200_optimizers(optimizer, inner, rtol):
    port = 29860 + hash((optimizer, inner)) % 30
    ref = _run("reference", optimizer, inner, port)
    ours = _run("b200", optimizer, inner, port + 40)
    assert ref["optimizer_module"].startswith("dd4ml.optimizers.")
    assert ours["optimizer_module"].startswith("dd4ml_b200.optimizers.")       # the overlay took effect
    assert ours["optimizer_class"] == ref["optimizer_class"]
    assert ours["device"].startswith("cuda") and ours["param_dtype"] == ref["param_dtype"] == "torch.float64"
    assert ours["gpu_launches"] > 0
    a, b = np.asarray(ref["log"]), np.asarray(ours["log"])
    assert a.shape == b.shape and a.shape[0] == 5                                # warm-up pass + 4 epochs
    np.testing.assert_array_equal(a[:, 0], b[:, 0])
    np.testing.assert_allclose(b[:, 1], a[:, 1], rtol=rtol)                      # loss per epoch
    np.testing.assert_allclose(b[:, 2], a[:, 2], rtol=1e-9)                      # optimizer.delta
    np.testing.assert_allclose(b[:, 3], a[:, 3], rtol=0, atol=1e-9)              # Trainer's grad_evals counter
    assert a[-1, 1] < a[0, 1]
