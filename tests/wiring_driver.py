"""Runs the UNMODIFIED reference driver path -- dd4ml.utility.trainer_setup.get_config_model_and_trainer
-> dd4ml.trainer.Trainer.run() (trainer_setup.py:37-432, trainer.py:606-613, 1198-1290) -- on the
built-in Allen-Cahn PINN dataset / PINNFFNN model / AllenCahnPINNLoss (BASELINE config 5), either with
the reference's own optimizers on the CPU (`--impl reference`) or, after `dd4ml_b200.install()`, with
the dd4ml_b200 optimizers on cuda:0 (`--impl b200`).  Prints one JSON line with the per-epoch
(loss, delta, grad_evals) the Trainer reports.  Needs baseline/_ref (the pip-installed reference)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True, choices=["reference", "b200"])
    ap.add_argument("--optimizer", default="apts_p")
    ap.add_argument("--inner", default="tr")
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--port", type=int, default=29871)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import dd4ml.utility  # noqa: F401  (import-order workaround)
    from unittest.mock import patch
    from dd4ml.utility import set_seed
    torch.set_num_threads(1)
    device = "cpu"
    if a.impl == "b200":
        import dd4ml_b200
        dd4ml_b200.install()               # the ONLY difference between the two arms
        device = "cuda:0"
        torch.cuda.set_device(0)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(a.port)
    dist.init_process_group("gloo", rank=0, world_size=1)   # mark_trainable needs a group (apts_utils.py:102)
    from dd4ml.utility.trainer_setup import get_config_model_and_trainer
    args = {"dataset_name": "allencahn1d", "model_name": "pinn_ffnn", "optimizer": a.optimizer,
            "criterion": "pinn_allencahn", "batch_size": 1002, "effective_batch_size": 1002, "epochs": a.epochs,
            "max_iters": 0, "learning_rate": 0.1, "seed": 42, "num_subdomains": 1, "num_stages": 1,
            "num_replicas_per_subdomain": 1, "gradient_accumulation": False, "accumulation_steps": 1,
            "batch_inc_factor": 1.0, "overlap": 0.0, "glob_opt": a.inner, "loc_opt": a.inner, "delta": 0.002,
            "min_delta": 1e-5, "max_delta": 0.5, "norm_type": "inf" if a.optimizer == "apts_p" else 2,
            "max_loc_iters": 3, "max_glob_iters": 1, "glob_pass": True, "foc": True, "glob_second_order": False,
            "loc_second_order": False, "glob_dogleg": False, "loc_dogleg": False, "tol": 1e-6, "mem_length": 5,
            "max_wolfe_iters": 5, "max_zoom_iters": 5, "paper_tr_update": True, "shuffle": False}
    set_seed(42)
    log = []
    real_stdout = os.dup(1)
    os.dup2(2, 1)                          # the reference prints banners; keep stdout for the JSON line
    with patch("dd4ml.utility.trainer_setup.get_device", return_value=device), \
            patch("dd4ml.trainer.dist.get_backend", return_value="gloo" if device == "cpu" else "nccl"):
        # (Trainer.__init__ picks cuda unless the backend is gloo, trainer.py:133-141; the 1-rank group here is
        # gloo in both arms -- it only exists because mark_trainable asks for the rank)
        cfg, model, trainer = get_config_model_and_trainer(args, None)

        def on_epoch_end(t):
            log.append((int(t.epoch_num), float(t.loss), float(getattr(t.optimizer, "delta", 0.0)),
                        float(t.grad_evals)))
        trainer.set_callback("on_epoch_end", on_epoch_end)
        trainer.run()
    os.dup2(real_stdout, 1)
    opt = trainer.optimizer
    out = {"impl": a.impl, "optimizer_module": type(opt).__module__, "optimizer_class": type(opt).__name__,
           "param_dtype": str(next(model.parameters()).dtype), "device": str(next(model.parameters()).device),
           "log": log}
    if a.impl == "b200":
        from dd4ml_b200 import _lib
        out["gpu_launches"] = _lib.get().launches
    sys.stdout.write(json.dumps(out) + "\n")
    sys.stdout.flush()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
