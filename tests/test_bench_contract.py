"""bench.py contract on a CPU-only box: the reference arm prints one JSON line with the agreed
keys (the unmodified reference from baseline/_ref timed on the host cores, or the oracle port when
that install is absent); the product arm refuses to run without a GPU."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "APTS_D outer iters/sec" and d["value"] > 0
    assert d["config"]["workload"].startswith("BASELINE configs[1]")
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "dd4ml"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["vs_baseline"] is None and d["scaling"] == "weak" and d["higher_is_better"] is True


def test_product_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU path" in (r.stderr + r.stdout)


def test_reference_arm_runs_two_gloo_subdomain_processes():
    """N = 2: rank 0 of the launch spawns one reference process per subdomain (gloo); other ranks exit 0."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "dd4ml")):
        return
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""          # a non-zero rank does no work
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--config", "pinn_apts_p", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                       timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "reference"
    assert "2 gloo process" in d["cpu_baseline"]["sample"] and d["value"] > 0


def test_reference_arm_under_the_real_torchrun_launcher():
    """The driver's launch for N > 1: `python -m torch.distributed.run --nproc-per-node 2 bench.py --impl
    reference`.  torchrun exports TORCHELASTIC_USE_AGENT_STORE, under which a child process group would never
    start its own TCPStore -- the arm must scrub the launcher's environment for its gloo subdomain processes."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "dd4ml")):
        return
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--config", "pinn_apts_p", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1                                        # rank 0 alone prints
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
