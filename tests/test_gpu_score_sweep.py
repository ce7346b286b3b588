"""GPU: BASELINE config 5 pipeline (tools/score_sweep.py) at a small size -- ResNetSE embeddings -> sharded cosine scoring -> parity
with the fp64 oracle pipeline (< 1e-4, asserted inside the tool).  Runs world-size 2 over NCCL when two GPUs are visible."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--trials", "24", "--enroll", "20", "--table", "64", "--pairs", "5000", "--chunk", "8", "--iters", "3", "--oracle-rows", "3"]


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def test_score_sweep_single_gpu(cuda):
    d = run([sys.executable, "tools/score_sweep.py"] + ARGS)
    assert d["n_gpus"] == 1 and d["all_pairs"]["pairs"] == 480
    assert d["max_abs_err_vs_oracle_pipeline"] < 1e-4
    assert d["all_pairs"]["max_abs_err_vs_fp64_cosine"] < 1e-4 and d["pair_list"]["max_abs_err_vs_fp64_cosine"] < 1e-4


def test_score_sweep_two_gpus(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible: the world-size-2 NCCL path needs two (tests/test_dist_cpu.py covers the host logic on gloo)")
    d = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
             "--master-port", "29533", "tools/score_sweep.py"] + ARGS)
    assert d["n_gpus"] == 2
    assert d["max_abs_err_vs_oracle_pipeline"] < 1e-4
