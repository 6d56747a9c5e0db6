"""The communicating multi-GPU paths (sequence-parallel chunk-global attention, hybrid chunk schedule) under pytest:
2 ranks under torchrun over NCCL, skipped on a box with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(what, tmp_path, nproc=2):
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    out = tmp_path / f"{what}.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "mgpu_worker.py"), what, str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return json.load(open(out))


def test_sequence_parallel_forward_matches_single_gpu(tmp_path):
    res = _run("sp", tmp_path)["sp"]
    assert res["even_blocking"]["bit_identical"]
    assert max(res["uneven_overlap"]["max_rel_l2"].values()) < 2e-2


def test_hybrid_schedule_chunk_files_match_single_gpu(tmp_path):
    res = _run("hybrid", tmp_path)["hybrid"]
    assert res["sp_chunks"], "the plan should contain a sequence-parallel wave"
    assert res["files"] == 5
