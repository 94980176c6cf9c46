"""Multi-rank parity of the distributed path (kman_b200/dist.py) under `pytest -m gpu`: rank-order
concatenation of the per-rank count tables / singleton lists must equal the oracle's global result
(kmermaid/seq.py:361-383 chunking with k-1 overlap, join.py:95-130 grouping), on multi-record inputs
with separators, duplicated stretches, N runs, IUPAC symbols and soft-masked bases.

Two topologies, both driven through tools/dist_gpu_check.py under torchrun:
  * two ranks sharing ONE GPU (gloo plumbing, CUDA-IPC peer buffers): runs on the single-GPU test box;
  * one rank per GPU over NCCL when the box has at least two GPUs.
Both exchange variants (single launch with shared cursors; exact per-source regions) and the plain
range-partition + all-to-all path are exercised."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, extra_env, port):
    env = dict(os.environ)
    env.update(extra_env)
    env["MASTER_ADDR"] = "127.0.0.1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_gpu_check.py")]
    p = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = (p.stdout + "\n" + p.stderr)[-4000:]
    assert p.returncode == 0 and "DIST_GPU_OK" in p.stdout, tail
    assert "MISMATCH" not in p.stdout, tail
    return p.stdout


@pytest.mark.parametrize("variant", ["shared", "exact", "alltoall"])
def test_two_ranks_on_one_gpu(variant):
    env = {"KMG_TEST_ONE_GPU": "1", "KMG_TEST_CASES": "0,1,2,4,5,7"}
    if variant == "shared":
        env["KMG_DIST_SHARED"] = "1"
    elif variant == "exact":
        env["KMG_DIST_SHARED"] = "0"
    else:
        env["KMG_DIST_P2P"] = "0"
    out = _run(2, env, 29611 + ["shared", "exact", "alltoall"].index(variant))
    assert ("p2p path: False" in out) == (variant == "alltoall"), out[-2000:]


def test_three_ranks_on_one_gpu_not_a_power_of_two():
    _run(3, {"KMG_TEST_ONE_GPU": "1", "KMG_TEST_CASES": "0,4", "KMG_DIST_SHARED": "0"}, 29621)


def test_one_rank_per_gpu_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    _run(2 if n < 4 else 4, {}, 29631)
