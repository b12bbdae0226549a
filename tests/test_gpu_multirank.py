"""Multi-rank GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the sharded path --
whole grid points per rank, the leftover point evaluated by both ranks together (hybrid
row-slab partition), NCCL exchanges -- against the single-rank run of the same calculation."""
import json
import os
import subprocess
import sys

import numpy
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multirank_worker.py")


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(nproc, tag, tmp, extra_env=None):
    out = os.path.join(str(tmp), "%s.json" % tag)
    env = dict(os.environ)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + (os.getpid() + nproc) % 300),
           WORKER, out]
    if nproc == 1:
        cmd = [sys.executable, WORKER, out]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return json.load(open(out))


@pytest.mark.parametrize("hybrid", ["1", "0"])
def test_two_ranks_match_one_rank(built, tmp_path, hybrid):
    """3 iterations of the amplitude loop from the MP2 guess, a Lambda solve, RDMs and E/S/N on
    2 ranks = the 1-rank trajectory to 1e-12, with the leftover grid point evaluated by both
    ranks together (hybrid) or by one owner; the two ranks hold bit-identical results."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    one = _run(1, "one", tmp_path)
    two = _run(2, "two" + hybrid, tmp_path, {"KB200_HYBRID": hybrid})
    assert two["world"] == 2 and two["sharded"] and two["hybrid_used"] == (hybrid == "1")
    for k in ("traj_E", "traj_res"):
        a, b = numpy.array(one[k]), numpy.array(two[k])
        assert a.shape == b.shape
        assert numpy.abs(a - b).max() <= 1e-12*max(1.0, numpy.abs(a).max()), k
    for k in ("omega", "E", "S", "N", "lam_norm", "t2_checksum", "n1rdm_trace"):
        assert abs(one[k] - two[k]) <= 1e-12*max(1.0, abs(one[k])), (k, one[k], two[k])
    assert two["ranks_agree"]


def test_owner_rows_as_one_batch(built, tmp_path):
    """ngrid 4 on 2 ranks without the shared evaluation: 3 evaluated points = 1 per rank + 1
    leftover, which rank 0 runs together with its own point as one strided batch of 2."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    env = {"KB200_TEST_NG": "4", "KB200_HYBRID": "0"}
    one = _run(1, "one4", tmp_path, env)
    two = _run(2, "two4", tmp_path, env)
    assert two["world"] == 2 and two["sharded"] and not two["hybrid_used"]
    for k in ("traj_E", "traj_res"):
        a, b = numpy.array(one[k]), numpy.array(two[k])
        assert numpy.abs(a - b).max() <= 1e-12*max(1.0, numpy.abs(a).max()), k
    for k in ("omega", "E", "S", "N", "lam_norm", "t2_checksum", "n1rdm_trace"):
        assert abs(one[k] - two[k]) <= 1e-12*max(1.0, abs(one[k])), (k, one[k], two[k])
    assert two["ranks_agree"]
