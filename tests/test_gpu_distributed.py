"""Split-model path on >= 2 GPUs (NCCL halo exchange + all-gathered reductions), checked against the oracle
run on the UNSPLIT model with the same block-Jacobi ILU0.  Skipped on a single-GPU box (the logic that does
not need GPUs is covered by tests/test_distributed_cpu.py with gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from modflow6_b200 import lib
    return lib.load().mf6gpu_device_count()


@pytest.mark.parametrize("cfg", ["3 24 30 1 2 0 1", "3 24 30 2 1 0 1", "3 24 30 1 2 1 1", "3 24 30 1 2 0 2"])
def test_two_rank_parity(cfg):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "dist_check.py")] + cfg.split()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "DIST_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_four_rank_2x2(gpu):
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "scripts", "dist_check.py"),
           "3", "24", "30", "2", "2", "0", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "DIST_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("shape,ordering", [("5 5 5", 0), ("4 6 5", 2)])
def test_two_rank_input_deck(tmp_path, shape, ordering):
    """`mf6 -p` analogue: the two-model par_gwf01 deck (autotest/test_par_gwf01.py), one model per GPU coupled
    through its GWF-GWF exchange; heads 1..10 (known answer) and equal to the unsplit single-GPU run"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    from modflow6_b200.output import read_head_file
    from tests import mf6_inputs
    shp = tuple(int(v) for v in shape.split())
    mf6_inputs.write_par_gwf01(str(tmp_path), shp)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29535", os.path.join(ROOT, "scripts", "dist_deck.py"), str(tmp_path),
           str(ordering), "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "DIST_DECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    for name, first in (("leftmodel", 1.0), ("rightmodel", 6.0)):
        recs = read_head_file(tmp_path / f"{name}.hds")
        assert len(recs) == shp[0]
        for rec in recs:
            np.testing.assert_array_almost_equal(rec["data"], np.broadcast_to(first + np.arange(5.0), shp[1:]))


@pytest.mark.parametrize("case,ordering", [("disv", 0), ("disv", 2), ("wetdry", 0), ("wetdry", 2)])
def test_two_rank_generic_models(case, ordering):
    """split-model path beyond DIS blocks: a hexagonal DISV model cut into stripes, and an unconfined model whose
    top layer dries up (wet/dry conversion + recharge hand-down with per-outer ibound exchange); both against the
    unsplit model solved on one GPU"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29536", os.path.join(ROOT, "scripts", "dist_generic_check.py"), case,
           str(ordering)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "DIST_GENERIC PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
