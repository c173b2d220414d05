"""MODFLOW 6 input decks run end to end on the device (reader -> mf6gpu_solution_* -> .hds / .cbc):
the reference's known answers and the oracle run of the same deck.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

from modflow6_b200 import ctypes_types as T
from modflow6_b200 import simulate
from modflow6_b200.output import read_budget_file, read_head_file
from tests import mf6_inputs
from tests.test_mf6io_cpu import oracle_class

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 1, 5), (1, 5, 5), (5, 5, 5)])
@pytest.mark.parametrize("ordering", [T.ORDER_NATURAL, T.ORDER_BLOCK_MULTICOLOR])
def test_par_gwf01_on_device(gpu, tmp_path, shape, ordering):
    """autotest/test_par_gwf01.py:200-212: heads 1..10 across two models coupled by a GWF-GWF exchange"""
    mf6_inputs.write_par_gwf01(str(tmp_path), shape)
    out = simulate.run(str(tmp_path), ordering=ordering)
    assert all(r["converged"] for r in out["reports"])
    for h, first in zip(out["heads"], (1.0, 6.0)):
        np.testing.assert_array_almost_equal(h, np.broadcast_to(first + np.arange(5.0), shape))


def test_transient_deck_device_vs_oracle(gpu, tmp_path):
    rng = np.random.default_rng(7)
    shape = (3, 12, 15)
    k = np.exp(rng.normal(1.0, 0.8, shape))
    chd = [((kk + 1, i + 1, 1), 12.0) for kk in range(3) for i in range(12)] + \
          [((kk + 1, i + 1, 15), 8.0) for kk in range(3) for i in range(12)]
    sto = dict(iconvert=0, ss=1e-4, sy=0.1, periods={1: "STEADY-STATE", 2: "TRANSIENT"})
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-6\n  OUTER_MAXIMUM 50\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 200\n  INNER_DVCLOSE 1e-7\n  INNER_RCLOSE 1e-4\n  LINEAR_ACCELERATION CG\nEND linear\n")
    for tag in ("gpu", "cpu"):
        d = tmp_path / tag
        d.mkdir()
        mf6_inputs.write_gwf(str(d), "m", shape, 50.0, 40.0, 0.0, [-10.0, -25.0, -45.0], k, chd={1: chd},
                             wel={2: [((2, 6, 8), -150.0)]}, sto=sto, strt=10.0, k33=0.5)
        mf6_inputs.write_sim(str(d), ["m"], [(1.0, 1, 1.0), (30.0, 4, 1.3)], ims)
    g = simulate.run(str(tmp_path / "gpu"), ordering=T.ORDER_NATURAL)
    c = simulate.run(str(tmp_path / "cpu"), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert len(g["reports"]) == len(c["reports"]) == 5
    for a, b in zip(g["reports"], c["reports"]):
        assert a["converged"] == b["converged"] == 1 and a["outer_iterations"] == b["outer_iterations"]
        assert abs(a["pdiffr"] - b["pdiffr"]) <= 1e-3
    assert np.abs(g["heads"][0] - c["heads"][0]).max() <= 0.1 * 1e-6
    hg, hc = read_head_file(tmp_path / "gpu" / "m.hds"), read_head_file(tmp_path / "cpu" / "m.hds")
    assert len(hg) == len(hc) == 15
    bg, bc = read_budget_file(tmp_path / "gpu" / "m.cbc"), read_budget_file(tmp_path / "cpu" / "m.cbc")
    assert [(r["text"], r["kper"], r["kstp"]) for r in bg] == [(r["text"], r["kper"], r["kstp"]) for r in bc]
    for a, b in zip(bg, bc):
        va, vb = (a["flow"], b["flow"]) if a["imeth"] == 1 else (a["q"], b["q"])
        assert np.allclose(va, vb, rtol=1e-4, atol=1e-5 * max(1.0, np.abs(vb).max()))


@pytest.mark.parametrize("ordering", [T.ORDER_NATURAL, T.ORDER_BLOCK_MULTICOLOR])
def test_chd02_known_answer_on_device(gpu, tmp_path, ordering):
    """autotest/test_gwf_chd02.py:72-87: literal heads of a 10-cell unconfined Picard solve"""
    from tests.test_mf6io_cpu import CHD02_HEADS, write_chd02
    write_chd02(str(tmp_path))
    out = simulate.run(str(tmp_path), ordering=ordering)
    assert np.allclose(CHD02_HEADS, out["heads"][0].ravel())


@pytest.mark.parametrize("irch", [None, [1, 1, 1, 1, 1], [2, 2, 1, 2, 2]])
def test_rch01_on_device(gpu, tmp_path, irch):
    """autotest/test_gwf_rch01.py:124-130 on the device: the top layer dries up in the first formulate
    (npf wet/dry conversion), recharge is handed down to the highest active cell, the budget file lists the
    cells the recharge acted on"""
    from tests.test_mf6io_cpu import write_rch01
    for tag in ("gpu", "cpu"):
        (tmp_path / tag).mkdir()
        write_rch01(str(tmp_path / tag), irch)
    g = simulate.run(str(tmp_path / "gpu"), ordering=T.ORDER_NATURAL)
    c = simulate.run(str(tmp_path / "cpu"), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert g["reports"][0]["converged"] == 1
    rec = [r for r in read_budget_file(tmp_path / "gpu" / "rch.cbc") if r["text"].strip() == "RCH"][0]
    assert rec["node"].tolist() == [6, 7, 3, 9, 10] and rec["node2"].tolist() == [1, 2, 3, 4, 5]
    assert np.allclose(rec["q"], [0.0, 0.1, 0.1, 0.1, 0.0])
    hg, hc = g["heads"][0].ravel(), c["heads"][0].ravel()
    assert (hg[[0, 1, 3, 4]] == -1.0e30).all() and np.abs(hg - hc).max() <= 1e-8
    assert g["reports"][0]["outer_iterations"] == c["reports"][0]["outer_iterations"]


def test_pertim_zero_length_period_on_device(gpu, tmp_path):
    """autotest/test_gwf_pertim.py:99-115 on the device: a steady period of length zero, two CHD packages with
    BOUNDNAMES; literal canal inflow 99928.4941 and river outflow 99928.5036"""
    from tests.test_mf6io_cpu import write_pertim
    write_pertim(str(tmp_path))
    out = simulate.run(str(tmp_path), ordering=T.ORDER_BLOCK_MULTICOLOR)
    assert out["reports"][0]["converged"] == 1
    canal, river = [r for r in read_budget_file(tmp_path / "gwf_pertim.cbc") if r["text"].strip() == "CHD"]
    assert np.allclose([canal["q"][canal["q"] > 0].sum()], [99928.4941])
    assert np.allclose([-river["q"][river["q"] < 0].sum()], [99928.5036])


def test_auxmult_time_series_on_device(gpu, tmp_path):
    """autotest/test_gwf_utl04_auxmult.py:163-182 on the device: stress lists re-evaluated from their time series every
    time step (set_packages per step), well rates 1, 0, 1, 0 ... 1"""
    from tests.test_mf6io_cpu import write_auxmult
    write_auxmult(str(tmp_path), 0)
    simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL)
    q = np.array([r["q"][0] for r in read_budget_file(tmp_path / "m.cbc") if r["text"].strip() == "WEL"])
    assert np.allclose(q, np.array(7 * [1.0, 0.0])[:-1])


@pytest.mark.parametrize("ordering", [T.ORDER_NATURAL, T.ORDER_BLOCK_MULTICOLOR])
def test_lgr_exchange_with_ghost_nodes_on_device(gpu, tmp_path, ordering):
    """parent + refined child model + GWF-GWF exchange + GNC6 from input files on the device: the linear head
    field to 1e-8"""
    from tests.test_mf6io_cpu import write_lgr_gnc
    hp, hc = write_lgr_gnc(str(tmp_path), True)
    out = simulate.run(str(tmp_path), ordering=ordering)
    assert out["reports"][0]["converged"] == 1
    assert np.abs(out["heads"][0].reshape(4, 3) - hp).max() < 1e-8
    assert np.abs(out["heads"][1].reshape(8, 4) - hc).max() < 1e-8


def test_rch03_on_device(gpu, tmp_path):
    """autotest/test_gwf_rch03.py:130-146 on the device: the literal RCH budget records of array-based recharge with
    IRCH over removed / pass-through / constant-head cells (reduced numbering, bound numbers kept)"""
    from tests.test_mf6io_cpu import _write_rch0203
    idom = np.array([[[0, 0, 0, 0, 0], [0, -1, 1, -1, 0], [0, -1, 1, -1, 0], [0, 0, 0, 0, 0]],
                     [[1, 1, 1, 1, 1], [1, 1, 1, -1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]]])
    irch = np.array([[1, 0, 0, 0, 0], [0, 1, 0, 1, 0], [0, 1, 0, 1, 0], [0, 0, 0, 0, 0]]) + 1
    _write_rch0203(str(tmp_path), idom, irch)
    g = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL)
    assert g["reports"][0]["converged"] == 1
    rec = [r for r in read_budget_file(tmp_path / "rch.cbc") if r["text"].strip() == "RCH"][0]
    assert rec["node"].tolist() == [21, 27, 8, 32, 13, 34] and rec["node2"].tolist() == [1, 7, 8, 12, 13, 14]
    assert np.allclose(rec["q"], [0.0, 7.0, 8.0, 12.0, 13.0, 14.0])


@pytest.mark.parametrize("idx", [6, 8])
def test_thickstrt_hfb_deck_on_device(gpu, tmp_path, idx):
    """autotest/test_gwf_npf_thickstrt.py cases 7 and 9 from input FILES on the device (NPF THICKSTRT, HFB6)"""
    from tests.helpers import NPF_THICKSTRT, npf_thickstrt_case
    _, hfb, heads, inflow = npf_thickstrt_case(idx)
    d = str(tmp_path)
    extra = [("HFB6", "hfb", "BEGIN dimensions\n  MAXHFB 1\nEND dimensions\n\nBEGIN period 1\n"
              "  1 1 3  1 1 4  1.0e-4\nEND period 1\n")]
    mf6_inputs.write_gwf(d, "flow", (1, 1, 6), 1.0, 1.0, 10.0, [0.0], 1.0, icelltype=NPF_THICKSTRT["icelltype"][idx],
                         chd={1: [((1, 1, 1), 6.0), ((1, 1, 6), 4.0)]}, strt=5.0, k33=1.0, extra_packages=extra)
    if NPF_THICKSTRT["thickstrt"][idx]:
        p = tmp_path / "flow.npf"
        p.write_text(p.read_text().replace("  SAVE_FLOWS\n", "  SAVE_FLOWS\n  THICKSTRT\n"))
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-6\n  OUTER_MAXIMUM 10\n  UNDER_RELAXATION NONE\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 5\n  INNER_DVCLOSE 1e-6\n  INNER_RCLOSE 1e-6\n  LINEAR_ACCELERATION CG\n"
           "  SCALING_METHOD NONE\n  REORDERING_METHOD NONE\n  RELAXATION_FACTOR 1.0\nEND linear\n")
    mf6_inputs.write_sim(d, ["flow"], [(1.0, 1, 1.0)], ims)
    out = simulate.run(d, ordering=T.ORDER_NATURAL)
    assert np.allclose(heads, out["heads"][0].ravel())
    cbc = read_budget_file(tmp_path / "flow.cbc")
    assert cbc[1]["text"].strip() == "CHD" and np.allclose(inflow, cbc[1]["q"][0])


def test_ex_gwf_bump_on_device(gpu):
    """the head file MODFLOW 6 itself wrote for autotest/test_gwf_newton_under_relaxation.py (tests/golden):
    Newton-Raphson + Newton under-relaxation + BiCGSTAB on the device against the reference's own output,
    with the reference test's criterion np.allclose(base_heads, heads)"""
    import os
    from modflow6_b200.solution import GpuNumericalSolution
    from tests.test_oracle_known_answers import BUMP, bump_case
    base = read_head_file(os.path.join(BUMP, "results.hds.cmp"))[0]["data"]
    m, chd, sln, ims = bump_case()
    G = GpuNumericalSolution(m, sln, ims)
    G.set_packages([chd])
    rep = G.timestep(1, 1, 1.0, 1)
    assert rep.converged == 1
    assert np.allclose(base, G.x.reshape(51, 51))
