"""Input reader (modflow6_b200/mf6io.py) + time loop (simulate.py) + binary writers, driven with the CPU
oracle as the solution class: the reference's own known answers reached from input FILES.

  * autotest/test_par_gwf01.py:200-212 -- two models joined by a GWF-GWF exchange, heads 1..10 (1-D, 2-D, 3-D)
  * autotest/test_gwf_chd01.py:126-127 -- heads == linspace(1, 0, 100)
  * a transient deck (STO periods, WEL from period 2, heterogeneous K as INTERNAL arrays, OC) must give exactly
    the heads of the same model assembled directly with grid.build_dis_model
"""
import numpy as np
import pytest

from modflow6_b200 import ctypes_types as T
from modflow6_b200 import mf6io, simulate
from modflow6_b200.grid import Package, build_dis_model, tdis_steps
from modflow6_b200.output import read_budget_file, read_head_file
from tests import mf6_inputs


def oracle_class():
    from oracle.oracle import OracleSolution
    return OracleSolution


@pytest.mark.parametrize("shape", [(1, 1, 5), (1, 5, 5), (5, 5, 5)])
def test_par_gwf01_known_answer(tmp_path, shape):
    mf6_inputs.write_par_gwf01(str(tmp_path), shape)
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert all(r["converged"] for r in out["reports"])
    left, right = out["heads"]
    for h, first in ((left, 1.0), (right, 6.0)):
        want = np.broadcast_to(first + np.arange(5.0), shape)
        np.testing.assert_array_almost_equal(h, want)          # 6 decimals, like the reference test
    # both head files exist and hold one record per layer
    recs = read_head_file(tmp_path / "leftmodel.hds")
    assert len(recs) == shape[0] and recs[0]["ncol"] == 5 and recs[0]["nrow"] == shape[1]
    np.testing.assert_array_almost_equal(recs[0]["data"][0], [1, 2, 3, 4, 5])


def test_chd01_known_answer(tmp_path):
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-6\n  OUTER_MAXIMUM 100\n  UNDER_RELAXATION NONE\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 300\n  INNER_DVCLOSE 1e-6\n  INNER_RCLOSE 1e-6\n  LINEAR_ACCELERATION CG\n"
           "  SCALING_METHOD NONE\n  REORDERING_METHOD NONE\n  RELAXATION_FACTOR 1.0\nEND linear\n")
    mf6_inputs.write_gwf(str(tmp_path), "gwf", (1, 1, 100), 1.0, 1.0, 1.0, [0.0], 1.0,
                         chd={1: [((1, 1, 1), 1.0), ((1, 1, 100), 0.0)]}, strt=1.0)
    mf6_inputs.write_sim(str(tmp_path), ["gwf"], [(5.0, 1, 1.0)], ims)
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert np.allclose(out["heads"][0].ravel(), np.linspace(1, 0, 100))
    sim = out["simulation"]
    assert sim.ims.relax == 1.0 and sim.ims.ilinmeth == 1 and sim.sln.mxiter == 100
    cbc = read_budget_file(tmp_path / "gwf.cbc")
    assert [r["text"].strip() for r in cbc] == ["FLOW-JA-FACE", "CHD"]
    assert cbc[1]["srcmodel"].strip() == "GWF" and cbc[1]["dstpackage"].strip() == "CHD_0"
    q = cbc[1]["q"]
    assert np.isclose(q[0], -q[1]) and np.isclose(q[0], 1.0 / 99.0)     # Darcy: K dh/dx * area


def test_transient_deck_equals_direct_model(tmp_path):
    rng = np.random.default_rng(5)
    shape = (2, 7, 9)
    k = np.exp(rng.normal(1.0, 0.7, shape))
    botm = [-10.0, -25.0]
    chd = [((kk + 1, i + 1, 1), 12.0) for kk in range(2) for i in range(7)] + \
          [((kk + 1, i + 1, 9), 8.0) for kk in range(2) for i in range(7)]
    wel = [((2, 4, 5), -150.0)]
    sto = dict(iconvert=0, ss=1e-4, sy=0.1, periods={1: "STEADY-STATE", 2: "TRANSIENT"})
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-7\n  OUTER_MAXIMUM 50\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 100\n  INNER_DVCLOSE 1e-8\n  INNER_RCLOSE 1e-4 STRICT\n"
           "  LINEAR_ACCELERATION BICGSTAB\nEND linear\n")
    d = str(tmp_path)
    mf6_inputs.write_gwf(d, "m", shape, 50.0, 40.0, 0.0, botm, k, chd={1: chd}, wel={2: wel}, sto=sto, strt=10.0,
                         k33=0.5)
    mf6_inputs.write_sim(d, ["m"], [(1.0, 1, 1.0), (30.0, 4, 1.3), (10.0, 2, 1.0)], ims)
    out = simulate.run(d, ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert len(out["reports"]) == 7 and all(r["converged"] for r in out["reports"])
    sim = out["simulation"]
    assert sim.ims.icnvgopt == 1 and sim.ims.ilinmeth == 2 and sim.models[0].sto_transient == {1: False, 2: True}
    # the same model without files
    m = build_dis_model(2, 7, 9, 50.0, 40.0, 0.0, botm, k, k33=0.5, strt=10.0, ss=1e-4, sy=0.1, iconvert=0)
    node = lambda c: ((c[0] - 1) * 7 + c[1] - 1) * 9 + c[2] - 1   # noqa: E731
    pc = Package(T.PKG_CHD, [node(c) for c, _ in chd], [v for _, v in chd])
    pw = Package(T.PKG_WEL, [node(c) for c, _ in wel], [v for _, v in wel])
    O = oracle_class()(m, sim.sln, sim.ims)
    steps = 0
    for kper, (perlen, nstp, tsm) in enumerate(sim.perioddata, start=1):
        O.set_packages([pc] if kper == 1 else [pc, pw])
        for kstp, delt in enumerate(tdis_steps(perlen, nstp, tsm), start=1):
            O.timestep(kper, kstp, delt, 1 if kper == 1 else 0)
            steps += 1
    assert np.array_equal(out["heads"][0].ravel(), O.x)
    # OC: SAVE HEAD ALL, SAVE BUDGET LAST (period-1 block stays in force)
    hds = read_head_file(tmp_path / "m.hds")
    assert len(hds) == 7 * 2 and hds[-1]["kper"] == 3 and np.isclose(hds[-1]["totim"], 41.0)
    cbc = read_budget_file(tmp_path / "m.cbc")
    assert sorted({(r["kper"], r["kstp"]) for r in cbc}) == [(1, 1), (2, 4), (3, 2)]
    assert [r["text"].strip() for r in cbc if r["kper"] == 2] == ["STO-SS", "FLOW-JA-FACE", "CHD", "WEL"]


def test_reader_rejects_what_the_path_cannot_honour(tmp_path):
    d = str(tmp_path)
    mf6_inputs.write_gwf(d, "m", (1, 2, 2), 1.0, 1.0, 0.0, [-1.0], 1.0, chd={1: [((1, 1, 1), 1.0)]})
    mf6_inputs.write_sim(d, ["m"], [(1.0, 1, 1.0)], "BEGIN linear\n  PRECONDITIONER_LEVELS 2\nEND linear\n")
    sim = mf6io.read_simulation(d)
    assert sim.ims.level == 2 and not any("ILUT" in w for w in sim.warnings)      # ILUT is honoured (IPC 3)
    npf = (tmp_path / "m.npf").read_text()
    (tmp_path / "m.npf").write_text(npf.replace("SAVE_FLOWS", "SAVE_FLOWS\n  XT3D"))
    with pytest.raises(mf6io.Mf6InputError, match="XT3D"):
        mf6io.read_simulation(d)
    (tmp_path / "m.npf").write_text(npf)
    with open(tmp_path / "m.nam", "w") as f:
        f.write("BEGIN packages\n  DIS6 m.dis\n  IC6 m.ic\n  NPF6 m.npf\n  LAK6 m.lak\nEND packages\n")
    with pytest.raises(mf6io.Mf6InputError, match="LAK6"):
        mf6io.read_simulation(d)


def test_array_control_records(tmp_path):
    p = tmp_path / "a.txt"
    p.write_text("1 2 3\n4 5 6\n")
    lines = [["k", "LAYERED"], ["CONSTANT", "2.5"], ["INTERNAL", "FACTOR", "2.0"], ["1", "2", "3"], ["4", "5", "6"],
             ["icelltype"], ["OPEN/CLOSE", "a.txt", "FACTOR", "1"]]
    g = mf6io.read_griddata(lines, str(tmp_path), {"K": ((2, 2, 3), np.float64), "ICELLTYPE": ((1, 2, 3), np.int32)})
    assert g["K"].tolist() == [2.5] * 6 + [2.0, 4.0, 6.0, 8.0, 10.0, 12.0]
    assert g["ICELLTYPE"].tolist() == [1, 2, 3, 4, 5, 6] and g["ICELLTYPE"].dtype == np.int32
    # OPEN/CLOSE (BINARY): one 52-byte header + m1 x m2 values per layer, doubles or 4-byte integers -- a head file
    # written by HeadFileWriter has exactly that layout, which is how the reference restarts from saved heads
    from modflow6_b200.output import HeadFileWriter
    hw = HeadFileWriter(str(tmp_path / "h.bin"), (2, 2, 3))
    hw.write(1, 1, 1.0, 1.0, np.arange(12.0))
    hw.close()
    import struct
    with open(tmp_path / "i.bin", "wb") as f:
        f.write(struct.pack("<iidd16siii", 1, 1, 1.0, 1.0, b"         IDOMAIN", 3, 2, 1) + np.arange(6, dtype="<i4").tobytes())
    lines = [["strt"], ["OPEN/CLOSE", "h.bin", "(BINARY)", "FACTOR", "2.0"], ["idomain"], ["OPEN/CLOSE", "i.bin", "(BINARY)"]]
    g = mf6io.read_griddata(lines, str(tmp_path), {"STRT": ((2, 2, 3), np.float64), "IDOMAIN": ((1, 2, 3), np.int32)})
    assert g["STRT"].tolist() == (2.0 * np.arange(12.0)).tolist() and g["IDOMAIN"].tolist() == [0, 1, 2, 3, 4, 5]
    assert mf6io._tokens("  SAVE  HEAD, 'my file.hds'  # trailing") == ["SAVE", "HEAD", "my file.hds"]
    assert mf6io._tokens("! comment") == [] and mf6io._tokens("// c") == []


def test_reference_minsim_deck(tmp_path):
    """the reference's own example deck (`.mf6minsim/`: two convertible 1x1x5 models, GWF-GWF exchange, IMS with
    ILUT levels): read unchanged from the reference tree when it is present (it is not on the
    GPU box).  1-D unconfined flow between CHD 1 and CHD 10 over a -100 m bottom: (h + 100)^2 falls on a
    straight line in x up to the discretisation of the saturated thickness"""
    import os
    import shutil
    src = "/root/reference/.mf6minsim"
    if not os.path.isdir(src):
        pytest.skip("reference tree not present")
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp_path)
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert out["simulation"].ims.level > 0            # the deck's ILUT levels are honoured (IPC 3), not downgraded
    assert all(r["converged"] for r in out["reports"]) and abs(out["reports"][0]["pdiffr"]) < 1e-5
    h = np.concatenate([x.ravel() for x in out["heads"]])
    assert h[0] == 1.0 and h[-1] == 10.0 and np.all(np.diff(h) > 0)
    t2 = (h + 100.0) ** 2
    fit = np.polyval(np.polyfit(np.arange(10.0), t2, 1), np.arange(10.0))
    assert np.abs(t2 - fit).max() / t2.mean() < 2e-4


def test_disv_and_disu_decks_of_a_rectangular_grid_equal_the_dis_deck(tmp_path):
    """VERTICES / CELL2D -> connections (vertexconnect, DisvGeom cprops): the same rectangular grid written as a
    DISV package must give the DIS connectivity (ia, ja, ihc, cl1, cl2, hwva, area) and therefore the same heads"""
    rng = np.random.default_rng(11)
    shape = (3, 5, 6)
    k = np.exp(rng.normal(0.5, 0.6, shape))
    delr = np.array([10.0, 20.0, 30.0, 15.0, 25.0, 10.0])
    delc = np.array([12.0, 8.0, 20.0, 16.0, 10.0])
    botm = [-5.0, -12.0, -30.0]
    chd = [((kk + 1, i + 1, 1), 5.0) for kk in range(3) for i in range(5)] + \
          [((kk + 1, i + 1, 6), 2.0) for kk in range(3) for i in range(5)]
    wel = [((3, 3, 4), -40.0)]
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-9\n  OUTER_MAXIMUM 50\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 200\n  INNER_DVCLOSE 1e-10\n  INNER_RCLOSE 1e-8\n  LINEAR_ACCELERATION CG\nEND linear\n")
    outs = {}
    for tag in ("dis", "disv", "disu"):
        d = tmp_path / tag
        d.mkdir()
        mf6_inputs.write_gwf(str(d), "m", shape, delr, delc, 0.0, botm, k, chd={1: chd}, wel={1: wel}, strt=3.0,
                             k33=0.3, disv=(tag == "disv"), disu=(tag == "disu"))
        mf6_inputs.write_sim(str(d), ["m"], [(1.0, 1, 1.0)], ims)
        outs[tag] = simulate.run(str(d), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    a, b = outs["dis"]["simulation"].models[0].model, outs["disv"]["simulation"].models[0].model
    assert np.array_equal(a.ia, b.ia) and np.array_equal(a.ja, b.ja) and np.array_equal(a.ihc, b.ihc)
    for name in ("cl1", "cl2", "hwva", "area", "top", "bot"):
        assert np.allclose(getattr(a, name), getattr(b, name), rtol=1e-13), name
    assert np.abs(outs["dis"]["heads"][0].ravel() - outs["disv"]["heads"][0].ravel()).max() < 1e-9
    recs = read_head_file(tmp_path / "disv" / "m.hds")
    assert len(recs) == 3 and recs[0]["ncol"] == 30 and recs[0]["nrow"] == 1      # DISV: ncol = ncpl, nrow = 1
    # and the DISU deck (CONNECTIONDATA: iac, ja, ihc, cl12, hwva -> disuconnections + con_finalize)
    c = outs["disu"]["simulation"].models[0].model
    assert np.array_equal(a.ia, c.ia) and np.array_equal(a.ja, c.ja) and np.array_equal(a.ihc, c.ihc)
    for name in ("cl1", "cl2", "hwva", "area", "top", "bot"):
        assert np.allclose(getattr(a, name), getattr(c, name), rtol=1e-13), name
    assert np.array_equal(a.ibotnode, c.ibotnode)
    assert np.abs(outs["dis"]["heads"][0].ravel() - outs["disu"]["heads"][0].ravel()).max() < 1e-9
    recs = read_head_file(tmp_path / "disu" / "m.hds")
    assert len(recs) == 1 and recs[0]["ncol"] == 90 and recs[0]["nrow"] == 1      # DISU: one record of all nodes


def test_multi_model_budget_files(tmp_path):
    """per-model .cbc of a two-model simulation: the model's own FLOW-JA-FACE, its packages, then the GWF-GWF
    exchange flows as a FLOW-JA-FACE list towards the other model (gwf_gwf_bdsav_model); the uniform flow field of
    par_gwf01 (gradient 1 per 100 m cell, K = 1, face area 100 x 100) carries 100 m3/d per cell face"""
    shape = (2, 3, 5)
    mf6_inputs.write_par_gwf01(str(tmp_path), shape)
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    left, right = (read_budget_file(tmp_path / f"{m}.cbc") for m in ("leftmodel", "rightmodel"))
    assert [r["text"].strip() for r in left] == ["FLOW-JA-FACE", "CHD", "FLOW-JA-FACE"]
    ml = out["simulation"].models[0].model
    assert left[0]["imeth"] == 1 and left[0]["flow"].size == ml.nja
    ex_l, ex_r = left[2], right[2]
    assert ex_l["imeth"] == 6 and ex_l["srcmodel"].strip() == "LEFTMODEL" and ex_l["dstmodel"].strip() == "RIGHTMODEL"
    assert ex_l["srcpackage"].strip() == ex_l["dstpackage"].strip() == "GWF-GWF_1"
    assert [a.strip() for a in ex_l["auxtxt"]] == ["ANGLDEGX", "CDIST"] and np.allclose(ex_l["aux"][:, 1], 100.0)
    # left boundary column 5 <-> right column 1, every layer and row
    cell = lambda k, i, j: (k * 3 + i) * 5 + j + 1   # noqa: E731
    assert ex_l["node"].tolist() == [cell(k, i, 4) for k in range(2) for i in range(3)]
    assert ex_l["node2"].tolist() == [cell(k, i, 0) for k in range(2) for i in range(3)]
    assert np.allclose(ex_l["q"], 100.0, rtol=1e-6)           # water enters the left model from the right (head 6 > 5)
    assert np.array_equal(ex_r["node"], ex_l["node2"]) and np.allclose(ex_r["q"], -ex_l["q"])
    # water balance of the left model: CHD outflow == exchange inflow; each cell's row of FLOW-JA-FACE closes
    assert np.isclose(left[1]["q"].sum(), -ex_l["q"].sum(), rtol=1e-6)
    resid = np.add.reduceat(left[0]["flow"], ml.ia[:-1])
    np.add.at(resid, left[1]["node"] - 1, left[1]["q"])
    np.add.at(resid, ex_l["node"] - 1, ex_l["q"])
    assert np.abs(resid).max() < 1e-4


def test_array_based_recharge_equals_the_list(tmp_path):
    """RCH with READASARRAYS (IRCH + RECHARGE arrays, omitted arrays keep their values) against the same
    recharge written as a list; convertible top layer + NEWTON so that recharge drives a water table"""
    rng = np.random.default_rng(3)
    shape = (2, 4, 5)
    rate = rng.uniform(1e-4, 5e-4, (4, 5))
    irch = np.ones((4, 5), dtype=int)
    irch[1, 2] = 2
    chd = [((2, i + 1, 1), 3.0) for i in range(4)]
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-8\n  OUTER_MAXIMUM 100\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 100\n  INNER_DVCLOSE 1e-9\n  INNER_RCLOSE 1e-7\n  LINEAR_ACCELERATION BICGSTAB\nEND linear\n")
    arr = lambda a: "\n".join("      " + " ".join(repr(float(v)) for v in row) for row in a)   # noqa: E731
    rcha = ("BEGIN options\n  READASARRAYS\nEND options\n\nBEGIN period 1\n  irch\n    INTERNAL FACTOR 1\n" + arr(irch)
            + "\n  recharge\n    INTERNAL FACTOR 1.0\n" + arr(rate) + "\nEND period 1\n\n"
            "BEGIN period 2\n  recharge\n    INTERNAL FACTOR 2.0\n" + arr(rate) + "\nEND period 2\n")
    rows = lambda f: "".join(f"  {irch[i, j]} {i + 1} {j + 1}  {float(f * rate[i, j])!r}\n"   # noqa: E731
                             for i in range(4) for j in range(5))
    rchl = ("BEGIN options\nEND options\n\nBEGIN dimensions\n  MAXBOUND 20\nEND dimensions\n\nBEGIN period 1\n" + rows(1.0)
            + "END period 1\n\nBEGIN period 2\n" + rows(2.0) + "END period 2\n")
    heads = {}
    for tag, text in (("arrays", rcha), ("list", rchl)):
        d = tmp_path / tag
        d.mkdir()
        mf6_inputs.write_gwf(str(d), "m", shape, 100.0, 100.0, 10.0, [0.0, -10.0], 2.0, chd={1: chd}, icelltype=1,
                             strt=5.0, newton=True, extra_packages=[("RCH6", "rch", text)])
        mf6_inputs.write_sim(str(d), ["m"], [(1.0, 1, 1.0), (1.0, 1, 1.0)], ims)
        out = simulate.run(str(d), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
        assert all(r["converged"] for r in out["reports"])
        heads[tag] = out["heads"][0]
        cbc = read_budget_file(d / "m.cbc")
        rch = [r for r in cbc if r["text"].strip() == "RCH"][-1]
        assert rch["node"].size == 20 and np.isclose(rch["q"].sum(), 2.0 * rate.sum() * 100.0 * 100.0)
    assert np.array_equal(heads["arrays"], heads["list"])
    assert heads["list"].max() > 3.0       # recharge mounds the water table above the CHD stage


def test_riv_ghb_drn_lists(tmp_path):
    """the head-dependent list packages (stage/cond/rbot, bhead/cond, elev/cond column order of the dfn files),
    with an AUXILIARY column and BOUNDNAMES that the path ignores, equal the directly built packages"""
    shape = (1, 6, 7)
    node = lambda c: ((c[0] - 1) * 6 + c[1] - 1) * 7 + c[2] - 1   # noqa: E731
    riv = [((1, i + 1, 1), 9.0, 50.0, 7.5) for i in range(6)]
    ghb = [((1, i + 1, 7), 6.0, 20.0) for i in range(6)]
    drn = [((1, 3, 4), 6.5, 80.0), ((1, 4, 4), 7.0, 60.0)]
    fmt = lambda rows, extra="": "".join("  " + " ".join(str(v) for v in r[0]) + "  "   # noqa: E731
                                         + "  ".join(repr(float(v)) for v in r[1:]) + extra + "\n" for r in rows)
    hdr = "BEGIN options\n{}END options\n\nBEGIN dimensions\n  MAXBOUND 10\nEND dimensions\n\nBEGIN period 1\n"
    extra = [("RIV6", "riv", hdr.format("  AUXILIARY conc\n  BOUNDNAMES\n") + fmt(riv, "  1.5  reach_a") + "END period 1\n"),
             ("GHB6", "ghb", hdr.format("") + fmt(ghb) + "END period 1\n"),
             ("DRN6", "drn", hdr.format("  PRINT_FLOWS\n") + fmt(drn) + "END period 1\n")]
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-9\n  OUTER_MAXIMUM 50\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 100\n  INNER_DVCLOSE 1e-10\n  INNER_RCLOSE 1e-8\n  LINEAR_ACCELERATION BICGSTAB\nEND linear\n")
    d = str(tmp_path)
    mf6_inputs.write_gwf(d, "m", shape, 50.0, 50.0, 10.0, [0.0], 3.0, strt=8.0, extra_packages=extra)
    mf6_inputs.write_sim(d, ["m"], [(1.0, 1, 1.0)], ims)
    out = simulate.run(d, ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    sim = out["simulation"]
    assert [p.ftype for p in sim.models[0].packages] == ["RIV", "GHB", "DRN"]
    m = build_dis_model(1, 6, 7, 50.0, 50.0, 10.0, [0.0], 3.0, strt=8.0)
    pk = [Package(T.PKG_RIV, [node(r[0]) for r in riv], [r[1] for r in riv], [r[2] for r in riv], [r[3] for r in riv]),
          Package(T.PKG_GHB, [node(r[0]) for r in ghb], [r[1] for r in ghb], [r[2] for r in ghb]),
          Package(T.PKG_DRN, [node(r[0]) for r in drn], [r[1] for r in drn], [r[2] for r in drn])]
    O = oracle_class()(m, sim.sln, sim.ims)
    O.set_packages(pk)
    O.timestep(1, 1, 1.0, 1)
    assert np.array_equal(out["heads"][0].ravel(), O.x)
    assert 6.0 < out["heads"][0].min() and out["heads"][0].max() < 9.0
    cbc = read_budget_file(tmp_path / "m.cbc")
    assert [r["text"].strip() for r in cbc] == ["FLOW-JA-FACE", "RIV", "GHB", "DRN"]
    assert cbc[1]["q"].sum() > 0 > cbc[2]["q"].sum() and (cbc[3]["q"] <= 0).all()      # river feeds, GHB / drains take


CHD02_HEADS = np.array([10.00, 9.57481441, 9.1298034, 8.66189866, 8.16714512, 7.64029169, 7.07409994, 6.45808724,
                        5.77600498, 5.000])


def write_chd02(d):
    """autotest/test_gwf_chd02.py:14-62: 1 x 1 x 10 convertible cells (top 10, bottom 0, K = 1), CHD 10 and 5 at
    the ends, IMS `COMPLEXITY SIMPLE`"""
    mf6_inputs.write_gwf(d, "chd02", (1, 1, 10), 1.0, 1.0, 10.0, [0.0], 1.0,
                         chd={1: [((1, 1, 1), 10.0), ((1, 1, 10), 5.0)]}, icelltype=1, strt=10.0)
    mf6_inputs.write_sim(d, ["chd02"], [(1.0, 1, 1.0)], "BEGIN options\n  COMPLEXITY SIMPLE\nEND options\n")


def test_chd02_known_answer(tmp_path):
    """the literal head array of autotest/test_gwf_chd02.py:72-87 (unconfined Picard iteration stopped by the
    SIMPLE preset's OUTER_DVCLOSE = 1e-3: reproducing it to np.allclose needs the same conductance formulation AND
    the same iteration path)"""
    write_chd02(str(tmp_path))
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    sim = out["simulation"]
    assert sim.sln.dvclose == 1e-3 and sim.sln.mxiter == 25 and sim.ims.iter1 == 50 and sim.ims.ilinmeth == 1
    assert np.allclose(CHD02_HEADS, out["heads"][0].ravel())
    assert np.abs(CHD02_HEADS - out["heads"][0].ravel()).max() < 1e-8


def test_chd02_binary_list_with_auxiliary(tmp_path):
    """exactly what autotest/test_gwf_chd02.py is about: the CHD list comes from a BINARY OPEN/CLOSE file that
    also carries two auxiliary columns; same literal answer"""
    write_chd02(str(tmp_path))
    rec = np.zeros(2, dtype=[("cellid", "<i4", (3,)), ("v", "<f8", (3,))])
    rec["cellid"] = [(1, 1, 1), (1, 1, 10)]
    rec["v"] = [(10.0, 1.0, 100.0), (5.0, 0.0, 100.0)]          # head, conc, something
    rec.tofile(tmp_path / "chd.bin")
    (tmp_path / "chd02.chd").write_text(
        "BEGIN options\n  AUXILIARY conc something\nEND options\n\nBEGIN dimensions\n  MAXBOUND 2\nEND dimensions\n\n"
        "BEGIN period 1\n  OPEN/CLOSE chd.bin (BINARY)\nEND period 1\n")
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert np.allclose(CHD02_HEADS, out["heads"][0].ravel())
    # the auxiliary columns travel into the CHD record of the budget file (save_print_model_flows: naux + 1, the
    # names, one row of values per boundary)
    chd = [r for r in read_budget_file(tmp_path / "chd02.cbc") if r["text"].strip() == "CHD"][-1]
    assert [a.strip() for a in chd["auxtxt"]] == ["CONC", "SOMETHING"]
    assert np.array_equal(chd["aux"], [[1.0, 100.0], [0.0, 100.0]]) and chd["node"].tolist() == [1, 10]
    # and the text flavour of OPEN/CLOSE
    (tmp_path / "chd.txt").write_text("1 1 1 10.0 1.0 100.0\n1 1 10 5.0 0.0 100.0\n")
    (tmp_path / "chd02.chd").write_text((tmp_path / "chd02.chd").read_text().replace("chd.bin (BINARY)", "chd.txt"))
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert np.allclose(CHD02_HEADS, out["heads"][0].ravel())


RCH01_IMS = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-9\n  OUTER_MAXIMUM 100\n  UNDER_RELAXATION DBD\nEND nonlinear\n\n"
             "BEGIN linear\n  INNER_MAXIMUM 300\n  INNER_DVCLOSE 1e-9\n  INNER_RCLOSE 1e-3\n  LINEAR_ACCELERATION BICGSTAB\n"
             "  SCALING_METHOD NONE\n  REORDERING_METHOD NONE\n  RELAXATION_FACTOR 0.97\nEND linear\n")


def write_rch01(d, irch):
    """autotest/test_gwf_rch01.py:21-120: 2 layers x 5 columns of convertible cells, the top layer dry except its
    middle cell (starting head 25 under a layer bottom of 50), array-based recharge 0.1 on the top layer"""
    strt = np.array([[[25.0, 25.0, 75.0, 25.0, 25.0]], [[25.0, 25.0, 75.0, 25.0, 25.0]]])
    per = "BEGIN period 1\n"
    if irch is not None:
        per += "  irch\n    INTERNAL FACTOR 1\n      " + " ".join(str(v) for v in irch) + "\n"
    per += "  recharge\n    CONSTANT 0.1\nEND period 1\n"
    mf6_inputs.write_gwf(d, "rch", (2, 1, 5), 1.0, 1.0, 100.0, [50.0, 0.0], 1.0, icelltype=1, strt=strt,
                         chd={1: [((2, 1, 1), 25.0), ((2, 1, 5), 25.0)]}, sto=dict(ss=1e-5, sy=0.1, periods={}),
                         extra_packages=[("RCH6", "rcha", "BEGIN options\n  READASARRAYS\nEND options\n\n" + per)])
    mf6_inputs.write_sim(d, ["rch"], [(0.01, 1, 1.0)], RCH01_IMS)


@pytest.mark.parametrize("irch", [None, [1, 1, 1, 1, 1], [2, 2, 1, 2, 2]])
def test_rch01_recharge_reaches_the_highest_active_cell(tmp_path, irch):
    """autotest/test_gwf_rch01.py:124-130: the RCH budget records -- recharge listed on dry top cells is applied
    to the cell below (node 7, 9), stays on the wet middle cell (node 3) and is dropped on the constant heads
    (nodes 6, 10): needs the wet/dry conversion of npf_cf and rch_cf's highest_active"""
    write_rch01(str(tmp_path), irch)
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert out["reports"][0]["converged"] == 1
    rec = [r for r in read_budget_file(tmp_path / "rch.cbc") if r["text"].strip() == "RCH"][0]
    assert rec["node"].tolist() == [6, 7, 3, 9, 10] and rec["node2"].tolist() == [1, 2, 3, 4, 5]
    assert np.allclose(rec["q"], [0.0, 0.1, 0.1, 0.1, 0.0])
    hds = read_head_file(tmp_path / "rch.hds")
    assert (hds[0]["data"].ravel()[[0, 1, 3, 4]] == -1.0e30).all()      # the dry cells carry HDRY
    assert hds[0]["data"].ravel()[2] > 50.0


def test_idomain_reduced_numbering_and_pass_through(tmp_path):
    """IDOMAIN like the reference treats it (disconnections, Connections.f90:463-700): cells with idomain <= 0 do
    not exist (reduced node numbers), idomain < 0 connects the cells above and below it directly.  A confined
    3-layer model whose whole middle layer is a pass-through equals the 2-layer model made of its outer layers;
    a hole (idomain == 0) shows up as 1e30 in the head file, list records keep USER node numbers."""
    rng = np.random.default_rng(8)
    nrow, ncol = 5, 6
    k3 = np.exp(rng.normal(0.3, 0.5, (3, nrow, ncol)))
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-10\n  OUTER_MAXIMUM 50\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 200\n  INNER_DVCLOSE 1e-11\n  INNER_RCLOSE 1e-9\n  LINEAR_ACCELERATION CG\nEND linear\n")
    idom = np.ones((3, nrow, ncol), dtype=int)
    idom[1] = -1
    idom[0, 2, 3] = 0                                      # a hole in the top layer
    chd3 = [((kk, i + 1, 1), 5.0) for kk in (1, 3) for i in range(nrow)] + \
           [((kk, i + 1, ncol), 2.0) for kk in (1, 3) for i in range(nrow)]
    wel3 = [((3, 3, 3), -30.0)]
    a = tmp_path / "three"
    a.mkdir()
    # a boundary in the hole is an input error, as in the reference (not silently dropped)
    mf6_inputs.write_gwf(str(a), "m", (3, nrow, ncol), 20.0, 25.0, 0.0, [-4.0, -10.0, -18.0], k3, chd={1: chd3},
                         wel={1: wel3 + [((1, 3, 4), -10.0)]}, strt=3.0, k33=0.2, idomain=idom)
    mf6_inputs.write_sim(str(a), ["m"], [(1.0, 1, 1.0)], ims)
    with pytest.raises(mf6io.Mf6InputError, match="IDOMAIN removes"):
        mf6io.read_simulation(str(a))
    mf6_inputs.write_gwf(str(a), "m", (3, nrow, ncol), 20.0, 25.0, 0.0, [-4.0, -10.0, -18.0], k3, chd={1: chd3},
                         wel={1: wel3}, strt=3.0, k33=0.2, idomain=idom)
    oa = simulate.run(str(a), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    gi = oa["simulation"].models[0]
    assert gi.model.nodes == 2 * nrow * ncol - 1 and gi.nodeuser.size == gi.model.nodes
    # the equivalent 2-layer model: thicknesses 4 and 8, same K per layer, the hole as idomain 0 again
    b = tmp_path / "two"
    b.mkdir()
    idom2 = np.ones((2, nrow, ncol), dtype=int)
    idom2[0, 2, 3] = 0
    chd2 = [((1 if c[0] == 1 else 2, c[1], c[2]), v) for c, v in chd3]
    mf6_inputs.write_gwf(str(b), "m", (2, nrow, ncol), 20.0, 25.0, 0.0, [-4.0, -12.0], k3[[0, 2]], chd={1: chd2},
                         wel={1: [((2, 3, 3), -30.0)]}, strt=3.0, k33=0.2, idomain=idom2)
    mf6_inputs.write_sim(str(b), ["m"], [(1.0, 1, 1.0)], ims)
    ob = simulate.run(str(b), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    ha, hb = oa["heads"][0], ob["heads"][0]
    assert np.array_equal(ha[0], hb[0]) and np.array_equal(ha[2], hb[1])         # identical systems
    assert (ha[1] == 1.0e30).all() and ha[0, 2, 3] == 1.0e30 and ha[0].min() > 1.9
    hds = read_head_file(a / "m.hds")
    assert len(hds) == 3 and (hds[1]["data"] == 1.0e30).all() and hds[0]["data"][2, 3] == 1.0e30
    cbc = read_budget_file(a / "m.cbc")
    fja, wel = cbc[0], [r for r in cbc if r["text"].strip() == "WEL"][0]
    assert fja["flow"].size == gi.model.nja                                     # the reduced connectivity
    assert wel["node"].tolist() == [(2 * nrow + 2) * ncol + 3] and np.isclose(wel["q"][0], -30.0)   # USER node


def test_models_with_different_options_are_refused(tmp_path):
    """a two-model solution shares one set of NPF / STO options on the GPU path: models that differ (here NEWTON
    in one of them) must be refused, not silently solved with the first model's options"""
    from modflow6_b200.grid import build_dis_model, merge_models
    a = build_dis_model(1, 1, 5, 1.0, 1.0, 0.0, [-1.0], 1.0)
    b = build_dis_model(1, 1, 5, 1.0, 1.0, 0.0, [-1.0], 1.0, icelltype=1, inewton=1)
    ex = dict(m1=0, m2=1, nodem1=[4], nodem2=[0], ihc=[1], cl1=[0.5], cl2=[0.5], hwva=[1.0])
    with pytest.raises(ValueError, match="inewton"):
        merge_models([a, b], [ex])
    merge_models([a, a], [ex])


def test_npf05_anisotropy_and_drn_depth_from_decks(tmp_path):
    """the deck reader hands K22 / K22OVERK / K33OVERK / ANGLEn (degrees -> radians) and the DRN AUXDEPTHNAME column
    to the path: (1) autotest/test_gwf_npf05_anisotropy.py written as input files gives its literal head array;
    (2) a strip along y with K22 and ANGLE1 = 90 equals the same strip along x (hy_eff)"""
    import os
    from tests.test_oracle_known_answers import NPF05_ANSWER
    d = str(tmp_path / "a")
    os.makedirs(d)
    mf6_inputs.write_gwf(d, "npf", (2, 1, 5), 1.0, 1.0, 100.0, [50.0, 0.0], 5.0, strt=100.0,
                         chd={1: [((1, 1, 1), 100.0), ((2, 1, 5), 110.0)]},
                         extra_packages=[("RCH6", "rcha", "BEGIN options\n  READASARRAYS\nEND options\n\nBEGIN period 1\n"
                                          "  RECHARGE\n    CONSTANT 0.01\nEND period 1\n")])
    npf = (tmp_path / "a" / "npf.npf").read_text()
    npf = npf.replace("SAVE_FLOWS", "SAVE_FLOWS\n  K22OVERK\n  K33OVERK").replace(
        "END griddata", "  k22\n    CONSTANT 0.1\n  k33\n    CONSTANT 0.01\nEND griddata")
    (tmp_path / "a" / "npf.npf").write_text(npf)
    mf6_inputs.write_sim(d, ["npf"], [(1.0, 1, 1.0)],
                         "BEGIN nonlinear\n  OUTER_DVCLOSE 1e-9\n  OUTER_MAXIMUM 100\n  UNDER_RELAXATION DBD\nEND nonlinear\n\n"
                         "BEGIN linear\n  INNER_MAXIMUM 300\n  INNER_DVCLOSE 1e-9\n  INNER_RCLOSE 1e-3\n"
                         "  LINEAR_ACCELERATION BICGSTAB\n  RELAXATION_FACTOR 0.97\nEND linear\n")
    sim = mf6io.read_simulation(d)
    m = sim.models[0].model
    assert np.allclose(m.k22, 0.5) and np.allclose(m.k33, 0.05) and m.conn_nx is not None
    out = simulate.run(d, ordering=T.ORDER_NATURAL, solution_class=oracle_class(), write_output=False)
    assert np.allclose(out["heads"][0].ravel(), NPF05_ANSWER)
    # DRN with a drainage-depth auxiliary column
    d2 = str(tmp_path / "b")
    os.makedirs(d2)
    drn = ("BEGIN options\n  AUXILIARY ddrn\n  AUXDEPTHNAME ddrn\nEND options\n\nBEGIN dimensions\n  MAXBOUND 1\n"
           "END dimensions\n\nBEGIN period 1\n  1 1 5  0.0  2.5  1.0\nEND period 1\n")
    mf6_inputs.write_gwf(d2, "m", (1, 1, 5), 1.0, 1.0, 10.0, [0.0], 1.0, strt=5.0, icelltype=1,
                         chd={1: [((1, 1, 1), 5.0)]}, extra_packages=[("DRN6", "drn", drn)])
    mf6_inputs.write_sim(d2, ["m"], [(1.0, 1, 1.0)], "")
    sp = [p for p in mf6io.read_simulation(d2).models[0].packages if p.ftype.upper().startswith("DRN")][0]
    pk = sp.periods[1]
    assert pk.b1.tolist() == [0.0] and pk.b2.tolist() == [2.5] and pk.b3.tolist() == [1.0] and pk.iflowred == 0


def test_npf02_rewet_from_deck(tmp_path):
    """autotest/test_gwf_npf02_rewet.py (case c: 3 layers) written as input files -- REWET record + WETDRY array
    through the deck reader, two stress periods with a changing CHD list -- gives the reference's literal heads"""
    from tests.test_oracle_known_answers import NPF02_3LAY, npf02_profile
    d = str(tmp_path)
    nlay, nrow, ncol = 3, 10, 15
    botm = [50.0, 0.0, -50.0]
    def chd(vl):
        rows = [((k + 1, i + 1, 1), vl) for k in range(nlay) for i in range(nrow) if botm[k] < vl]
        return rows + [((k + 1, i + 1, ncol), -40.0) for k in range(nlay) for i in range(nrow) if botm[k] < -40.0]
    mf6_inputs.write_gwf(d, "m", (nlay, nrow, ncol), 500.0, 500.0, 150.0, botm, 10.0, strt=-40.0, icelltype=1,
                         chd={1: chd(100.0), 2: chd(25.0)})
    npf = (tmp_path / "m.npf").read_text()
    npf = npf.replace("SAVE_FLOWS", "SAVE_FLOWS\n  REWET WETFCT 1.0 IWETIT 1 IHDWET 1").replace(
        "END griddata", "  wetdry\n    CONSTANT -0.001\nEND griddata")
    (tmp_path / "m.npf").write_text(npf)
    mf6_inputs.write_sim(d, ["m"], [(1.0, 1, 1.0), (1.0, 1, 1.0)],
                         "BEGIN nonlinear\n  OUTER_DVCLOSE 1e-1\n  OUTER_MAXIMUM 1000\nEND nonlinear\n\n"
                         "BEGIN linear\n  INNER_MAXIMUM 100\n  INNER_DVCLOSE 1e-1\n  INNER_RCLOSE 0.01\n"
                         "  LINEAR_ACCELERATION CG\n  RELAXATION_FACTOR 1.0\nEND linear\n")
    out = simulate.run(d, ordering=T.ORDER_NATURAL, solution_class=oracle_class(), write_output=False)
    assert all(r["converged"] for r in out["reports"])
    assert np.abs(npf02_profile(out["heads"][0].ravel(), nlay) - NPF02_3LAY[1]).max() < 1e-9


@pytest.mark.parametrize("idx", [2, 6, 8])
def test_npf_thickstrt_and_hfb_from_decks(tmp_path, idx):
    """autotest/test_gwf_npf_thickstrt.py from input FILES: NPF THICKSTRT with a negative ICELLTYPE and the HFB6
    package (PERIOD list of cellid1 cellid2 hydchr) reach the reference's literal heads and CHD inflow"""
    from tests.helpers import NPF_THICKSTRT, npf_thickstrt_case
    _, hfb, heads, inflow = npf_thickstrt_case(idx)
    d = str(tmp_path)
    extra = []
    if hfb:
        extra.append(("HFB6", "hfb", "BEGIN options\n  PRINT_INPUT\nEND options\n\nBEGIN dimensions\n  MAXHFB 1\n"
                      "END dimensions\n\nBEGIN period 1\n  1 1 3  1 1 4  1.0e-4\nEND period 1\n"))
    mf6_inputs.write_gwf(d, "flow", (1, 1, 6), 1.0, 1.0, 10.0, [0.0], 1.0, icelltype=NPF_THICKSTRT["icelltype"][idx],
                         chd={1: [((1, 1, 1), 6.0), ((1, 1, 6), 4.0)]}, strt=5.0, k33=1.0, extra_packages=extra)
    if NPF_THICKSTRT["thickstrt"][idx]:
        p = tmp_path / "flow.npf"
        p.write_text(p.read_text().replace("  SAVE_FLOWS\n", "  SAVE_FLOWS\n  THICKSTRT\n"))
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-6\n  OUTER_MAXIMUM 10\n  UNDER_RELAXATION NONE\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 5\n  INNER_DVCLOSE 1e-6\n  INNER_RCLOSE 1e-6\n  LINEAR_ACCELERATION CG\n"
           "  SCALING_METHOD NONE\n  REORDERING_METHOD NONE\n  RELAXATION_FACTOR 1.0\nEND linear\n")
    mf6_inputs.write_sim(d, ["flow"], [(1.0, 1, 1.0)], ims)
    out = simulate.run(d, ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert np.allclose(heads, out["heads"][0].ravel())
    cbc = read_budget_file(tmp_path / "flow.cbc")
    assert cbc[1]["text"].strip() == "CHD" and np.allclose(inflow, cbc[1]["q"][0])
    # a barrier between cells that are not connected is an input error (check_data, gwf-hfb.f90:714-766)
    if hfb:
        (tmp_path / "flow.hfb").write_text("BEGIN dimensions\n  MAXHFB 1\nEND dimensions\n\nBEGIN period 1\n"
                                           "  1 1 2  1 1 4  1.0e-4\nEND period 1\n")
        with pytest.raises(mf6io.Mf6InputError, match="not connected"):
            simulate.run(d, ordering=T.ORDER_NATURAL, solution_class=oracle_class(), write_output=False)


def test_idomain_holes_on_disv_and_disu_decks_equal_the_dis_deck(tmp_path):
    """IDOMAIN == 0 on DISV / DISU: reduced node numbering like on DIS (grid.reduce_model) -- the same rectangular
    grid with the same holes written as DIS, DISV and DISU gives the same reduced connectivity, bottom nodes, heads
    and FLOW-JA-FACE length; the head files carry 1e30 in the holes"""
    rng = np.random.default_rng(21)
    shape = (3, 4, 5)
    k = np.exp(rng.normal(0.5, 0.6, shape))
    idom = np.ones(shape, dtype=int)
    idom[0, 1, 2] = idom[1, 1, 2] = idom[2, 3, 0] = idom[1, 0, 4] = 0
    chd = [((kk + 1, i + 1, 1), 5.0) for kk in range(3) for i in range(4) if idom[kk, i, 0]] + \
          [((kk + 1, i + 1, 5), 2.0) for kk in range(3) for i in range(4) if idom[kk, i, 4]]
    wel = [((3, 3, 3), -40.0)]
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-9\n  OUTER_MAXIMUM 50\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 200\n  INNER_DVCLOSE 1e-10\n  INNER_RCLOSE 1e-8\n  LINEAR_ACCELERATION CG\nEND linear\n")
    outs = {}
    for tag in ("dis", "disv", "disu"):
        d = tmp_path / tag
        d.mkdir()
        mf6_inputs.write_gwf(str(d), "m", shape, 10.0, 12.0, 0.0, [-5.0, -12.0, -30.0], k, chd={1: chd},
                             wel={1: wel}, strt=3.0, k33=0.3, disv=(tag == "disv"), disu=(tag == "disu"),
                             idomain=idom if tag == "dis" else None)
        if tag != "dis":
            p = d / "m.dis"
            arr = mf6_inputs._arr("idomain", idom.astype(float), layered=True) if tag == "disv" else \
                mf6_inputs._arr("idomain", idom.astype(float).reshape(1, -1))
            p.write_text(p.read_text().replace("END griddata\n", arr + "END griddata\n"))
        mf6_inputs.write_sim(str(d), ["m"], [(1.0, 1, 1.0)], ims)
        outs[tag] = simulate.run(str(d), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    a = outs["dis"]["simulation"].models[0].model
    assert a.nodes == idom.sum()
    for tag in ("disv", "disu"):
        gi = outs[tag]["simulation"].models[0]
        b = gi.model
        assert b.nodes == a.nodes and np.array_equal(gi.nodeuser, np.nonzero(idom.reshape(-1))[0])
        for name in ("ia", "ja", "jas", "isym", "ihc", "ibotnode"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), (tag, name)
        for name in ("cl1", "cl2", "hwva", "area", "top", "bot", "k11"):
            assert np.allclose(getattr(a, name), getattr(b, name), rtol=1e-13), (tag, name)
        ha, hb = outs["dis"]["heads"][0].ravel(), outs[tag]["heads"][0].ravel()
        assert np.array_equal(ha == 1.0e30, hb == 1.0e30) and (hb == 1.0e30).sum() == 4
        assert np.abs(ha - hb).max() < 1e-9
        cbc = read_budget_file(tmp_path / tag / "m.cbc")
        assert cbc[0]["flow"].size == a.nja
        w = [r for r in cbc if r["text"].strip() == "WEL"][0]
        assert w["node"].tolist() == [(2 * 4 + 2) * 5 + 3]          # USER node number


def _write_rch0203(d, idomain, irch):
    """autotest/test_gwf_rch02.py:14-106 / test_gwf_rch03.py:14-128: 2 x 4 x 5 confined cells (top 100, bottoms 50 and
    0, K 1), CHD 100 at (2, 1, 1), array-based recharge = 2-D cell number (1 .. 20), IDOMAIN with removed (0) and
    pass-through (-1) cells in the top layer"""
    per = "BEGIN period 1\n"
    if irch is not None:
        per += "  irch\n    INTERNAL FACTOR 1\n" + "".join("      " + " ".join(str(v) for v in row) + "\n" for row in irch)
    per += "  recharge\n    INTERNAL FACTOR 1.0\n" + "".join(
        "      " + " ".join(repr(float(v)) for v in row) + "\n" for row in np.arange(20).reshape(4, 5) + 1.0)
    per += "END period 1\n"
    mf6_inputs.write_gwf(d, "rch", (2, 4, 5), 1.0, 1.0, 100.0, [50.0, 0.0], 1.0, icelltype=0, strt=100.0,
                         chd={1: [((2, 1, 1), 100.0)]}, idomain=idomain,
                         extra_packages=[("RCH6", "rcha", "BEGIN options\n  READASARRAYS\nEND options\n\n" + per)])
    mf6_inputs.write_sim(d, ["rch"], [(1.0, 1, 1.0)], RCH01_IMS)


def test_rch02_array_recharge_over_pass_through_cells(tmp_path):
    """autotest/test_gwf_rch02.py:109-122: recharge given for pass-through cells (IDOMAIN -1) goes nowhere; every
    record that is written has node == node2 == q (the recharge array holds the 2-D cell numbers)"""
    idom = np.ones((2, 4, 5), dtype=int)
    idom[0, 1:3, 1:4] = -1
    _write_rch0203(str(tmp_path), idom, None)
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert out["reports"][0]["converged"] == 1
    rec = [r for r in read_budget_file(tmp_path / "rch.cbc") if r["text"].strip() == "RCH"][0]
    assert rec["node"].size == 14
    assert np.allclose(rec["node"].astype(float), rec["q"]) and np.allclose(rec["node2"], rec["node"])


def test_rch03_irch_and_idomain_literal_records(tmp_path):
    """autotest/test_gwf_rch03.py:130-146: the literal RCH budget records (user nodes 21 27 8 32 13 34, bound
    numbers 1 7 8 12 13 14, rates 0 7 8 12 13 14) of IRCH pointing at removed, pass-through and constant-head cells"""
    idom = np.array([[[0, 0, 0, 0, 0], [0, -1, 1, -1, 0], [0, -1, 1, -1, 0], [0, 0, 0, 0, 0]],
                     [[1, 1, 1, 1, 1], [1, 1, 1, -1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]]])
    irch = np.array([[1, 0, 0, 0, 0], [0, 1, 0, 1, 0], [0, 1, 0, 1, 0], [0, 0, 0, 0, 0]]) + 1   # flopy writes 1-based
    _write_rch0203(str(tmp_path), idom, irch)
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert out["reports"][0]["converged"] == 1
    rec = [r for r in read_budget_file(tmp_path / "rch.cbc") if r["text"].strip() == "RCH"][0]
    assert rec["node"].tolist() == [21, 27, 8, 32, 13, 34]
    assert rec["node2"].tolist() == [1, 7, 8, 12, 13, 14]
    assert np.allclose(rec["q"], [0.0, 7.0, 8.0, 12.0, 13.0, 14.0])


def write_pertim(d):
    """autotest/test_gwf_pertim.py:10-96: 3 x 21 x 20 confined cells (K 50 / 0.01 / 200, K33 10 / 0.01 / 20), two CHD
    packages with BOUNDNAMES (canal 330 on the first column, river 320 on the last), ONE steady period of LENGTH 0,
    IMS COMPLEXITY SIMPLE"""
    nlay, nrow, ncol = 3, 21, 20
    mf6_inputs.write_gwf(d, "gwf_pertim", (nlay, nrow, ncol), 500.0, 500.0, 330.0, [220.0, 200.0, 0.0],
                         np.broadcast_to(np.array([50.0, 0.01, 200.0])[:, None, None], (nlay, nrow, ncol)),
                         strt=330.0,
                         k33=np.broadcast_to(np.array([10.0, 0.01, 20.0])[:, None, None], (nlay, nrow, ncol)))
    for tag, col, head in (("canal", 1, 330.0), ("river", ncol, 320.0)):
        with open(f"{d}/gwf_pertim_{tag}.chd", "w") as f:
            f.write(f"BEGIN options\n  BOUNDNAMES\nEND options\n\nBEGIN dimensions\n  MAXBOUND {nrow}\nEND dimensions\n\n"
                    "BEGIN period 1\n" + "".join(f"  1 {i + 1} {col} {head!r} {tag}\n" for i in range(nrow))
                    + "END period 1\n")
    nam = f"{d}/gwf_pertim.nam"
    text = open(nam).read().replace("END packages", "  CHD6  gwf_pertim_canal.chd  CHD-CANAL\n"
                                    "  CHD6  gwf_pertim_river.chd  CHD-RIVER\nEND packages")
    open(nam, "w").write(text)
    oc = f"{d}/gwf_pertim.oc"
    text = open(oc).read().replace("  PRINT HEAD LAST\n", "  PRINT HEAD LAST\n  PRINT BUDGET ALL\n")
    open(oc, "w").write(text)
    mf6_inputs.write_sim(d, ["gwf_pertim"], [(0.0, 1, 1.0)], "BEGIN options\n  COMPLEXITY SIMPLE\nEND options\n")


def test_pertim_zero_length_period_literal_chd_flows(tmp_path):
    """autotest/test_gwf_pertim.py:99-115: a steady stress period of length ZERO; the listing budget's CHD_IN
    99928.4941 and CHD2_OUT 99928.5036 (np.allclose) = inflow through the canal, outflow through the river"""
    write_pertim(str(tmp_path))
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert out["reports"][0]["converged"] == 1
    cbc = read_budget_file(tmp_path / "gwf_pertim.cbc")
    canal, river = [r for r in cbc if r["text"].strip() == "CHD"]
    assert canal["srcpackage"].strip() == "CHD-CANAL" and river["srcpackage"].strip() == "CHD-RIVER"
    assert np.allclose([canal["q"][canal["q"] > 0].sum()], [99928.4941])
    assert np.allclose([-river["q"][river["q"] < 0].sum()], [99928.5036])
    # and where the reference test reads them: the VOLUME BUDGET table of the model listing file (CHD_IN of the
    # first CHD package, CHD2_OUT of the second)
    from modflow6_b200.output import read_listing_budgets
    bud = read_listing_budgets(tmp_path / "gwf_pertim.lst")
    assert len(bud) == 1 and (bud[0]["kstp"], bud[0]["kper"]) == (1, 1)
    assert np.allclose([bud[0]["rates_in"][("CHD", "CHD-CANAL")]], [99928.4941])
    assert np.allclose([bud[0]["rates_out"][("CHD", "CHD-RIVER")]], [99928.5036])
    assert bud[0]["volumes_in"][("CHD", "CHD-CANAL")] == 0.0          # a period of length zero moves no volume
    assert abs(bud[0]["pdiffr"]) < 0.01 and bud[0]["totim_seconds"] == 0.0


def write_auxmult(d, idx):
    """autotest/test_gwf_utl04_auxmult.py:14-160: 1 x 3 x 3 confined cells, CHD 1 at (1,1,1), one well at (1,3,3) whose
    rate is a time series ("tsq", constant 1) or the number 1, multiplied by the auxiliary variable AUXMULT, which is
    the time series "tsqfact" (0, 1, 0, 1 ... at t = 0, 0.1, ... 1, 2, 3, 4; LINEAREND); 10 + 1 + 1 + 1 time steps"""
    wel = ("BEGIN options\n  AUXILIARY auxmult\n  AUXMULTNAME auxmult\n  PRINT_INPUT\n  TS6 FILEIN m.wel.ts\nEND options\n\n"
           "BEGIN dimensions\n  MAXBOUND 1\nEND dimensions\n\nBEGIN period 1\n  1 3 3 "
           + ("tsq" if idx == 0 else "1.0000000") + " tsqfact\nEND period 1\n")
    mf6_inputs.write_gwf(d, "m", (1, 3, 3), 100.0, 100.0, 0.0, [-1.0], 1.0, chd={1: [((1, 1, 1), 1.0)]}, strt=0.0,
                         k33=1.0, extra_packages=[("WEL6", "wel", wel)])
    t = [0.1 * i for i in range(11)] + [2.0, 3.0, 4.0]
    with open(f"{d}/m.wel.ts", "w") as f:
        f.write("BEGIN attributes\n  NAMES tsqfact tsq\n  METHODS linearend linearend\nEND attributes\n\nBEGIN timeseries\n"
                + "".join(f"  {tt!r}  {float(i % 2)!r}  1.0\n" for i, tt in enumerate(t)) + "END timeseries\n")
    oc = f"{d}/m.oc"
    text = open(oc).read().replace("SAVE BUDGET LAST", "SAVE BUDGET ALL")
    open(oc, "w").write(text)
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-6\n  OUTER_MAXIMUM 100\n  UNDER_RELAXATION NONE\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 300\n  INNER_DVCLOSE 1e-6\n  INNER_RCLOSE 1e-3\n  LINEAR_ACCELERATION CG\n"
           "  SCALING_METHOD NONE\n  REORDERING_METHOD NONE\n  RELAXATION_FACTOR 1.0\nEND linear\n")
    mf6_inputs.write_sim(d, ["m"], [(1.0, 10, 1.0), (1.0, 1, 1.0), (1.0, 1, 1.0), (1.0, 1, 1.0)], ims)


@pytest.mark.parametrize("idx", [0, 1])
def test_auxmult_with_time_series_literal_rates(tmp_path, idx):
    """autotest/test_gwf_utl04_auxmult.py:163-182: the well rate of every time step in the budget file is
    1, 0, 1, 0 ... 1 (13 steps) -- TS6 time series (LINEAREND = the value at the end of the step) and AUXMULTNAME"""
    write_auxmult(str(tmp_path), idx)
    simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    recs = [r for r in read_budget_file(tmp_path / "m.cbc") if r["text"].strip() == "WEL"]
    q = np.array([r["q"][0] for r in recs])
    assert np.allclose(q, np.array(7 * [1.0, 0.0])[:-1])
    assert [a.strip() for a in recs[0]["auxtxt"]] == ["AUXMULT"]
    assert np.allclose([r["aux"][0, 0] for r in recs], np.array(7 * [1.0, 0.0])[:-1])


def test_time_series_interpolation_methods():
    """TimeSeries.f90 GetValue: STEPWISE and LINEAR give the time-weighted average over the step, LINEAREND the value
    at its end; a step beyond the last record is an error (no extension)"""
    from modflow6_b200.timeseries import TimeSeries, TimeSeriesError
    t, v = [0.0, 1.0, 3.0], [2.0, 4.0, 0.0]
    sw, li, le = (TimeSeries("s", m, t, v) for m in ("STEPWISE", "LINEAR", "LINEAREND"))
    assert sw.value(0.0, 1.0) == 2.0 and np.isclose(sw.value(0.5, 2.0), (0.5 * 2.0 + 1.0 * 4.0) / 1.5)
    assert np.isclose(li.value(0.0, 1.0), 3.0) and np.isclose(li.value(0.5, 2.0), (0.5 * 3.5 + 1.0 * 3.0) / 1.5)
    assert np.isclose(le.value(0.5, 2.0), 2.0) and le.value(0.0, 3.0) == 0.0
    assert sw.value(1.0, 1.0) == 4.0 and np.isclose(li.value(2.0, 2.0), 2.0)
    for s in (sw, li, le):
        with pytest.raises(TimeSeriesError):
            s.value(2.0, 3.5)


def write_lgr_gnc(d, gnc=True):
    """local grid refinement the way MODFLOW 6 decks do it: a parent DIS model (4 x 3 cells of 2 x 2) and a child DIS
    model (8 x 4 cells of 1 x 1) joined by a GWF-GWF exchange with a GNC6 file (one contributing parent cell per
    connection, alpha 0.25); constant heads on h = 10 + 0.7 x + 0.3 y along the outer boundary.
    Returns the exact heads of both models."""
    hp = np.array([[10 + 0.7 * (2 * j - 1) + 0.3 * (9 - 2 * i) for j in range(1, 4)] for i in range(1, 5)])
    hc = np.array([[10 + 0.7 * (5.5 + j) + 0.3 * (8.5 - i) for j in range(1, 5)] for i in range(1, 9)])
    chd_p = [((1, i, j), float(hp[i - 1, j - 1])) for i in range(1, 5) for j in range(1, 4) if j == 1 or i in (1, 4)]
    chd_c = [((1, i, j), float(hc[i - 1, j - 1])) for i in range(1, 9) for j in range(1, 5) if j == 4 or i in (1, 8)]
    mf6_inputs.write_gwf(d, "parent", (1, 4, 3), 2.0, 2.0, 1.0, [0.0], 1.0, chd={1: chd_p}, strt=10.0)
    mf6_inputs.write_gwf(d, "child", (1, 8, 4), 1.0, 1.0, 1.0, [0.0], 1.0, chd={1: chd_c}, strt=10.0)
    rows, gncrows = [], []
    for i in range(1, 5):
        for ic, ij in ((2 * i - 1, i - 1), (2 * i, i + 1)):      # upper / lower child row, parent row on that side
            rows.append(((1, i, 3), (1, ic, 1), 1, 1.0, 0.5, 1.0))
            if 1 <= ij <= 4:
                gncrows.append(f"  1 {i} 3  1 {ic} 1  1 {ij} 3  0.25\n")
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-10\n  OUTER_MAXIMUM 200\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 200\n  INNER_DVCLOSE 1e-12\n  INNER_RCLOSE 1e-12\n  LINEAR_ACCELERATION CG\nEND linear\n")
    mf6_inputs.write_sim(d, ["parent", "child"], [(1.0, 1, 1.0)], ims, exchanges=[("lgr", "parent", "child", rows)])
    if gnc:
        p = f"{d}/lgr.gwfgwf"
        text = open(p).read().replace("  SAVE_FLOWS\n", "  SAVE_FLOWS\n  GNC6 FILEIN lgr.gnc\n")
        open(p, "w").write(text)
        with open(f"{d}/lgr.gnc", "w") as f:
            f.write("BEGIN options\n  EXPLICIT\nEND options\n\nBEGIN dimensions\n"
                    f"  NUMGNC {len(gncrows)}\n  NUMALPHAJ 1\nEND dimensions\n\nBEGIN gncdata\n" + "".join(gncrows)
                    + "END gncdata\n")
    return hp, hc


def test_lgr_exchange_with_ghost_nodes_from_decks(tmp_path):
    """parent + refined child model + GWF-GWF exchange + GNC6 from input FILES: the linear head field is reproduced
    exactly with the ghost nodes and is centimetres off without them"""
    for tag, gnc in (("gnc", True), ("plain", False)):
        d = tmp_path / tag
        d.mkdir()
        hp, hc = write_lgr_gnc(str(d), gnc)
        out = simulate.run(str(d), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
        assert out["reports"][0]["converged"] == 1
        err = max(np.abs(out["heads"][0].reshape(4, 3) - hp).max(), np.abs(out["heads"][1].reshape(8, 4) - hc).max())
        assert (err < 1e-8) if gnc else (err > 1e-2)


@pytest.mark.parametrize("grid", ["disv", "disu"])
def test_disv_disu_connectivity_with_an_inactive_cell(tmp_path, grid):
    """autotest/test_gwf_disv.py:56-72 and test_gwf_disu.py:56-72, case b: a 3 x 3 x 3 grid of 10 x 10 x 10 cells
    (DISV with vertex coordinates offset by 1e8) whose cell 2 has IDOMAIN 0 -- the reference asserts the connectivity
    of its binary grid file in USER numbering: ia[0:4] = 1 4 4 7 (the removed cell's row is empty), ja[:6] =
    1 4 10 3 6 12, ia[-1] = 127, 28 / 126 entries.  Rebuilt here from the reduced model and nodeuser."""
    shape = (3, 3, 3)
    d = str(tmp_path)
    mf6_inputs.write_gwf(d, "m", shape, 10.0, 10.0, 0.0, [-10.0, -20.0, -30.0], 1.0, strt=0.0,
                         chd={1: [((1, 1, 1), 1.0), ((1, 3, 3), 0.0)]}, disv=(grid == "disv"), disu=(grid == "disu"))
    idom = np.ones(shape)
    idom[0, 0, 1] = 0
    p = tmp_path / "m.dis"
    if grid == "disv":
        p.write_text("# test\n" + mf6_inputs._disv_text(shape, 10.0, 10.0, 0.0, [-10.0, -20.0, -30.0], 1.0e8, 1.0e8))
        arr = mf6_inputs._arr("idomain", idom, layered=True)
    else:
        arr = mf6_inputs._arr("idomain", idom.reshape(1, -1))
    p.write_text(p.read_text().replace("END griddata\n", arr + "END griddata\n"))
    mf6_inputs.write_sim(d, ["m"], [(1.0, 1, 1.0)], "BEGIN options\n  PRINT_OPTION SUMMARY\nEND options\n")
    gi = mf6io.read_simulation(d).models[0]
    m, nodeuser = gi.model, gi.nodeuser
    assert m.nodes == 26 and np.allclose(m.area, 100.0, rtol=1e-12)     # the 1e8 offsets do not hurt the areas
    ia = np.zeros(28, dtype=int)
    ja = []
    for r in range(m.nodes):
        row = nodeuser[m.ja[m.ia[r]:m.ia[r + 1]]] + 1
        ja += [int(row[0])] + sorted(int(v) for v in row[1:])
        ia[nodeuser[r] + 1] = m.ia[r + 1] - m.ia[r]
    ia = 1 + np.concatenate([[0], np.cumsum(ia[1:])])
    assert np.array_equal(ia[0:4], [1, 4, 4, 7]) and ia[-1] == 127 and ia.shape[0] == 28
    assert ja[:6] == [1, 4, 10, 3, 6, 12] and len(ja) == 126
    out = simulate.run(d, ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert out["reports"][0]["converged"] == 1
    # the binary grid file the run writes (write_grb): exactly what the reference test reads back
    from modflow6_b200.output import read_grb
    grb = read_grb(tmp_path / "m.dis.grb")
    assert grb["GRID"] == grid.upper() and grb["NJA"] == 126
    assert np.array_equal(grb["IA"][0:4], [1, 4, 4, 7]) and grb["IA"][-1] == 127 and grb["IA"].shape[0] == 28
    assert np.array_equal(grb["JA"][:6], [1, 4, 10, 3, 6, 12]) and grb["JA"].shape[0] == 126
    assert np.array_equal(grb["IDOMAIN"], [1, 0] + 25 * [1]) and grb["ICELLTYPE"].shape[0] == 27
    if grid == "disv":
        assert grb["NCPL"] == 9 and grb["NVERT"] == 16 and grb["IAVERT"][-1] == grb["NJAVERT"] + 1 == 46
        assert grb["VERTICES"].shape[0] == 32 and grb["VERTICES"].min() >= 1.0e8
    raw = open(tmp_path / "m.dis.grb", "rb").read()
    assert raw[:50] == (f"GRID {grid.upper()}".ljust(49) + "\n").encode() and raw[50:100].startswith(b"VERSION 1")


def test_listing_budget_of_two_models_with_an_exchange(tmp_path):
    """the model listing files of a two-model simulation (par_gwf01): each model's VOLUME BUDGET carries its CHD
    package and the GWF-GWF exchange as a FLOW-JA-FACE entry (gwf_gwf_bd), what leaves the left model through the
    exchange enters the right one, and each budget closes"""
    from modflow6_b200.output import read_listing_budgets
    mf6_inputs.write_par_gwf01(str(tmp_path), (1, 5, 5))
    for m in ("leftmodel", "rightmodel"):
        p = tmp_path / f"{m}.oc"
        text = p.read_text()
        assert "PRINT" not in text.upper() or "PRINT BUDGET" not in text.upper()
        p.write_text(text.replace("END period 1", "  PRINT BUDGET ALL\nEND period 1").replace("END period  1", "  PRINT BUDGET ALL\nEND period  1"))
    out = simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert all(r["converged"] for r in out["reports"])
    left, right = (read_listing_budgets(tmp_path / f"{m}.lst")[-1] for m in ("leftmodel", "rightmodel"))
    ex_l = [k for k in left["rates_out"] if k[0] == "FLOW-JA-FACE"]
    ex_r = [k for k in right["rates_in"] if k[0] == "FLOW-JA-FACE"]
    assert len(ex_l) == 1 and ex_l == ex_r and ex_l[0][1].startswith("GWF-GWF")
    # heads fall from the right model's CHD (10) to the left model's (1): water crosses the exchange right -> left
    q = right["rates_out"][ex_r[0]]
    assert q > 0 and np.isclose(left["rates_in"][ex_l[0]], q, rtol=1e-9)
    assert np.isclose(left["total_in"], left["total_out"], rtol=1e-6) and abs(left["pdiffr"]) < 0.01
    assert any(k[0] == "CHD" for k in left["rates_out"]) and any(k[0] == "CHD" for k in right["rates_in"])


def test_time_series_that_ends_early_is_an_input_error(tmp_path):
    """TimeSeries.f90 get_value_at_time / get_integrated_value stop the run when a step reaches past the last record
    of a series (no extension to the end of the simulation for stress packages)"""
    write_auxmult(str(tmp_path), 0)
    p = tmp_path / "m.wel.ts"
    rows = p.read_text().split("\n")
    p.write_text("\n".join(r for r in rows if not r.strip().startswith(("3.0", "4.0"))))
    with pytest.raises(mf6io.Mf6InputError, match="period 3 step 1"):
        simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class(), write_output=False)


def test_convergence_failure_stops_the_simulation_unless_continue(tmp_path):
    """converge_check (Sim.f90:401-433): a time step that does not converge ends the simulation after its output --
    unless mfsim.nam says CONTINUE"""
    d = str(tmp_path)
    mf6_inputs.write_gwf(d, "m", (1, 1, 10), 1.0, 1.0, 10.0, [0.0], 1.0, icelltype=1, strt=10.0,
                         chd={1: [((1, 1, 1), 10.0), ((1, 1, 10), 5.0)]})
    ims = ("BEGIN nonlinear\n  OUTER_DVCLOSE 1e-9\n  OUTER_MAXIMUM 1\nEND nonlinear\n\n"
           "BEGIN linear\n  INNER_MAXIMUM 50\n  INNER_DVCLOSE 1e-10\n  INNER_RCLOSE 1e-8\nEND linear\n")
    mf6_inputs.write_sim(d, ["m"], [(3.0, 3, 1.0)], ims)
    out = simulate.run(d, ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert len(out["reports"]) == 1 and out["reports"][0]["converged"] == 0
    assert len(read_head_file(tmp_path / "m.hds")) == 1                  # the failed step's output is written
    nam = tmp_path / "mfsim.nam"
    nam.write_text(nam.read_text().replace("BEGIN options\n", "BEGIN options\n  CONTINUE\n", 1))
    out = simulate.run(d, ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    assert len(out["reports"]) == 3 and out["simulation"].continue_


def test_budget_csv_file(tmp_path):
    """OC BUDGETCSV FILEOUT: header and one row per time step in the layout of Budget.f90 writecsv / write_csv_header"""
    write_auxmult(str(tmp_path), 0)
    p = tmp_path / "m.oc"
    p.write_text(p.read_text().replace("  BUDGET FILEOUT m.cbc\n", "  BUDGET FILEOUT m.cbc\n  BUDGETCSV FILEOUT m.bud.csv\n"))
    simulate.run(str(tmp_path), ordering=T.ORDER_NATURAL, solution_class=oracle_class())
    rows = (tmp_path / "m.bud.csv").read_text().strip().split("\n")
    assert rows[0] == ("time,CHD(CHD_0)_IN,WEL(WEL_0)_IN,CHD(CHD_0)_OUT,WEL(WEL_0)_OUT,"
                       "TOTAL_IN,TOTAL_OUT,PERCENT_DIFFERENCE")
    data = np.array([[float(v) for v in r.split(",")] for r in rows[1:]])
    assert data.shape == (13, 8) and np.allclose(data[:, 0], [0.1 * i for i in range(1, 11)] + [2.0, 3.0, 4.0])
    assert np.allclose(data[:, 2], np.array(7 * [1.0, 0.0])[:-1])           # the well injects 1, 0, 1, ...
    on = data[:, 2] > 0.5                                                   # (idle steps: in and out are round-off)
    assert np.allclose(data[:, 5], data[:, 6], atol=1e-6) and np.abs(data[on, 7]).max() < 1e-3


def test_time_series_average_equals_quadrature():
    """STEPWISE / LINEAR values over a step are time-weighted averages: compared with a fine midpoint quadrature of
    the pointwise interpolant (get_value_at_time) on seeded random series and intervals"""
    from modflow6_b200.timeseries import TimeSeries
    rng = np.random.default_rng(7)
    for method in ("STEPWISE", "LINEAR"):
        for _ in range(20):
            t = np.cumsum(rng.uniform(0.1, 2.0, 8))
            s = TimeSeries("x", method, t, rng.normal(size=8))
            a, b = np.sort(rng.uniform(t[0], t[-1], 2))
            if b - a < 1e-3:
                continue
            mid = a + (np.arange(20000) + 0.5) * (b - a) / 20000
            quad = np.mean([s.value_at(x) for x in mid])
            assert abs(s.value(a, b) - quad) < 2e-3, (method, a, b)
