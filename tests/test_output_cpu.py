"""Binary .hds / .cbc writers (modflow6_b200/output.py) against hand-packed records of the reference's
stream-access formats (InputOutput.f90:924-1071), round trips through the readers (HeadFileReader.f90,
BudgetFileReader.f90), and a water-balance property of the records written from an oracle run."""
import struct

import numpy as np

from modflow6_b200 import configs
from modflow6_b200 import ctypes_types as T
from modflow6_b200.grid import tdis_steps
from modflow6_b200.output import BudgetFileWriter, HeadFileWriter, read_budget_file, read_head_file


def test_head_record_bytes(tmp_path):
    """ulasav: kstp, kper, pertim, totim, text(16), ncol, nrow, ilay, then ncol*nrow doubles -- per layer"""
    p = tmp_path / "m.hds"
    w = HeadFileWriter(p, (2, 2, 3))
    h = np.arange(12, dtype=float) + 0.5
    w.write(3, 2, 1.5, 11.5, h)
    w.close()
    want = b""
    for k in range(2):
        want += struct.pack("<iidd", 3, 2, 1.5, 11.5) + b"HEAD            " + struct.pack("<iii", 3, 2, k + 1)
        want += struct.pack("<6d", *h[6 * k:6 * k + 6])
    assert p.read_bytes() == want
    recs = read_head_file(p)
    assert [r["ilay"] for r in recs] == [1, 2] and recs[1]["text"] == "HEAD            "
    assert np.array_equal(np.concatenate([r["data"].ravel() for r in recs]), h)


def test_budget_record_bytes(tmp_path):
    p = tmp_path / "m.cbc"
    w = BudgetFileWriter(p, (1, 2, 2), "gwf_1")
    w.write_flowja(1, 1, 2.0, 2.0, 2.0, [0.0, 1.0, -1.0])
    w.write_array(1, 1, 2.0, 2.0, 2.0, "STO-SS", [1.0, 2.0, 3.0, 4.0])
    w.write_list(1, 1, 2.0, 2.0, 2.0, "WEL", "wel-1", np.array([3, 0]), np.array([-5.0, -7.0]))
    w.close()
    want = (struct.pack("<ii", 1, 1) + b"    FLOW-JA-FACE" + struct.pack("<iii", 3, 1, -1)
            + struct.pack("<iddd", 1, 2.0, 2.0, 2.0) + struct.pack("<3d", 0.0, 1.0, -1.0))
    want += (struct.pack("<ii", 1, 1) + b"          STO-SS" + struct.pack("<iii", 2, 2, -1)
             + struct.pack("<iddd", 1, 2.0, 2.0, 2.0) + struct.pack("<4d", 1.0, 2.0, 3.0, 4.0))
    want += (struct.pack("<ii", 1, 1) + b"             WEL" + struct.pack("<iii", 2, 2, -1)
             + struct.pack("<iddd", 6, 2.0, 2.0, 2.0)
             + b"GWF_1           " + b"WEL-1           " + b"GWF_1           " + b"WEL-1           "
             + struct.pack("<i", 1) + struct.pack("<i", 2)
             + struct.pack("<iid", 4, 1, -5.0) + struct.pack("<iid", 1, 2, -7.0))
    assert p.read_bytes() == want
    recs = read_budget_file(p)
    assert [r["text"].strip() for r in recs] == ["FLOW-JA-FACE", "STO-SS", "WEL"]
    assert recs[0]["flow"].size == 3 and recs[1]["flow"].size == 4
    assert recs[2]["node"].tolist() == [4, 1] and recs[2]["node2"].tolist() == [1, 2]
    assert recs[2]["srcpackage"] == "WEL-1           "


def run_and_write(solution, cfg, tmp_path, tag, max_steps=None):
    """the Mf6DoTimestep loop of configs.run_simulation with gwf_ot_dv / gwf_ot_flow after every step"""
    m = cfg.model
    shape = getattr(m, "shape", None) or (m.nodes,)
    hw = HeadFileWriter(tmp_path / f"{tag}.hds", shape)
    bw = BudgetFileWriter(tmp_path / f"{tag}.cbc", shape, "model")
    totim, nsteps = 0.0, 0
    for kper, per in enumerate(cfg.periods, start=1):
        solution.set_packages(per.packages)
        pertim = 0.0
        for kstp, delt in enumerate(tdis_steps(per.perlen, per.nstp, per.tsmult), start=1):
            solution.timestep(kper, kstp, delt, 1 if per.steady else 0)
            pertim += delt
            totim += delt
            hw.write(kstp, kper, pertim, totim, solution.x)
            bw.write_step(kstp, kper, delt, pertim, totim, solution, per.packages)
            nsteps += 1
            if max_steps and nsteps >= max_steps:
                hw.close()
                bw.close()
                return
    hw.close()
    bw.close()


def check_water_balance(cfg, hds, cbc, rtol=1e-6):
    """per cell and time step: sum of FLOW-JA-FACE over the row + storage + package rates == 0 up to the
    closure of the solve (what `zonbud`-style post-processors assume about these files)"""
    m = cfg.model
    steps = {}
    for r in cbc:
        steps.setdefault((r["kper"], r["kstp"]), []).append(r)
    assert len(steps) * 1 == len({(r["kper"], r["kstp"]) for r in hds})
    for key, recs in steps.items():
        resid = np.zeros(m.nodes)
        scale = 0.0
        chd_cells = np.zeros(m.nodes, bool)
        for r in recs:
            t = r["text"].strip()
            if t == "FLOW-JA-FACE":
                assert r["flow"].size == m.nja
                resid += np.add.reduceat(r["flow"], m.ia[:-1])
                scale = max(scale, np.abs(r["flow"]).max())
            elif r["imeth"] == 1:
                assert r["flow"].size == m.nodes
                resid += r["flow"]
            else:
                np.add.at(resid, r["node"] - 1, r["q"])
                if t == "CHD":
                    chd_cells[r["node"] - 1] = True
        assert np.abs(resid[~chd_cells]).max() <= rtol * max(scale, 1.0) + 0.02   # rclose-sized residual
        assert np.abs(resid[chd_cells]).max() <= 1e-6 * max(scale, 1.0)           # CHD rate closes its cell exactly


def test_oracle_run_writes_consistent_files(tmp_path):
    from oracle.oracle import OracleSolution
    cfg = configs.c1_npf01("a", T.ORDER_NATURAL)          # unconfined: STO-SS and STO-SY, CHD, WEL
    O = OracleSolution(cfg.model, cfg.sln, cfg.ims)
    run_and_write(O, cfg, tmp_path, "c1a")
    hds = read_head_file(tmp_path / "c1a.hds")
    cbc = read_budget_file(tmp_path / "c1a.cbc")
    assert len(hds) == 12 and all(r["text"].strip() == "HEAD" for r in hds)       # 1 + 10 + 1 steps, 1 layer
    assert np.isclose(hds[-1]["totim"], 1002.0) and hds[1]["kper"] == 2
    texts = [r["text"].strip() for r in cbc if (r["kper"], r["kstp"]) == (2, 1)]
    assert texts == ["STO-SS", "STO-SY", "FLOW-JA-FACE", "CHD", "WEL"]
    assert np.allclose(hds[-1]["data"].ravel(), O.x)
    check_water_balance(cfg, hds, cbc)


def test_binary_grid_file_of_a_dis_grid(tmp_path):
    """write_grb (Dis.f90:547-659): header lines, the sixteen variables in the reference's order, DELR / DELC / TOP /
    BOTM and the user-numbered connectivity of a DIS grid with a removed and a pass-through cell"""
    import numpy as np
    from modflow6_b200 import mf6io
    from modflow6_b200.output import read_grb, write_grb
    from tests import mf6_inputs
    idom = np.ones((2, 3, 4), dtype=int)
    idom[0, 1, 1] = 0
    idom[0, 2, 2] = -1
    mf6_inputs.write_gwf(str(tmp_path), "m", (2, 3, 4), [10.0, 20.0, 30.0, 40.0], [5.0, 6.0, 7.0], 3.0, [-1.0, -4.0], 1.0,
                         chd={1: [((1, 1, 1), 1.0)]}, strt=0.0, idomain=idom)
    mf6_inputs.write_sim(str(tmp_path), ["m"], [(1.0, 1, 1.0)], "BEGIN options\nEND options\n")
    gi = mf6io.read_simulation(str(tmp_path)).models[0]
    write_grb(tmp_path / "g.grb", gi.grid, gi.model, gi.nodeuser)
    g = read_grb(tmp_path / "g.grb")
    assert list(g)[:1] == ["GRID"] and g["GRID"] == "DIS"
    assert list(g)[1:] == ["NCELLS", "NLAY", "NROW", "NCOL", "NJA", "XORIGIN", "YORIGIN", "ANGROT", "DELR", "DELC", "TOP",
                           "BOTM", "IA", "JA", "IDOMAIN", "ICELLTYPE"]
    assert (g["NCELLS"], g["NLAY"], g["NROW"], g["NCOL"]) == (24, 2, 3, 4) and g["NJA"] == gi.model.nja
    assert np.array_equal(g["DELR"], [10, 20, 30, 40]) and np.array_equal(g["DELC"], [5, 6, 7])
    assert g["TOP"].size == 12 and g["BOTM"].size == 24 and np.array_equal(g["IDOMAIN"], idom.reshape(-1))
    ia, ja = g["IA"], g["JA"]
    assert ia[0] == 1 and ia[-1] == gi.model.nja + 1 and ja.size == gi.model.nja
    assert ia[5] == ia[6] and ia[10] == ia[11]                       # the removed / pass-through cells: empty rows
    row = lambda n: ja[ia[n - 1] - 1:ia[n] - 1].tolist()             # noqa: E731
    assert row(1) == [1, 2, 5, 13]                                   # itself, right, front, below
    assert row(23)[0] == 23 and 11 not in row(23)                    # below the pass-through cell: nothing above
    assert row(18)[0] == 18 and 6 not in row(18)                     # below the hole


def test_listing_file_budget_table_and_time_summary(tmp_path):
    """ListingFileWriter against the formats of budget_ot (Budget.f90:292-310) and tdis_ot (tdis.f90:283-297): fixed
    columns, F17.4 / 1PE17.4 switching of value_to_string, cumulative volumes = rate x delt summed, 1P G12.5 times"""
    from modflow6_b200.output import ListingFileWriter, fortran_g, read_listing_budgets
    w = ListingFileWriter(tmp_path / "m.lst", "m", "DAYS")
    e = [("STO-SS", 0.0, 0.05, "STORAGE"), ("WEL", 0.0, 2500.0, "WEL_0"), ("CHD", 2500.05, 1.0e12, "CHD_0")]
    w.write_budget(1, 1, 0.5, 0.5, 0.5, e)
    w.write_budget(2, 1, 1.0, 1.5, 1.5, e)
    w.close()
    lines = open(tmp_path / "m.lst").read().split("\n")
    hdr = [i for i, ln in enumerate(lines) if "VOLUME BUDGET" in ln]
    assert lines[hdr[0]] == "  VOLUME BUDGET FOR ENTIRE MODEL AT END OF TIME STEP    1, STRESS PERIOD   1"
    assert lines[hdr[0] + 1] == "  " + 99 * "-"
    assert lines[hdr[0] + 3].startswith("     CUMULATIVE VOLUME      L**3       RATES FOR THIS TIME STEP      L**3/T")
    row = [ln for ln in lines if ln.startswith("                 WEL =")]
    assert row[1] == "                 WEL =        1250.0000                   WEL =        2500.0000     WEL_0"
    assert "          STO-SS =       2.5000E-02" in "\n".join(lines)            # below 0.1: exponent form
    assert "             CHD =       1.0000E+12" in "\n".join(lines)            # above 9.99999e11 too
    b = read_listing_budgets(tmp_path / "m.lst")
    assert [x["kstp"] for x in b] == [1, 2]
    assert b[1]["volumes_out"][("WEL", "WEL_0")] == 2500.0 * 1.5 and b[1]["rates_in"][("CHD", "CHD_0")] == 2500.05
    assert b[0]["totim_seconds"] == 43200.0 and b[1]["totim_seconds"] == 129600.0
    assert " STRESS PERIOD TIME 1.29600E+05  2160.0      36.000      1.5000     4.10678E-03" in lines
    assert "         TIME SUMMARY AT END OF TIME STEP    2 IN STRESS PERIOD    1" in lines
    assert [fortran_g(v, 12, 5) for v in (86400.0, 24.0, 0.0)] == ["  86400.    ", "  24.000    ", "  0.0000    "]
