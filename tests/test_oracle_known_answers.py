"""Pins the CPU oracle to the reference's own known-answer tests and to
invariants of the formulation.  Runs on CPU."""
import os

import numpy as np
import pytest

from modflow6_b200 import configs
from modflow6_b200 import ctypes_types as T
from modflow6_b200.grid import Package, build_dis_model, dis_connectivity, merge_models, tdis_steps
from oracle.oracle import OracleIlu0, OracleIms, OracleSolution, amux
from tests.helpers import assembled_system, chd_west_east, hetero_dis, permute_csr, well_center


def test_chd01_linear_head_profile():
    """autotest/test_gwf_chd01.py:12-60,126-127 -- 1x1x100, K=1, CHD 1/0, CG + relax 1.0 => linspace(1,0,100)"""
    m = build_dis_model(1, 1, 100, 1.0, 1.0, top=1.0, botm=[0.0], k11=1.0, k33=1.0, icelltype=0, strt=1.0)
    ims = T.ImsSettings.make(dvclose=1e-6, rclose=1e-6, iter1=300, ilinmeth=1, relax=1.0)
    sln = T.SlnSettings.make(dvclose=1e-6, mxiter=100)
    S = OracleSolution(m, sln, ims)
    S.set_packages([Package(T.PKG_CHD, [0, 99], [1.0, 0.0])])
    rep = S.timestep(1, 1, 5.0, 1)
    assert rep.converged == 1
    assert np.allclose(S.x, np.linspace(1, 0, 100))
    assert abs(rep.pdiffr) < 1e-6


@pytest.mark.parametrize("shape", [(1, 1, 10), (1, 5, 10), (5, 5, 10)])
@pytest.mark.parametrize("meth,relax", [(2, 0.97), (1, 0.0)])
def test_par_gwf01_heads_one_to_ten(shape, meth, relax):
    """autotest/test_par_gwf01.py:20-29,200-212 (and test_par_petsc01.py) -- two 5-column models joined by a
    GWF-GWF exchange == one 10-column model: CHD 1 / 10 on the outer columns => heads 1..10 (6 decimals)"""
    nlay, nrow, ncol = shape
    m = build_dis_model(nlay, nrow, ncol, 100.0, 100.0, 0.0, -10.0 * np.arange(1, nlay + 1), 1.0, strt=1.0)
    chd = chd_west_east(m, 1.0, 10.0)
    ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-6, iter1=100, ilinmeth=meth, relax=relax)
    S = OracleSolution(m, T.SlnSettings.make(dvclose=1e-9, mxiter=50), ims)
    S.set_packages([chd])
    assert S.timestep().converged == 1
    h = S.x.reshape(shape)
    assert np.allclose(h, np.arange(1.0, 11.0)[None, None, :], atol=1e-6)


def test_newton01_perched_recharge():
    """autotest/test_gwf_newton01.py:8-21,95-103 -- NEWTON, RCH 1.0, CHD 7 in layer 2 => H1 = 8, H2 = 7
    (the COMPLEX preset's ILUT + backtracking are replaced by BICGSTAB + ILU0 + DBD: the answer is analytic)"""
    m = build_dis_model(2, 3, 3, 1.0, 1.0, top=20.0, botm=[10.0, 0.0], k11=10.0, icelltype=1, strt=7.0, inewton=1)
    ims = T.ImsSettings.make(dvclose=1e-9, rclose=0.1, iter1=500, ilinmeth=2, relax=0.0, north=2)
    sln = T.SlnSettings.make(dvclose=1e-9, mxiter=100, nonmeth=3, theta=0.8, akappa=1e-4)
    S = OracleSolution(m, sln, ims)
    S.set_packages([Package(T.PKG_CHD, np.arange(9, 18), np.full(9, 7.0)), Package(T.PKG_RCH, np.arange(9), np.full(9, 1.0))])
    rep = S.timestep(1, 1, 1.0, 1)
    assert rep.converged == 1
    assert np.allclose(S.x[:9], 8.0) and np.allclose(S.x[9:], 7.0)
    assert abs(rep.pdiffr) < 1e-5


def test_dis_connectivity_matches_disconnections():
    """Connections.f90:510-688: row layout, symmetric numbering, geometry"""
    c = dis_connectivity(2, 3, 4)
    ia, ja, jas, isym = c["ia"], c["ja"], c["jas"], c["isym"]
    n = 24
    assert ia[-1] == ja.size == n + 2 * c["njas"]
    for r in range(n):
        row = ja[ia[r]:ia[r + 1]]
        assert row[0] == r and np.all(np.diff(row[1:]) > 0)
        for p in range(ia[r] + 1, ia[r + 1]):
            q = isym[p]
            assert ja[q] == r and ja[p] == np.searchsorted(ia, q, side="right") - 1
            assert jas[p] == jas[q]
    # upper-triangle connections are numbered consecutively in (k,i,j) order
    up = [jas[p] for r in range(n) for p in range(ia[r] + 1, ia[r + 1]) if ja[p] > r]
    assert up == list(range(c["njas"]))
    m = build_dis_model(2, 3, 4, [1.0, 2.0, 3.0, 4.0], [10.0, 20.0, 30.0], 5.0, [0.0, -7.0], 1.0)
    # first connection of cell 0 is "right": cl1 = delr0/2, cl2 = delr1/2, hwva = delc0
    assert (m.ihc[0], m.cl1[0], m.cl2[0], m.hwva[0]) == (1, 0.5, 1.0, 10.0)
    # second is "front": cl1 = delc0/2, cl2 = delc1/2, hwva = delr0 ; third "down": area
    assert (m.ihc[1], m.cl1[1], m.cl2[1], m.hwva[1]) == (1, 5.0, 10.0, 1.0)
    assert (m.ihc[2], m.cl1[2], m.cl2[2], m.hwva[2]) == (0, 2.5, 3.5, 10.0)


def test_condsat_closed_forms():
    """harmonic horizontal, series vertical (gwf-npf.f90:2010-2035)"""
    m = build_dis_model(2, 1, 2, 100.0, 50.0, 0.0, [-10.0, -30.0], np.array([[[2.0, 8.0]], [[1.0, 1.0]]]),
                        k33=np.array([[[0.2, 0.8]], [[0.1, 0.1]]]))
    S = OracleSolution(m, T.SlnSettings.make(), T.ImsSettings.make())
    cs = S.condsat
    t1, t2 = 2.0 * 10.0, 8.0 * 10.0
    assert np.isclose(cs[0], 50.0 * t1 * t2 / (t1 * 50.0 + t2 * 50.0))
    assert np.isclose(cs[1], 100.0 * 50.0 / (0.5 * 10.0 / 0.2 + 0.5 * 20.0 / 0.1))


def test_assembled_matrix_invariants():
    """non-Newton A is symmetric; interior confined rows sum to zero; CHD rows are identity"""
    m = hetero_dis(3, 12, 14, seed=5)
    chd = chd_west_east(m)
    a, b, x = assembled_system(m, [chd, well_center(m)], ilinmeth=2)   # BICGSTAB: no symmetric elimination
    ia, ja = m.ia, m.ja
    dense = np.zeros((m.nodes, m.nodes))
    for r in range(m.nodes):
        dense[r, ja[ia[r]:ia[r + 1]]] = a[ia[r]:ia[r + 1]]
    isch = np.zeros(m.nodes, bool)
    isch[chd.nodelist] = True
    free = ~isch
    sub = dense[np.ix_(free, free)]
    assert np.allclose(sub, sub.T)
    rowsum = dense[free].sum(axis=1)
    assert np.abs(rowsum).max() < 1e-9 * np.abs(dense).max()
    assert np.allclose(dense[isch], np.eye(m.nodes)[isch])
    assert np.allclose(b[isch], x[isch])


def test_ilu0_is_exact_on_tridiagonal_and_reordering_consistent():
    m = hetero_dis(1, 1, 50, seed=2)
    a, b, x0 = assembled_system(m, [chd_west_east(m)])
    P = OracleIlu0(m.ia, m.ja)
    assert P.factor(a, 0.0) == 0
    r = np.random.default_rng(0).normal(size=m.nodes)
    z = P.apply(r)
    assert np.allclose(amux(m.ia, m.ja, a, z), r)
    # a symmetric permutation of the system gives the same solution
    m2 = hetero_dis(2, 9, 11, seed=4)
    a, b, x0 = assembled_system(m2, [chd_west_east(m2), well_center(m2)])
    ims = T.ImsSettings.make(dvclose=1e-10, rclose=1e-8, iter1=500)
    perm = np.random.default_rng(1).permutation(m2.nodes).astype(np.int32)
    xa, xb = x0.copy(), x0.copy()
    ita, cva = OracleIms(m2.ia, m2.ja, ims).solve(a, xa, b)
    itb, cvb = OracleIms(m2.ia, m2.ja, ims, perm=perm).solve(a, xb, b)
    assert cva == 1 and cvb == 1
    assert np.abs(xa - xb).max() < 1e-8
    ia2, ja2, a2 = permute_csr(m2.ia, m2.ja, a, perm)
    assert np.allclose(amux(ia2, ja2, a2, xa[perm]), amux(m2.ia, m2.ja, a, xa)[perm])


def test_testcnvg_options_and_epfact():
    """ImsLinearBase.f90:1101-1146, 1316-1333 through full solves with every ICNVGOPT"""
    m = hetero_dis(2, 10, 10, seed=7)
    a, b, x0 = assembled_system(m, [chd_west_east(m), well_center(m)])
    its = {}
    for opt in range(5):
        ims = T.ImsSettings.make(dvclose=1e-7, rclose=1e-3, iter1=400, icnvgopt=opt)
        x = x0.copy()
        its[opt] = OracleIms(m.ia, m.ja, ims).solve(a, x, b)
    assert its[0][1] == 1
    assert its[1] == (its[0][0], 0)  # STRICT: same iterations, "converged" only if the first inner iteration is
    # L2NORM options leave early with ICNVG = -1 -> 0 once the residual dropped by EPFACT (0.01 at kstp 1)
    for opt in (2, 3, 4):
        assert its[opt][0] < its[0][0]
    assert its[2][0] <= its[4][0]    # OR-criterion stops no later than the AND-criterion


def test_tdis_step_lengths():
    """src/Timing/tdis.f90:255-267"""
    d = tdis_steps(1000.0, 10, 1.5)
    assert np.isclose(sum(d), 1000.0) and np.isclose(d[1] / d[0], 1.5)
    assert tdis_steps(5.0, 1, 1.0) == [5.0]


def test_c1_npf01_runs_like_the_reference_case():
    """autotest/test_gwf_npf01_75x75.py (BASELINE config 1): all 12 time steps converge, budget closes,
    heads stay between the two constant heads except for the pumping drawdown cone."""
    for case in ("a", "b"):
        cfg = configs.c1_npf01(case)
        O = OracleSolution(cfg.model, cfg.sln, cfg.ims)
        reps = configs.run_simulation(O, cfg, collect_heads=True)
        assert len(reps) == 12 and all(r["converged"] == 1 for r in reps)
        assert max(abs(r["pdiffr"]) for r in reps) < 1e-2     # rclose = 0.01 on ~1e5 m3/d of flow
        h = O.x.reshape(75, 75)
        assert np.allclose(h[:, 0], 48.0) and np.allclose(h[:, -1], 40.0)
        wellnode = 38 * 75 + 38
        assert reps[-1]["head"][wellnode] < reps[0]["head"][wellnode]  # drawdown once the well is on


def test_packages_riv_ghb_drn_budget_closes():
    m = hetero_dis(2, 12, 12, seed=9, strt=10.0, top=20.0)
    rng = np.random.default_rng(3)
    riv = Package(T.PKG_RIV, rng.choice(144, 10, replace=False), np.full(10, 12.0), np.full(10, 50.0), np.full(10, 8.0))
    ghb = Package(T.PKG_GHB, [5, 50, 100], [9.0, 9.5, 10.5], [20.0, 20.0, 20.0])
    drn = Package(T.PKG_DRN, [20, 21, 22], [9.0, 9.0, 30.0], [40.0, 40.0, 40.0])
    rch = Package(T.PKG_RCH, np.arange(144), np.full(144, 1e-4))
    ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-6, iter1=300, ilinmeth=2)
    S = OracleSolution(m, T.SlnSettings.make(dvclose=1e-8, mxiter=50), ims)
    S.set_packages([riv, ghb, drn, rch])
    rep = S.timestep()
    assert rep.converged == 1 and abs(rep.pdiffr) < 1e-4
    d = rep.as_dict()["terms"]
    assert len(d) == 4


def _two_model_case(nlay, nrow):
    """autotest/test_par_gwf01.py:20-120 literally: two nlay x nrow x 5 models side by side, joined by a GWF-GWF
    exchange between column 5 of the left and column 1 of the right model, CHD 1 on the far left, 10 on the far
    right"""
    mk = lambda strt: build_dis_model(nlay, nrow, 5, 100.0, 100.0, 0.0, -10.0 * np.arange(1, nlay + 1), 1.0, strt=strt)  # noqa: E731
    left, right = mk(1.0), mk(10.0)
    kk, ii = np.meshgrid(np.arange(nlay), np.arange(nrow), indexing="ij")
    n1 = ((kk * nrow + ii) * 5 + 4).reshape(-1)
    n2 = ((kk * nrow + ii) * 5 + 0).reshape(-1)
    ex = dict(m1=0, m2=1, nodem1=n1, nodem2=n2, ihc=np.ones(n1.size, np.int32), cl1=np.full(n1.size, 50.0),
              cl2=np.full(n1.size, 50.0), hwva=np.full(n1.size, 100.0))
    merged, offs = merge_models([left, right], [ex])
    chd = Package(T.PKG_CHD, np.concatenate([n2 + offs[0], n1 + offs[1]]),
                  np.concatenate([np.full(n2.size, 1.0), np.full(n1.size, 10.0)]))
    return merged, offs, chd


@pytest.mark.parametrize("shape", [(1, 1), (1, 5), (5, 5)])
def test_par_gwf01_two_models_and_exchange(shape):
    """the same answer through the serial multi-model path (models merged into one solution matrix)"""
    merged, offs, chd = _two_model_case(*shape)
    ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-6, iter1=100, ilinmeth=2, relax=0.97)
    S = OracleSolution(merged, T.SlnSettings.make(dvclose=1e-9, mxiter=50), ims)
    S.set_packages([chd])
    assert S.timestep().converged == 1
    h = S.x
    left = h[:offs[1]].reshape(shape[0], shape[1], 5)
    right = h[offs[1]:].reshape(shape[0], shape[1], 5)
    assert np.allclose(left, np.arange(1.0, 6.0)[None, None, :], atol=1e-6)
    assert np.allclose(right, np.arange(6.0, 11.0)[None, None, :], atol=1e-6)


def test_backtracking_is_exercised():
    """sln_backtracking (NumericalSolution.f90:2680-2842): a tight BACKTRACKING_TOLERANCE forces steps"""
    cfg = configs.c1_npf01("a")
    cfg.sln = T.SlnSettings.make(dvclose=1e-6, mxiter=100, nonmeth=0, numtrack=5, btol=0.3, breduc=0.5, res_lim=1e-9)
    O = OracleSolution(cfg.model, cfg.sln, cfg.ims)
    reps = configs.run_simulation(O, cfg, max_steps=3)
    assert all(r["converged"] == 1 for r in reps) and sum(r["nbacktracks"] for r in reps) >= 2


def test_ifmod_newton_split_equals_single():
    """autotest/test_gwf_ifmod_newton.py:346-380 -- the reference's criterion for interface models: a Newton,
    convertible-cell model split into two models joined by a GWF-GWF exchange gives the heads of the single
    model (max |dh| < 10 x 1e-9 ... here 1e-8) and a budget that closes (|percent discrepancy| < 1e-5)"""
    from modflow6_b200.grid import build_dis_model, merge_models
    rng = np.random.default_rng(21)
    nlay, nrow, ncol, half = 2, 4, 10, 5
    k = np.exp(rng.normal(0.0, 0.5, (nlay, nrow, ncol)))
    opts = dict(k33=None, icelltype=1, strt=6.0, inewton=1, inewtonur=1)
    single = build_dis_model(nlay, nrow, ncol, 50.0, 50.0, 10.0, [0.0, -10.0], k, **opts)
    parts = [build_dis_model(nlay, nrow, half, 50.0, 50.0, 10.0, [0.0, -10.0], k[:, :, :half], **opts),
             build_dis_model(nlay, nrow, ncol - half, 50.0, 50.0, 10.0, [0.0, -10.0], k[:, :, half:], **opts)]
    node = lambda kk, i, j, nc: (kk * nrow + i) * nc + j   # noqa: E731
    ki = [(kk, i) for kk in range(nlay) for i in range(nrow)]
    exg = dict(m1=0, m2=1, nodem1=np.array([node(kk, i, half - 1, half) for kk, i in ki]),
               nodem2=np.array([node(kk, i, 0, ncol - half) for kk, i in ki]), ihc=np.ones(len(ki), np.int32),
               cl1=np.full(len(ki), 25.0), cl2=np.full(len(ki), 25.0), hwva=np.full(len(ki), 50.0))
    merged, offs = merge_models(parts, [exg])
    # global cell (kk, i, j) -> merged numbering
    def to_merged(kk, i, j):
        return node(kk, i, j, half) if j < half else int(offs[1]) + node(kk, i, j - half, ncol - half)
    chd_cells = [(1, i, 0) for i in range(nrow)]
    rch_cells = [(0, i, j) for i in range(nrow) for j in range(ncol)]
    def pkgs(f):
        return [Package(T.PKG_CHD, [f(*c) for c in chd_cells], np.full(len(chd_cells), 2.0)),
                Package(T.PKG_RCH, [f(*c) for c in rch_cells], np.full(len(rch_cells), 2e-3))]
    sln = T.SlnSettings.make(dvclose=1e-10, mxiter=200, nonmeth=3, theta=0.9, akappa=1e-4, iallowptc=0)
    ims = T.ImsSettings.make(dvclose=1e-11, rclose=1e-9, iter1=200, ilinmeth=2, relax=0.0)
    A = OracleSolution(single, sln, ims)
    A.set_packages(pkgs(lambda kk, i, j: node(kk, i, j, ncol)))
    ra = A.timestep()
    B = OracleSolution(merged, sln, ims)
    B.set_packages(pkgs(to_merged))
    rb = B.timestep()
    assert ra.converged == 1 and rb.converged == 1
    hb = np.array([B.x[to_merged(kk, i, j)] for kk in range(nlay) for i in range(nrow) for j in range(ncol)])
    assert np.abs(A.x - hb).max() < 1e-8
    assert abs(ra.pdiffr) < 1e-5 and abs(rb.pdiffr) < 1e-5
    assert A.x.max() > 2.5 and A.x[: nrow * ncol].min() > 0.0   # a real water table in the convertible top layer


BUMP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ex-gwf-bump")


def bump_case():
    """autotest/test_gwf_newton_under_relaxation.py:9-101 (ex-gwf-bump): 51 x 51 convertible cells over a bumpy
    bottom, CHD 7.5 / 2.5 on the west / east columns, NEWTON UNDER_RELAXATION, IMS defaults (SIMPLE) with BICGSTAB,
    OUTER_DVCLOSE 1e-8, OUTER_MAXIMUM 75, INNER 100 / 1e-9 / 1e-3, NO_PTC ALL"""
    from modflow6_b200.grid import build_dis_model
    botm = np.loadtxt(os.path.join(BUMP, "bottom.txt")).reshape(1, 51, 51)
    m = build_dis_model(1, 51, 51, 100.0 / 51, 100.0 / 51, 25.0, botm, 1.0, icelltype=1, strt=7.5, inewton=1,
                        inewtonur=1)
    chd = Package(T.PKG_CHD, [i * 51 for i in range(51)] + [i * 51 + 50 for i in range(51)], [7.5] * 51 + [2.5] * 51)
    sln = T.SlnSettings.make(dvclose=1e-8, mxiter=75, nonmeth=0, theta=1.0, akappa=0.0, gamma=1.0, amomentum=0.0,
                             iallowptc=0, numtrack=0, btol=0.0, breduc=0.0, res_lim=0.0)
    ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-3, iter1=100, ilinmeth=2, relax=0.0)
    return m, chd, sln, ims


def test_ex_gwf_bump_reference_heads():
    """the head file MODFLOW 6 itself wrote for this model (tests/golden/README.md): Newton-Raphson terms
    (npf_fn), Newton under-relaxation (npf_nur), quadratic saturation smoothing, BiCGSTAB + ILU0 -- the oracle
    lands on the reference's heads (max |dh| 2e-8), the bar of the reference test being np.allclose"""
    from modflow6_b200.output import read_head_file
    base = read_head_file(os.path.join(BUMP, "results.hds.cmp"))
    assert len(base) == 1 and base[0]["text"] == "HEAD            " and (base[0]["nrow"], base[0]["ncol"]) == (51, 51)
    m, chd, sln, ims = bump_case()
    O = OracleSolution(m, sln, ims)
    O.set_packages([chd])
    rep = O.timestep(1, 1, 1.0, 1)
    assert rep.converged == 1
    assert np.allclose(base[0]["data"], O.x.reshape(51, 51))
    assert np.abs(base[0]["data"] - O.x.reshape(51, 51)).max() < 1e-6


@pytest.mark.parametrize("newton", [False, True])
def test_drn_ddrn01_discharge_scaling(newton):
    """autotest/test_gwf_drn_ddrn01.py:120-190 -- the reference's own criterion: the simulated drain discharge of
    every time step equals the analytic scaling (linear for Picard, cubic under NEWTON) of the drain cell's head
    to 1e-6 (drainage depth, get_drain_factor and drn_fn, gwf-drn.f90:420-574)"""
    from modflow6_b200 import configs
    from tests.helpers import drn_ddrn01_case
    cfg, analytic = drn_ddrn01_case(newton)
    S = OracleSolution(cfg.model, cfg.sln, cfg.ims)
    S.set_packages(cfg.periods[0].packages)
    from modflow6_b200.grid import tdis_steps
    heads, q = [], []
    for kstp, delt in enumerate(tdis_steps(100.0, 100, 1.1), start=1):
        rep = S.timestep(1, kstp, delt, 0)
        assert rep.converged == 1
        heads.append(S.x[-1])
        q.append(S.simvals[0][0])
    heads, q = np.array(heads), np.array(q)
    assert np.abs(q - analytic(heads)).max() < 1e-6
    assert q.min() < -1e-3 and 0.0 < heads[-1] < 1.0    # the drain works inside its scaling range


NPF05_ANSWER = np.array([100.0, 100.00031999, 100.00055998, 100.00071997, 100.00079997,
                         109.99960002, 109.99964002, 109.99972002, 109.99984001, 110.0])


def npf05_model(**aniso):
    """autotest/test_gwf_npf05_anisotropy.py:16-118: 2 x 1 x 5, top 100, botm 50 / 0, K 5, K22 0.5, K33 0.05 (given
    as K22OVERK / K33OVERK ratios there), CHD 100 at (1,1,1) and 110 at (2,1,5), recharge 0.01"""
    m = build_dis_model(2, 1, 5, 1.0, 1.0, 100.0, [50.0, 0.0], 5.0, k33=0.05, icelltype=0, strt=100.0, **aniso)
    pk = [Package(T.PKG_CHD, [0, 9], [100.0, 110.0]), Package(T.PKG_RCH, np.arange(5), np.full(5, 0.01))]
    ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-3, iter1=300, ilinmeth=2, relax=0.97)
    sln = T.SlnSettings.make(dvclose=1e-9, mxiter=100, nonmeth=3, theta=0.7, akappa=0.1, gamma=0.2, amomentum=0.001)
    return m, pk, sln, ims


def test_npf05_anisotropy_literal_heads():
    """autotest/test_gwf_npf05_anisotropy.py:127-141 -- the literal head array; with one row every horizontal
    connection runs along x, so hy_eff must return K11 whatever K22 is (hyeff, HGeoUtil.f90:29-108)"""
    m, pk, sln, ims = npf05_model(k22=0.5)
    S = OracleSolution(m, sln, ims)
    S.set_packages(pk)
    assert S.timestep().converged == 1
    assert np.allclose(S.x, NPF05_ANSWER)


def _strip(axis, k11, k22=None, angle1=None, n=12):
    """confined strip of n cells along x (axis 0) or y (axis 1) between CHD 10 and 0: returns the heads"""
    shape = (1, 1, n) if axis == 0 else (1, n, 1)
    rng = np.random.default_rng(4)
    kk = k11 * np.exp(rng.normal(0.0, 0.5, size=shape))
    opts = {}
    if k22 is not None:
        opts["k22"] = k22 * np.exp(rng.normal(0.0, 0.5, size=shape))
    if angle1 is not None:
        opts["angle1"] = angle1
    m = build_dis_model(*shape, 10.0, 10.0, 0.0, [-5.0], kk, strt=5.0, **opts)
    S = OracleSolution(m, T.SlnSettings.make(dvclose=1e-10, mxiter=50),
                       T.ImsSettings.make(dvclose=1e-11, rclose=1e-9, iter1=100, ilinmeth=2))
    S.set_packages([Package(T.PKG_CHD, [0, n - 1], [10.0, 0.0])])
    assert S.timestep().converged == 1
    return S.x.copy(), kk, opts.get("k22")


def test_hy_eff_directional_conductivity():
    """hy_eff (gwf-npf.f90:2280-2355): flow along y sees K22, along x K11; ANGLE1 = 90 degrees swaps them"""
    hy_ref, kk, k22 = _strip(1, 3.0, k22=0.7)
    # the same strip with K11 := that K22 field and no anisotropy
    m = build_dis_model(1, 12, 1, 10.0, 10.0, 0.0, [-5.0], k22, strt=5.0)
    S = OracleSolution(m, T.SlnSettings.make(dvclose=1e-10, mxiter=50),
                       T.ImsSettings.make(dvclose=1e-11, rclose=1e-9, iter1=100, ilinmeth=2))
    S.set_packages([Package(T.PKG_CHD, [0, 11], [10.0, 0.0])])
    S.timestep()
    assert np.array_equal(hy_ref, S.x)                    # exactly K22: the unit normal is exactly (0, -1)
    hx, _, _ = _strip(0, 3.0, k22=0.7)                    # along x K22 is invisible
    hx0, _, _ = _strip(0, 3.0)
    assert np.array_equal(hx, hx0)
    hrot, _, _ = _strip(0, 3.0, k22=0.7, angle1=np.arctan(1.0) * 2.0)   # ellipse turned by 90 degrees: x sees K22
    assert np.allclose(hrot, hy_ref, rtol=0, atol=1e-9)


def test_ilut_known_answers():
    """sparskit2/ilut.f90:48-548 restated (oracle/ilut.c), pinned by what the algorithm must do by construction:
    (1) with no dropping (droptol 0, lfil = n) ILUT is the complete LU factorisation: lusol solves exactly;
    (2) on a tridiagonal matrix there is no fill, so ILUT(lfil >= 1, 0) equals ILU0 -- values and apply;
    (3) MILUT (relax > 0) still solves the system;
    (4) the COMPLEX preset of the reference (LEVEL 5, DROPTOL 1e-4, BICGSTAB; ImsLinearSettings.f90:103-113) converges
        in far fewer iterations than ILU0 on the same system."""
    import scipy.sparse as sp
    from oracle.oracle import OracleIlu0, OracleIlut, OracleIms
    from tests.helpers import assembled_system, chd_west_east, hetero_dis, well_center
    m = hetero_dis(2, 9, 11, seed=3)
    a, b, x0 = assembled_system(m, [chd_west_east(m), well_center(m)])
    n = m.nodes
    A = sp.csr_matrix((a, m.ja, m.ia), shape=(n, n))
    r = np.random.default_rng(0).normal(size=n)
    P = OracleIlut(m.ia, m.ja, n, 0.0)
    assert P.factor(a, 0.0) == 0
    z = P.apply(r)
    assert np.abs(A @ z - r).max() <= 1e-9 * np.abs(r).max()
    # tridiagonal
    nt = 40
    ia = np.zeros(nt + 1, np.int32)
    ja, av = [], []
    rng = np.random.default_rng(1)
    for i in range(nt):
        ja.append(i)
        av.append(4.0 + rng.random())
        for j in (i - 1, i + 1):
            if 0 <= j < nt:
                ja.append(j)
                av.append(-1.0 - 0.3 * rng.random())
        ia[i + 1] = len(ja)
    ja, av = np.array(ja, np.int32), np.array(av)
    rt = rng.normal(size=nt)
    Pt, P0 = OracleIlut(ia, ja, 3, 0.0), OracleIlu0(ia, ja)
    assert Pt.factor(av, 0.0) == 0 and P0.factor(av, 0.0) == 0
    assert np.array_equal(Pt.apply(rt), P0.apply(rt))
    # MILUT (relax > 0: the dropped terms go to the pivot, ilut.f90:365) still solves the system
    sm = T.ImsSettings.make(dvclose=1e-9, rclose=1e-7, iter1=500, ilinmeth=2, level=2, droptol=1e-3, relax=0.97)
    xm = x0.copy()
    assert OracleIms(m.ia, m.ja, sm).solve(a, xm, b)[1] == 1
    assert np.abs(A @ xm - b).max() < 1e-6
    # COMPLEX preset
    s5 = T.ImsSettings.make(dvclose=1e-9, rclose=1e-7, iter1=500, ilinmeth=2, level=5, droptol=1e-4)
    s0 = T.ImsSettings.make(dvclose=1e-9, rclose=1e-7, iter1=500, ilinmeth=2)
    x5, xz = x0.copy(), x0.copy()
    it5, cv5 = OracleIms(m.ia, m.ja, s5).solve(a, x5, b)
    it0, cv0 = OracleIms(m.ia, m.ja, s0).solve(a, xz, b)
    assert cv5 == 1 and cv0 == 1 and it5 < it0
    assert np.abs(x5 - xz).max() < 1e-7


NPF02_1LAY = np.array([
    [1.000000000000000000e02, 9.491194679807708212e01, 8.963852725425633139e01, 8.415783778939784554e01,
     7.844327180838388358e01, 7.246196989134719502e01, 6.617253575516674857e01, 5.952154961171697778e01,
     5.243800230005604845e01, 4.482387907284233819e01, 3.653700567841701030e01, 2.735652405614735727e01,
     1.690257501779389671e01, 4.399393234039389533e00, -4.000000000000000000e01],
    [2.500000000000000000e01, 2.236586340188107513e01, 1.963190756499453116e01, 1.678583852445553148e01,
     1.381261073936328998e01, 1.069347358389677538e01, 7.404548318470112633e00, 3.914611143252753056e00,
     1.814550305069945468e-01, -3.854475465269350920e00, -8.282216319385574010e00, -1.324637250326845006e01,
     -1.901458294740395516e01, -2.622133464488191024e01, -4.000000000000000000e01]])
NPF02_3LAY = np.array([
    [1.000000000000000000e02, 9.496635368428880497e01, 8.974621521357816789e01, 8.432145824563308167e01,
     7.866533338822696919e01, 7.274409668492738490e01, 6.651345136714729733e01, 5.991038738691469945e01,
     5.278755954306133447e01, 4.516004619735397796e01, 3.696996838160407606e01, 2.789393025475152399e01,
     1.752424435308674333e01, 4.558574627023233461e00, -4.000000000000000000e01],
    [2.500000000000000000e01, 2.237276762221516435e01, 1.964557102731594540e01, 1.680659328791805862e01,
     1.384042482892330028e01, 1.072732402849822897e01, 7.440599125420748194e00, 3.939435701073743079e00,
     9.093214618820914807e-02, -3.940381765174159057e00, -8.354930726606033531e00, -1.330380168293016219e01,
     -1.905635937367108212e01, -2.625249959368398223e01, -4.000000000000000000e01]])


def npf02_rewet_case(nlay):
    """autotest/test_gwf_npf02_rewet.py:8-200 (cases a / c, one model): 10 rows x 15 columns, 1 or 3 convertible
    layers, K 10, strt -40, REWET WETFCT 1 IWETIT 1 IHDWET 1, WETDRY -0.001, CHD 100 (period 2: 25) on the left
    wherever the layer bottom lies below it and -40 on the right; CG + MILU0 (relax 1), closure 0.1 / 0.01.
    Returns (model, [packages of period 1, of period 2], sln, ims)."""
    nrow, ncol = 10, 15
    delr, delc = 15.0 * 500.0 / ncol, 10.0 * 500.0 / nrow
    botm = [-50.0] if nlay == 1 else [50.0, 0.0, -50.0]
    m = build_dis_model(nlay, nrow, ncol, delr, delc, 150.0, botm, 10.0, icelltype=1, strt=-40.0,
                        wetdry=-0.001, irewet=1, wetfct=1.0, iwetit=1, ihdwet=1)
    def chd(vl):
        left = [(k * nrow + i) * ncol for k in range(nlay) for i in range(nrow) if botm[k] < vl]
        right = [(k * nrow + i) * ncol + ncol - 1 for k in range(nlay) for i in range(nrow) if botm[k] < -40.0]
        return [Package(T.PKG_CHD, left, np.full(len(left), vl)), Package(T.PKG_CHD, right, np.full(len(right), -40.0))]
    ims = T.ImsSettings.make(dvclose=1e-1, rclose=0.01, iter1=100, ilinmeth=1, relax=1.0)
    sln = T.SlnSettings.make(dvclose=1e-1, mxiter=1000, nonmeth=0)
    return m, [chd(100.0), chd(25.0)], sln, ims


def npf02_profile(x, nlay):
    """the test's own reduction: the head of the highest wet layer along the middle row (:281-303)"""
    h = np.asarray(x).reshape(nlay, 10, 15)[:, 5, :]
    ht = np.full(15, 1e30)
    for k in range(nlay):
        sel = (ht == 1e30) & (h[k] != -1e30)
        ht[sel] = h[k][sel]
    return ht


@pytest.mark.parametrize("nlay", [1, 3])
def test_npf02_rewet_literal_heads(nlay):
    """autotest/test_gwf_npf02_rewet.py:203-327 -- the literal head profiles of both stress periods, to the
    reference's own tolerance 1e-9: rewetting sweep (rewet_check) + drying + npf_ad + the chd_rp reset"""
    m, periods, sln, ims = npf02_rewet_case(nlay)
    S = OracleSolution(m, sln, ims)
    want = NPF02_1LAY if nlay == 1 else NPF02_3LAY
    for kper, pk in enumerate(periods, start=1):
        S.set_packages(pk)
        rep = S.timestep(kper, 1, 1.0, 1)
        assert rep.converged == 1
        assert np.abs(npf02_profile(S.x, nlay) - want[kper - 1]).max() < 1e-9


def npf02_two_model_case(nlay):
    """cases b / d of autotest/test_gwf_npf02_rewet.py: the same grid as two models (10 + 5 columns) joined by a
    GWF-GWF exchange, in ONE solution.  Returns (merged model, [packages per period], sln, ims, offsets, ncols)."""
    from modflow6_b200.grid import merge_models
    nrow, ncols = 10, [10, 5]
    botm = [-50.0] if nlay == 1 else [50.0, 0.0, -50.0]
    ms = [build_dis_model(nlay, nrow, nc, 500.0, 500.0, 150.0, botm, 10.0, icelltype=1, strt=-40.0, wetdry=-0.001,
                          irewet=1, wetfct=1.0, iwetit=1, ihdwet=1) for nc in ncols]
    kk, ii = np.meshgrid(np.arange(nlay), np.arange(nrow), indexing="ij")
    n1 = ((kk * nrow + ii) * ncols[0] + ncols[0] - 1).reshape(-1)
    n2 = ((kk * nrow + ii) * ncols[1]).reshape(-1)
    ex = dict(m1=0, m2=1, nodem1=n1, nodem2=n2, ihc=np.ones(n1.size, np.int32), cl1=np.full(n1.size, 250.0),
              cl2=np.full(n1.size, 250.0), hwva=np.full(n1.size, 500.0))
    merged, offs = merge_models(ms, [ex])
    def chd(vl):
        left = [int(offs[0]) + (k * nrow + i) * ncols[0] for k in range(nlay) for i in range(nrow) if botm[k] < vl]
        right = [int(offs[1]) + (k * nrow + i) * ncols[1] + ncols[1] - 1 for k in range(nlay) for i in range(nrow)
                 if botm[k] < -40.0]
        return [Package(T.PKG_CHD, left, np.full(len(left), vl)), Package(T.PKG_CHD, right, np.full(len(right), -40.0))]
    ims = T.ImsSettings.make(dvclose=1e-1, rclose=0.01, iter1=100, ilinmeth=1, relax=1.0)
    sln = T.SlnSettings.make(dvclose=1e-1, mxiter=1000, nonmeth=0)
    return merged, [chd(100.0), chd(25.0)], sln, ims, offs, ncols


def npf02_two_model_profile(x, nlay, offs, ncols):
    out = []
    for j, nc in enumerate(ncols):
        h = np.asarray(x)[int(offs[j]):int(offs[j + 1])].reshape(nlay, 10, nc)[:, 5, :]
        ht = np.full(nc, 1e30)
        for k in range(nlay):
            sel = (ht == 1e30) & (h[k] != -1e30)
            ht[sel] = h[k][sel]
        out.append(ht)
    return np.concatenate(out)


@pytest.mark.parametrize("nlay", [1, 3])
def test_npf02_rewet_two_models_literal_heads(nlay):
    """cases b / d: rewetting across a GWF-GWF exchange (the exchange is an ordinary connection of the merged
    system, so `rewet_check` sees the neighbour model's cells) -- the same literal heads, 1e-9"""
    m, periods, sln, ims, offs, ncols = npf02_two_model_case(nlay)
    S = OracleSolution(m, sln, ims)
    want = NPF02_1LAY if nlay == 1 else NPF02_3LAY
    for kper, pk in enumerate(periods, start=1):
        S.set_packages(pk)
        assert S.timestep(kper, 1, 1.0, 1).converged == 1
        assert np.abs(npf02_two_model_profile(S.x, nlay, offs, ncols) - want[kper - 1]).max() < 1e-9


@pytest.mark.parametrize("idx", range(9))
def test_npf_thickstrt_hfb_literal_answers(idx):
    """autotest/test_gwf_npf_thickstrt.py:129-194 -- the reference's own assertions (np.allclose on the six heads
    and on the first CHD's inflow) for ICELLTYPE 0 / 1 / -1 x THICKSTRT x one horizontal flow barrier: prepcheck's
    treatment of a negative ICELLTYPE, calc_initial_sat (gwf-npf.f90:1838-1882, 2046-2057), condsat_modify, hfb_fc
    and hfb_cq (gwf-hfb.f90:149-450, 770-832)"""
    from tests.helpers import npf_thickstrt_case
    cfg, hfb, heads, inflow = npf_thickstrt_case(idx)
    S = OracleSolution(cfg.model, cfg.sln, cfg.ims)
    S.set_packages(cfg.periods[0].packages)
    if hfb:
        S.set_hfb(*hfb)
    S.timestep(1, 1, 1.0, 1)
    assert np.allclose(heads, S.x)
    assert np.allclose(inflow, S.simvals[0][0])


def test_gnc_restores_the_linear_field_on_a_nested_grid():
    """Ghost node correction (GhostNode.f90 gnc_fc explicit branch :280-324, gnc_cq :478-542).  The reference has no
    GNC case with literal answers in autotest/, so the pin is the property the method is built on (Panday and
    Langevin 2012): with the ghost nodes interpolated correctly a LINEAR head field satisfies the discrete equations
    of a locally refined grid exactly; without them it does not"""
    from tests.helpers import nested_grid_case
    m, chd, gnc, exact = nested_grid_case()
    ims = T.ImsSettings.make(dvclose=1e-12, rclose=1e-12, iter1=200, ilinmeth=1)
    sln = T.SlnSettings.make(dvclose=1e-10, mxiter=200)
    plain = OracleSolution(m, sln, ims)
    plain.set_packages([chd])
    assert plain.timestep(1, 1, 1.0, 1).converged == 1
    assert np.abs(plain.x - exact).max() > 1e-2             # the two-point flux is wrong across the refinement
    S = OracleSolution(m, sln, ims)
    S.set_packages([chd])
    S.set_gnc(*gnc)
    rep = S.timestep(1, 1, 1.0, 1)
    assert rep.converged == 1 and rep.outer_iterations > 2   # explicit: the correction lags one outer iteration
    assert np.abs(S.x - exact).max() < 1e-8
    # flowja carries the correction: every free cell's flows balance, and the interface flows are the exact Darcy
    # fluxes K * 0.7 * (face width 1) into the fine cells
    row = np.repeat(np.arange(m.nodes), np.diff(m.ia))
    for a, b in zip(gnc[0], gnc[1]):
        q = S.flowja[(row == b) & (m.ja == a)][0]
        assert abs(q - (-0.7)) < 1e-7 or abs(q - 0.7) < 1e-7
    assert abs(rep.pdiffr) < 1e-6


def test_gnc_on_an_unconfined_nested_grid_picard_and_newton():
    """the nested grid unconfined (water table between 10 and 20 in cells from 0 to 25), Picard (gnc_fc) and NEWTON
    (gnc_fc + gnc_fn, :340-443): both converge, the correction moves the heads by centimetres in both formulations,
    and the corrected interface flows balance (budget closes).  The two formulations discretise the saturated
    thickness differently (averaged vs upstream), so they are not compared with each other."""
    from tests.helpers import nested_grid_case
    for newton in (0, 1):
        out = {}
        for use_gnc in (True, False):
            m, chd, gnc, _ = nested_grid_case()
            m.top[:], m.icelltype[:], m.inewton = 25.0, 1, newton
            ims = T.ImsSettings.make(dvclose=1e-12, rclose=1e-10, iter1=200, ilinmeth=2)
            sln = T.SlnSettings.make(dvclose=1e-10, mxiter=300)
            S = OracleSolution(m, sln, ims)
            S.set_packages([chd])
            if use_gnc:
                S.set_gnc(*gnc)
            rep = S.timestep(1, 1, 1.0, 1)
            assert rep.converged == 1 and abs(rep.pdiffr) < 1e-6, (newton, use_gnc)
            out[use_gnc] = S.x.copy()
        assert 1e-3 < np.abs(out[True] - out[False]).max() < 0.1


@pytest.mark.parametrize("newton", [False, True])
def test_sto03_storage_is_invariant_under_an_elevation_offset(newton):
    """autotest/test_gwf_sto03.py:13-255 -- ONE convertible cell (top 0, bottom -100, ss 1e-5, sy 0) filled and
    emptied by a well in alternating stress periods (6 periods x 50 steps x 1.1), Picard (CG, SIMPLE under-relaxation
    0.95) and NEWTON UNDER_RELAXATION (BiCGSTAB).  The reference's own criteria: the heads of the same model lifted by
    15999.1 are the same once the offset is removed (atol 1e-6), the head reached at the end of every filling period
    is the same (t = 1, 3, 5), and the STO-SS rates of the two models agree -- the convertible-cell specific-storage
    formulation (SsTerms, gwf-sto.f90 / GwfStorageUtils.f90) must not depend on the datum."""
    from modflow6_b200 import configs
    from modflow6_b200.grid import tdis_steps
    ss, off = 1e-5, 15999.1
    absrate = 1.1 * ss * 100.0 * 90.0
    out = {}
    for tag, offset in (("base", 0.0), ("offset", off)):
        m = build_dis_model(1, 1, 1, 1.0, 1.0, 0.0 + offset, [-100.0 + offset], 1.0, icelltype=1,
                            strt=-100.0 + 1e-7 + offset, ss=ss, sy=0.0, iconvert=1, inewton=1 if newton else 0,
                            inewtonur=1 if newton else 0)
        ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-6, iter1=300, ilinmeth=2 if newton else 1, relax=1.0)
        sln = T.SlnSettings.make(dvclose=1e-9, mxiter=500, nonmeth=1, gamma=1.0 if newton else 0.95)
        S = OracleSolution(m, sln, ims)
        heads, rates, ends = [], [], []
        for kper in range(1, 7):
            S.set_packages([Package(T.PKG_WEL, [0], [absrate if kper % 2 == 1 else -absrate])])
            for kstp, delt in enumerate(tdis_steps(1.0, 50, 1.1), start=1):
                rep = S.timestep(kper, kstp, delt, 0)
                assert rep.converged == 1, (tag, kper, kstp)
                heads.append(S.x[0])
                rates.append(S.storage_rates[0][0])
            ends.append(S.x[0])
        out[tag] = (np.array(heads), np.array(rates), np.array(ends))
    hb, rb, eb = out["base"]
    ho, ro, eo = out["offset"]
    assert np.allclose(hb, ho - off, atol=1e-6)
    assert np.allclose(rb, ro)
    assert np.allclose(eb[[0, 2, 4]], eb[0]) and np.allclose(eo[[0, 2, 4]], eo[0])
    assert eb[0] > -100.0 + 1.0                    # the cell did fill (the test is not vacuous)
