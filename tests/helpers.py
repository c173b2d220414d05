"""Shared problem builders for the tests (seeded, sized to run in seconds)."""
import numpy as np

from modflow6_b200 import ctypes_types as T
from modflow6_b200.grid import Package, build_dis_model


def hetero_dis(nlay, nrow, ncol, seed=1, sigma=1.0, icelltype=0, top=0.0, dz=10.0, delr=100.0,
               k33_ratio=0.1, strt=44.0, **opts):
    rng = np.random.default_rng(seed)
    k = np.exp(rng.normal(np.log(10.0), sigma, size=(nlay, nrow, ncol)))
    botm = top - dz * np.arange(1, nlay + 1)
    return build_dis_model(nlay, nrow, ncol, delr, delr, top, botm, k, k33=k33_ratio * k,
                           icelltype=icelltype, strt=strt, **opts)


def chd_west_east(m, hw=48.0, he=40.0):
    nlay, nrow, ncol = m.shape
    kk, ii = np.meshgrid(np.arange(nlay), np.arange(nrow), indexing="ij")
    west = ((kk * nrow + ii) * ncol).reshape(-1)
    east = west + ncol - 1
    return Package(T.PKG_CHD, np.concatenate([west, east]),
                   np.concatenate([np.full(west.size, hw), np.full(east.size, he)]))


def well_center(m, q=-1000.0, layer=None):
    nlay, nrow, ncol = m.shape
    k = nlay // 2 if layer is None else layer
    return Package(T.PKG_WEL, [m.node(k, nrow // 2, ncol // 2)], [q])


def permute_csr(ia, ja, a, perm):
    """B = P A P^T, rows stored diagonal first then ascending (what the oracle builds)."""
    n = ia.size - 1
    iperm = np.empty(n, np.int64)
    iperm[perm] = np.arange(n)
    ia2 = np.zeros(n + 1, np.int32)
    ja2 = np.empty_like(ja)
    a2 = np.empty_like(a)
    pos = 0
    for r in range(n):
        o = perm[r]
        s, e = ia[o], ia[o + 1]
        cols = iperm[ja[s + 1:e]]
        order = np.argsort(cols, kind="stable")
        ja2[pos] = r
        a2[pos] = a[s]
        m = e - s - 1
        ja2[pos + 1:pos + 1 + m] = cols[order]
        a2[pos + 1:pos + 1 + m] = a[s + 1:e][order]
        pos += m + 1
        ia2[r + 1] = pos
    return ia2, ja2, a2


def assembled_system(m, pkgs, ilinmeth=1):
    """Formulated (amat, rhs, x) of the first outer iteration, from the oracle."""
    from oracle.oracle import OracleSolution
    ims = T.ImsSettings.make(ilinmeth=ilinmeth)
    sln = T.SlnSettings.make()
    S = OracleSolution(m, sln, ims)
    S.set_packages(pkgs)
    # apply chd_ad so that x carries the constant heads
    x = S.x
    for p in pkgs:
        if p.type == T.PKG_CHD:
            x[p.nodelist] = p.b1
    S.formulate(kiter=1, delt=1.0, iss=1)
    return S.amat.copy(), S.rhs.copy(), S.x.copy()


def drn_ddrn01_case(newton):
    """autotest/test_gwf_drn_ddrn01.py:17-118 literally: 1 x 1 x 100 unconfined strip (xlen 1000, K 10, sy 0.1,
    ss 1e-5), parabolic initial heads 9 -> 1, one drain in the last cell with elevation 0, conductance
    kh*delc*ddrn/(delr/2) and DRAINAGE DEPTH ddrn = 1 (AUXDEPTHNAME); 100 transient steps x 1.1 over 100 days;
    BICGSTAB, outer_dvclose 1e-6, inner_dvclose 1e-9, rclose 0.01 STRICT.  Case b adds NEWTON (cubic scaling,
    drn_fn terms).  Returns (SimConfig, analytic drain discharge as a function of the drain cell's head)."""
    from modflow6_b200 import configs
    ncol, xlen = 100, 1000.0
    delr, delc = xlen / ncol, 1.0
    kh, h0, h1 = 10.0, 9.0, 1.0
    delev, ddrn = 0.0, h1
    dcond = kh * delc * ddrn / (0.5 * delr)
    x = np.arange(0, xlen - delr / 2, delr)
    strt = np.sqrt(h0 ** 2 + x * (h1 ** 2 - h0 ** 2) / (xlen - delr))
    m = build_dis_model(1, 1, ncol, delr, delc, 10.0, [0.0], kh, icelltype=1, strt=strt.reshape(1, 1, ncol),
                        ss=1e-5, sy=0.1, iconvert=1, inewton=1 if newton else 0)
    drn = Package(T.PKG_DRN, [ncol - 1], [delev], [dcond], [ddrn], iflowred=1 if newton else 0)
    ims = T.ImsSettings.make(dvclose=1e-9, rclose=0.01, icnvgopt=1, iter1=200, ilinmeth=2)
    sln = T.SlnSettings.make(dvclose=1e-6, mxiter=200)
    cfg = configs.SimConfig("drn_ddrn01" + ("b" if newton else "a"), m,
                            [configs.Period(100.0, 100, 1.1, False, [drn])], sln, ims)

    def analytic(h):
        xd = np.asarray(h) - delev
        sat = xd / ddrn
        f = (-1.0 / ddrn ** 3) * xd ** 3 + (2.0 / ddrn ** 2) * xd ** 2 if newton else sat.copy()
        f = np.where(sat < 0, 0.0, np.where(sat > 1, 1.0, f))
        return f * dcond * (delev - np.asarray(h))
    return cfg, analytic


NPF_THICKSTRT = {"thickstrt": [False, True, True, False, False, False, True, False, False],
                 "icelltype": [0, 0, -1, 1, -1, 0, -1, 1, -1],
                 "hfb": [False, False, False, False, False, True, True, True, True]}


def npf_thickstrt_case(idx):
    """autotest/test_gwf_npf_thickstrt.py:8-127 literally: 1 x 1 x 6 strip, top 10, bottom 0, K 1, starting head 5,
    constant heads 6 / 4 at the ends, one steady step; CG with relaxation 1 (MILU0), closures 1e-6, 10 outer x 5 inner
    iterations.  The nine cases vary ICELLTYPE (0 / 1 / -1), THICKSTRT and one horizontal flow barrier between cells 3
    and 4 (hydraulic characteristic 1e-4).  Returns (SimConfig, hfb or None, heads, CHD inflow) with the reference
    test's literal answers (:140-194)."""
    from modflow6_b200 import configs
    m = build_dis_model(1, 1, 6, 1.0, 1.0, 10.0, [0.0], 1.0, icelltype=NPF_THICKSTRT["icelltype"][idx], strt=5.0,
                        ithickstrt=1 if NPF_THICKSTRT["thickstrt"][idx] else 0)
    chd = Package(T.PKG_CHD, [0, 5], [6.0, 4.0])
    ims = T.ImsSettings.make(dvclose=1e-6, rclose=1e-6, iter1=5, ilinmeth=1, relax=1.0)
    sln = T.SlnSettings.make(dvclose=1e-6, mxiter=10)
    cfg = configs.SimConfig(f"npf_thickstrt{idx + 1:02d}", m, [configs.Period(1.0, 1, 1.0, True, [chd])], sln, ims)
    hfb = ([2], [3], [1.0e-4]) if NPF_THICKSTRT["hfb"][idx] else None
    linear = np.linspace(6, 4, 6)
    water_table = np.array((6.0, 5.65716, 5.29206, 4.89969, 4.47276, 4.0))
    confined_hfb = np.array((6.0, 5.9998, 5.9996, 4.0004, 4.0002, 4.0))
    thickstrt_hfb = np.array((6.0, 5.9996004, 5.9992008, 4.0007992, 4.0003996, 4.0))
    unconfined_hfb = np.array((6.0, 5.99983342, 5.99966683, 4.00049971, 4.00024986, 4.0))
    heads = [linear, linear, linear, water_table, water_table, confined_hfb, thickstrt_hfb, unconfined_hfb,
             unconfined_hfb][idx]
    inflow = [4.0, 4.0, 2.0, 1.9965396769631871, 1.9965396769631871, 1.9990e-03, 1.9980e-03, 9.9949e-04,
              9.9949e-04][idx]
    return cfg, hfb, heads, inflow


def nested_grid_case():
    """A locally refined 2-D grid as ONE DISU model, the textbook case for the ghost node correction: 3 x 4 coarse
    cells of 2 x 2 (x 0..6) next to 4 x 8 fine cells of 1 x 1 (x 6..10), one confined layer, K = 1.  Every coarse cell
    of the last coarse column touches TWO fine cells whose centres sit 0.5 above / below its own, so the plain
    two-point flux is wrong for any head field with a y-gradient; the ghost node on the coarse side, interpolated
    between the coarse cell and its neighbour above / below with alpha = 0.5 / 2, removes that error exactly for a
    LINEAR field.  All boundary cells are constant heads on h = 10 + 0.7 x + 0.3 y.
    Returns (model, chd package, gnc tuple (noden, nodem, nodesj, alphasj), exact heads)."""
    from modflow6_b200.grid import build_disu_model
    ncx, ncy, nfx, nfy = 3, 4, 4, 8
    nc = ncx * ncy
    n = nc + nfx * nfy
    cid = lambda r, c: r * ncx + c                      # noqa: E731
    fid = lambda r, c: nc + r * nfx + c                 # noqa: E731
    x, y, size = np.zeros(n), np.zeros(n), np.zeros(n)
    for r in range(ncy):
        for c in range(ncx):
            x[cid(r, c)], y[cid(r, c)], size[cid(r, c)] = 2 * c + 1.0, 2 * r + 1.0, 2.0
    for r in range(nfy):
        for c in range(nfx):
            x[fid(r, c)], y[fid(r, c)], size[fid(r, c)] = 6.5 + c, r + 0.5, 1.0
    nbr = [dict() for _ in range(n)]                    # node -> {neighbour: (cl of this side, width)}

    def link(a, b, cla, clb, w):
        nbr[a][b] = (cla, w)
        nbr[b][a] = (clb, w)
    for r in range(ncy):
        for c in range(ncx):
            if c + 1 < ncx:
                link(cid(r, c), cid(r, c + 1), 1.0, 1.0, 2.0)
            if r + 1 < ncy:
                link(cid(r, c), cid(r + 1, c), 1.0, 1.0, 2.0)
    for r in range(nfy):
        for c in range(nfx):
            if c + 1 < nfx:
                link(fid(r, c), fid(r, c + 1), 0.5, 0.5, 1.0)
            if r + 1 < nfy:
                link(fid(r, c), fid(r + 1, c), 0.5, 0.5, 1.0)
    gn, gm, gj, ga = [], [], [], []
    for r in range(ncy):
        for half in (0, 1):
            a, b = cid(r, ncx - 1), fid(2 * r + half, 0)
            link(a, b, 1.0, 0.5, 1.0)
            rj = r + (1 if half else -1)                # the coarse neighbour on the side of the fine cell's centre
            if 0 <= rj < ncy:
                gn.append(a); gm.append(b); gj.append([cid(rj, ncx - 1)]); ga.append([0.25])
    iac, ja, ihc, cl12, hwva = [], [], [], [], []
    for a in range(n):
        cols = sorted(nbr[a])
        iac.append(1 + len(cols))
        ja += [a] + cols
        ihc += [1] * (1 + len(cols))
        cl12 += [0.0] + [nbr[a][b][0] for b in cols]
        hwva += [0.0] + [nbr[a][b][1] for b in cols]
    exact = 10.0 + 0.7 * x + 0.3 * y
    m = build_disu_model(np.array(iac), np.array(ja), np.array(ihc), np.array(cl12), np.array(hwva), 1.0, 0.0,
                         size * size, 1.0, icelltype=0, strt=10.0)
    edge = [a for a in range(n) if (x[a] - size[a] / 2 <= 0 or x[a] + size[a] / 2 >= 10
                                    or y[a] - size[a] / 2 <= 0 or y[a] + size[a] / 2 >= 8)]
    chd = Package(T.PKG_CHD, edge, exact[edge])
    return m, chd, (np.array(gn), np.array(gm), np.array(gj), np.array(ga)), exact
