"""Synthetic GWF model builders: produce exactly the arrays the reference's
ConnectionsType / GwfNpfType / GwfStoType / BndType would hand to the solver.

DIS connectivity follows src/Model/ModelUtilities/Connections.f90:463-700
(`disconnections`): node = k*nrow*ncol + i*ncol + j (0-based here), CSR row =
[diag, up(k-1), back(i-1), left(j-1), right(j+1), front(i+1), down(k+1)],
upper-triangle connections numbered in (k,i,j) order as right, front, down;
cl1/cl2 = half cell sizes, hwva = perpendicular width (horizontal) or area
(vertical), ihc = 1 horizontal / 0 vertical.

All index arrays produced here are 0-based (index_base = 0).
"""
from dataclasses import dataclass, field

import numpy as np

from . import ctypes_types as T


@dataclass
class Package:
    type: int
    nodelist: np.ndarray
    b1: np.ndarray
    b2: np.ndarray = None
    b3: np.ndarray = None
    iflowred: int = 0
    flowred: float = 0.1
    auxnames: tuple = ()          # AUXILIARY variable names and their values [nbound, naux]: carried to the
    aux: np.ndarray = None        # budget file records (save_print_model_flows), not used by the solve
    bound_index: np.ndarray = None   # 0-based position of every boundary in the package's own list, when entries of
                                     # that list do not exist in the model (array-based recharge over removed cells)
    # time series (TS6) and AUXMULTNAME of an input deck, resolved per time step by at_time():
    ts_links: list = None            # [(target, row, SERIES NAME)], target = "b1" | "b2" | "b3" | ("aux", j)
    series: dict = None              # {NAME: timeseries.TimeSeries}
    mult: tuple = None               # (column "b1" | "b2", j): that column is multiplied by auxiliary variable j

    def with_nodes(self, nodelist, keep=None):
        """the same boundaries at other node numbers (model offset in a merged solution, user -> reduced);
        keep: boolean mask of the boundaries that exist (the others are dropped, their list positions remembered)"""
        k = slice(None) if keep is None else np.asarray(keep, dtype=bool)
        bi = self.bound_index if self.bound_index is not None else (None if keep is None else np.arange(self.nodelist.size))
        links = self.ts_links
        if links and keep is not None:
            newrow = np.cumsum(k) - 1
            links = [(t, int(newrow[r]), nm) for t, r, nm in links if k[r]]
        return Package(self.type, np.asarray(nodelist)[k], self.b1[k], self.b2[k], self.b3[k], iflowred=self.iflowred,
                       flowred=self.flowred, auxnames=self.auxnames, aux=None if self.aux is None else self.aux[k],
                       bound_index=None if bi is None else np.asarray(bi)[k], ts_links=links, series=self.series,
                       mult=self.mult)

    @property
    def time_dependent(self):
        return bool(self.ts_links)

    def at_time(self, time0, time1):
        """the package as the solve sees it during the time step [time0, time1]: every entry linked to a time series
        takes the series' value (tsmgr_ad, TimeSeriesManager.f90:134-260: auxiliary links first, because one of
        them may be the multiplier), then the multiplier column is applied (bnd `*_mult` functions, e.g.
        gwf-wel.f90 q_mult, gwf-chd.f90 head_mult, gwf-riv.f90 cond_mult)"""
        if not self.ts_links and self.mult is None:
            return self
        cols = {"b1": self.b1.copy(), "b2": self.b2.copy(), "b3": self.b3.copy()}
        aux = None if self.aux is None else self.aux.copy()
        for target, row, name in self.ts_links or ():
            val = self.series[name].value(time0, time1)
            if isinstance(target, tuple):
                aux[row, target[1]] = val
            else:
                cols[target][row] = val
        if self.mult is not None:
            col, j = self.mult
            cols[col] = cols[col] * aux[:, j]
        return Package(self.type, self.nodelist, cols["b1"], cols["b2"], cols["b3"], iflowred=self.iflowred,
                       flowred=self.flowred, auxnames=self.auxnames, aux=aux, bound_index=self.bound_index)

    def __post_init__(self):
        self.nodelist = T.as_i32(self.nodelist)
        self.b1 = T.as_f64(self.b1)
        nb = self.nodelist.size
        self.b2 = T.as_f64(self.b2) if self.b2 is not None else np.zeros(nb)
        self.b3 = T.as_f64(self.b3) if self.b3 is not None else np.zeros(nb)

    def struct(self):
        return T.BndPackageStruct(self.type, self.nodelist.size, 0, self.iflowred, self.flowred,
                                  T.ptr_i32(self.nodelist), T.ptr_f64(self.b1),
                                  T.ptr_f64(self.b2), T.ptr_f64(self.b3))


def package_array(pkgs):
    arr = (T.BndPackageStruct * max(1, len(pkgs)))()
    for i, p in enumerate(pkgs):
        arr[i] = p.struct()
    return arr


@dataclass
class GwfModel:
    """Arrays of one GWF model (0-based indices)."""

    nodes: int
    ia: np.ndarray
    ja: np.ndarray
    jas: np.ndarray
    isym: np.ndarray
    ihc: np.ndarray
    cl1: np.ndarray
    cl2: np.ndarray
    hwva: np.ndarray
    top: np.ndarray
    bot: np.ndarray
    area: np.ndarray
    k11: np.ndarray
    k33: np.ndarray
    icelltype: np.ndarray
    strt: np.ndarray
    ibound: np.ndarray = None
    ibotnode: np.ndarray = None
    ss: np.ndarray = None
    sy: np.ndarray = None
    iconvert: np.ndarray = None
    icellavg: int = 0
    inewton: int = 0
    inewtonur: int = 0
    iperched: int = 0
    ivarcv: int = 0
    idewatcv: int = 0
    insto: int = 0
    istor_coef: int = 0
    iconf_ss: int = 0
    iorig_ss: int = 0
    ithickstrt: int = 0      # NPF THICKSTRT: cells with icelltype < 0 are confined with the thickness of their starting head
    shape: tuple = None
    meta: dict = field(default_factory=dict)
    # NPF anisotropy (all None = K22 == K, no rotation): k22 [nodes], angle1/2/3 [nodes] in radians, and the unit
    # normal of every connection from its lower- to its higher-numbered cell
    k22: np.ndarray = None
    angle1: np.ndarray = None
    angle2: np.ndarray = None
    angle3: np.ndarray = None
    conn_nx: np.ndarray = None
    conn_ny: np.ndarray = None
    # NPF REWET: wetdry [nodes] (None = no rewetting) and the REWET record
    wetdry: np.ndarray = None
    irewet: int = 0
    wetfct: float = 1.0
    iwetit: int = 1
    ihdwet: int = 0

    def __post_init__(self):
        n = self.nodes
        for name in ("ia", "ja", "jas", "isym", "ihc", "icelltype"):
            setattr(self, name, T.as_i32(getattr(self, name)))
        for name in ("cl1", "cl2", "hwva", "top", "bot", "area", "k11", "k33", "strt"):
            setattr(self, name, T.as_f64(getattr(self, name)))
        self.ibound = T.as_i32(self.ibound) if self.ibound is not None else np.ones(n, np.int32)
        self.ibotnode = T.as_i32(self.ibotnode) if self.ibotnode is not None else np.arange(n, dtype=np.int32)
        self.ss = T.as_f64(self.ss) if self.ss is not None else np.zeros(n)
        self.sy = T.as_f64(self.sy) if self.sy is not None else np.zeros(n)
        self.iconvert = T.as_i32(self.iconvert) if self.iconvert is not None else np.zeros(n, np.int32)
        for name in ("k22", "angle1", "angle2", "angle3", "conn_nx", "conn_ny", "wetdry"):
            v = getattr(self, name)
            if v is not None:
                size = self.ihc.size if name.startswith("conn_") else n
                setattr(self, name, T.as_f64(np.broadcast_to(np.asarray(v, dtype=np.float64), (size,)).copy()))

    @property
    def nja(self):
        return int(self.ja.size)

    @property
    def njas(self):
        return int(self.ihc.size)

    def struct(self):
        s = T.GwfModelStruct()
        s.index_base = 0
        s.nodes, s.nja, s.njas = self.nodes, self.nja, self.njas
        for name in ("ia", "ja", "jas", "isym", "ihc", "ibound", "icelltype", "ibotnode", "iconvert"):
            setattr(s, name, T.ptr_i32(getattr(self, name)))
        for name in ("cl1", "cl2", "hwva", "top", "bot", "area", "strt", "k11", "k33", "ss", "sy"):
            setattr(s, name, T.ptr_f64(getattr(self, name)))
        for name in ("icellavg", "inewton", "inewtonur", "iperched", "ivarcv", "idewatcv", "insto",
                     "istor_coef", "iconf_ss", "iorig_ss"):
            setattr(s, name, int(getattr(self, name)))
        s.ithickstrt = int(self.ithickstrt)
        for name in ("k22", "angle1", "angle2", "angle3", "conn_nx", "conn_ny", "wetdry"):
            v = getattr(self, name)
            if v is not None:
                setattr(s, name, T.ptr_f64(v))
        s.irewet, s.iwetit, s.ihdwet, s.wetfct = int(self.irewet), int(self.iwetit), int(self.ihdwet), float(self.wetfct)
        return s

    def node(self, k, i, j):
        nlay, nrow, ncol = self.shape
        return (k * nrow + i) * ncol + j


def _bcast(v, shape):
    a = np.asarray(v, dtype=np.float64)
    return np.ascontiguousarray(np.broadcast_to(a, shape)).reshape(-1)


def dis_connectivity(nlay, nrow, ncol):
    """ia, ja, jas, isym (0-based) and per-direction masks for a full DIS grid."""
    n = nlay * nrow * ncol
    nrc = nrow * ncol
    idx = np.arange(n, dtype=np.int64)
    k = idx // nrc
    rem = idx - k * nrc
    i = rem // ncol
    j = rem - i * ncol
    # direction order: up, back, left, right, front, down
    exist = np.empty((n, 6), dtype=bool)
    exist[:, 0] = k > 0
    exist[:, 1] = i > 0
    exist[:, 2] = j > 0
    exist[:, 3] = j < ncol - 1
    exist[:, 4] = i < nrow - 1
    exist[:, 5] = k < nlay - 1
    offs = np.array([-nrc, -ncol, -1, 1, ncol, nrc], dtype=np.int64)
    cnt = 1 + exist.sum(axis=1)
    ia = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=ia[1:])
    nja = int(ia[-1])
    # position of direction d inside its row (1 + number of earlier existing directions)
    pos_in_row = 1 + np.cumsum(exist, axis=1) - exist  # (n,6)
    # upper connection numbering
    hasr, hasf, hasd = exist[:, 3], exist[:, 4], exist[:, 5]
    ucnt = hasr.astype(np.int64) + hasf + hasd
    ustart = np.zeros(n, dtype=np.int64)
    np.cumsum(ucnt[:-1], out=ustart[1:])
    njas = int(ucnt.sum())
    # candidate arrays (n,7): col, jas, isym
    col = np.empty((n, 7), dtype=np.int64)
    jas = np.empty((n, 7), dtype=np.int64)
    isym = np.empty((n, 7), dtype=np.int64)
    mask = np.empty((n, 7), dtype=bool)
    col[:, 0] = idx
    jas[:, 0] = -1
    isym[:, 0] = ia[:-1]
    mask[:, 0] = True
    opp = [5, 4, 3, 2, 1, 0]
    # upper index of the connection leaving cell c in direction right/front/down
    uidx = {3: ustart, 4: ustart + hasr, 5: ustart + hasr + hasf}
    for d in range(6):
        m = idx + offs[d]
        ok = exist[:, d]
        msafe = np.where(ok, m, 0)
        col[:, d + 1] = msafe
        mask[:, d + 1] = ok
        isym[:, d + 1] = ia[msafe] + pos_in_row[msafe, opp[d]]
        if d >= 3:
            jas[:, d + 1] = uidx[d]
        else:
            jas[:, d + 1] = uidx[opp[d]][msafe]
    ja = col[mask].astype(np.int32)
    jas_f = jas[mask].astype(np.int32)
    isym_f = isym[mask].astype(np.int32)
    return dict(n=n, nja=nja, njas=njas, ia=ia.astype(np.int32), ja=ja, jas=jas_f, isym=isym_f,
                k=k, i=i, j=j, hasr=hasr, hasf=hasf, hasd=hasd, ustart=ustart)


def build_dis_model(nlay, nrow, ncol, delr, delc, top, botm, k11, k33=None, icelltype=0,
                    strt=0.0, ss=None, sy=None, iconvert=None, **opts):
    """Full (no idomain holes) DIS model.  `top` scalar/(nrow,ncol); `botm`
    (nlay,) / (nlay,nrow,ncol) layer bottoms; k11/k33/icelltype/strt scalar or
    (nlay,nrow,ncol)."""
    c = dis_connectivity(nlay, nrow, ncol)
    n = c["n"]
    shp = (nlay, nrow, ncol)
    delr = np.broadcast_to(np.asarray(delr, dtype=np.float64), (ncol,)).copy()
    delc = np.broadcast_to(np.asarray(delc, dtype=np.float64), (nrow,)).copy()
    botm = np.asarray(botm, dtype=np.float64)
    if botm.ndim == 1:
        botm = botm[:, None, None]
    bot3 = np.broadcast_to(botm, shp)
    top3 = np.empty(shp)
    top3[0] = np.broadcast_to(np.asarray(top, dtype=np.float64), (nrow, ncol))
    if nlay > 1:
        top3[1:] = bot3[:-1]
    topv = top3.reshape(-1).copy()
    botv = np.ascontiguousarray(bot3).reshape(-1).copy()
    area = np.broadcast_to(delc[:, None] * delr[None, :], shp).reshape(-1).copy()
    # per upper connection geometry, in (k,i,j) order right, front, down
    k, i, j = c["k"], c["i"], c["j"]
    hasr, hasf, hasd, ustart = c["hasr"], c["hasf"], c["hasd"], c["ustart"]
    njas = c["njas"]
    ihc = np.empty(njas, dtype=np.int32)
    cl1 = np.empty(njas)
    cl2 = np.empty(njas)
    hwva = np.empty(njas)
    nrc = nrow * ncol
    # right
    s = np.nonzero(hasr)[0]
    u = ustart[s]
    ihc[u] = 1
    cl1[u] = 0.5 * delr[j[s]]
    cl2[u] = 0.5 * delr[j[s] + 1]
    hwva[u] = delc[i[s]]
    # front
    s = np.nonzero(hasf)[0]
    u = ustart[s] + hasr[s]
    ihc[u] = 1
    cl1[u] = 0.5 * delc[i[s]]
    cl2[u] = 0.5 * delc[i[s] + 1]
    hwva[u] = delr[j[s]]
    # down
    s = np.nonzero(hasd)[0]
    u = ustart[s] + hasr[s] + hasf[s]
    ihc[u] = 0
    cl1[u] = 0.5 * (topv[s] - botv[s])
    cl2[u] = 0.5 * (topv[s + nrc] - botv[s + nrc])
    hwva[u] = delr[j[s]] * delc[i[s]]
    k11v = _bcast(k11, shp)
    k33v = _bcast(k33 if k33 is not None else k11, shp)
    ict = np.ascontiguousarray(np.broadcast_to(np.asarray(icelltype, dtype=np.int32), shp)).reshape(-1)
    ibot = (np.arange(n, dtype=np.int64) % nrc + (nlay - 1) * nrc).astype(np.int32)
    # anisotropy: per-cell arrays + the connection normals of DisType%connection_normal (Dis.f90:1039-1085):
    # towards the next column (1, 0), towards the next row ("front") (0, -1), vertical (0, 0)
    for name in ("k22", "angle1", "angle2", "angle3", "wetdry"):
        if opts.get(name) is not None:
            opts[name] = _bcast(opts[name], shp)
    if opts.get("k22") is not None or opts.get("angle1") is not None:
        nx, ny = np.zeros(njas), np.zeros(njas)
        sr = np.nonzero(hasr)[0]
        nx[ustart[sr]] = 1.0
        sf = np.nonzero(hasf)[0]
        ny[ustart[sf] + hasr[sf]] = -1.0
        opts["conn_nx"], opts["conn_ny"] = nx, ny
    m = GwfModel(nodes=n, ia=c["ia"], ja=c["ja"], jas=c["jas"], isym=c["isym"], ihc=ihc, cl1=cl1,
                 cl2=cl2, hwva=hwva, top=topv, bot=botv, area=area, k11=k11v, k33=k33v,
                 icelltype=ict, strt=_bcast(strt, shp), ibotnode=ibot,
                 ss=_bcast(ss, shp) if ss is not None else None,
                 sy=_bcast(sy, shp) if sy is not None else None,
                 iconvert=(np.ascontiguousarray(np.broadcast_to(np.asarray(iconvert, dtype=np.int32), shp)).reshape(-1)
                           if iconvert is not None else None),
                 shape=shp, **opts)
    if ss is not None or sy is not None:
        m.insto = 1
    return m


def tdis_steps(perlen, nstp, tsmult):
    """Time-step lengths of one stress period (src/Timing/tdis.f90:255-267)."""
    if tsmult == 1.0:
        d0 = perlen / nstp
    else:
        d0 = perlen * (1.0 - tsmult) / (1.0 - tsmult ** nstp)
    out = [d0]
    for _ in range(nstp - 1):
        out.append(out[-1] * tsmult)
    return out


def merge_models(models, exchanges):
    """Serial multi-model solution: N GWF models + GWF-GWF exchanges in ONE system matrix, the way
    sln_connect lays the models out one after the other (NumericalSolution.f90:2336-2381) and gwf_gwf_ac /
    gwf_gwf_fc add the cross-model terms (exg-gwfgwf.f90:363-411, 488-550).

    `exchanges`: list of dicts(m1=, m2=, nodem1=, nodem2=, ihc=, cl1=, cl2=, hwva=) with 0-based cell ids of the
    EXCHANGEDATA block.  Returns the merged GwfModel (cell n of model k becomes offset[k] + n) and the offsets.
    All models must share the NPF / STO options."""
    offs = np.concatenate([[0], np.cumsum([m.nodes for m in models])]).astype(np.int64)
    n = int(offs[-1])
    # upper-triangle connection list of the merged model: (n, m, ihc, cl1, cl2, hwva)
    rows, cols, ihc, cl1, cl2, hw = [], [], [], [], [], []
    for k, m in enumerate(models):
        r = np.repeat(np.arange(m.nodes, dtype=np.int64), np.diff(m.ia))
        up = m.ja > r
        j = m.jas[up]
        rows.append(r[up] + offs[k]); cols.append(m.ja[up].astype(np.int64) + offs[k])
        ihc.append(m.ihc[j]); cl1.append(m.cl1[j]); cl2.append(m.cl2[j]); hw.append(m.hwva[j])
    for e in exchanges:
        a = np.asarray(e["nodem1"], dtype=np.int64) + offs[e["m1"]]
        b = np.asarray(e["nodem2"], dtype=np.int64) + offs[e["m2"]]
        swap = a > b
        rows.append(np.where(swap, b, a)); cols.append(np.where(swap, a, b))
        ihc.append(np.asarray(e["ihc"], dtype=np.int32))
        c1, c2 = np.asarray(e["cl1"], dtype=np.float64), np.asarray(e["cl2"], dtype=np.float64)
        cl1.append(np.where(swap, c2, c1)); cl2.append(np.where(swap, c1, c2))
        hw.append(np.asarray(e["hwva"], dtype=np.float64))
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    ihc, cl1, cl2, hw = np.concatenate(ihc), np.concatenate(cl1), np.concatenate(cl2), np.concatenate(hw)
    order = np.lexsort((cols, rows))            # (n, ascending m): the filljas numbering
    rows, cols, ihc, cl1, cl2, hw = rows[order], cols[order], ihc[order], cl1[order], cl2[order], hw[order]
    njas = rows.size
    # full CSR: diagonal first, then ascending columns
    allr = np.concatenate([np.arange(n), rows, cols])
    allc = np.concatenate([np.arange(n), cols, rows])
    alljas = np.concatenate([np.full(n, -1), np.arange(njas), np.arange(njas)])
    isdiag = np.concatenate([np.zeros(n), np.ones(2 * njas)])
    o2 = np.lexsort((allc, isdiag, allr))
    ja, jas = allc[o2], alljas[o2]
    ia = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(allr, minlength=n), out=ia[1:])
    r2 = allr[o2]
    key = r2 * n + ja
    ko = np.argsort(key, kind="stable")
    isym = ko[np.searchsorted(key[ko], ja * n + r2)]
    cat = lambda name: np.concatenate([getattr(m, name) for m in models])  # noqa: E731
    m0 = models[0]
    # The merged system carries ONE set of NPF / STO options; the reference keeps them per model.  Models that
    # differ would be solved with model 0's options (silently wrong heads), so refuse them.
    for name in ("icellavg", "inewton", "inewtonur", "iperched", "ivarcv", "idewatcv", "insto", "istor_coef",
                 "iconf_ss", "iorig_ss", "irewet", "iwetit", "ihdwet", "wetfct"):
        vals = {float(getattr(m, name)) for m in models}
        if len(vals) > 1:
            raise ValueError(f"merge_models: the models differ in option `{name}` ({sorted(vals)}); one solution "
                             "matrix on the GPU path needs the same NPF / STO options in every model")
    if any(getattr(m, a) is not None for m in models for a in ("k22", "angle1", "angle2", "angle3")):
        raise ValueError("merge_models: K22 / ANGLE anisotropy is not supported in multi-model solutions")
    # THICKSTRT is a per-model option: a model without it turns its negative ICELLTYPE into 1 (prepcheck,
    # gwf-npf.f90:1838-1846), so doing that here lets the merged model carry THICKSTRT for the models that set it
    ict = np.concatenate([m.icelltype if m.ithickstrt else np.where(m.icelltype < 0, 1, m.icelltype)
                          for m in models]).astype(np.int32)
    return GwfModel(nodes=n, ia=ia, ja=ja, jas=jas, isym=isym, ihc=ihc, cl1=cl1, cl2=cl2, hwva=hw,
                    top=cat("top"), bot=cat("bot"), area=cat("area"), k11=cat("k11"), k33=cat("k33"),
                    icelltype=ict, ithickstrt=max(int(m.ithickstrt) for m in models),
                    strt=cat("strt"), ibound=cat("ibound"),
                    ibotnode=np.concatenate([m.ibotnode + offs[k] for k, m in enumerate(models)]),
                    ss=cat("ss"), sy=cat("sy"), iconvert=cat("iconvert"), icellavg=m0.icellavg,
                    inewton=m0.inewton, inewtonur=m0.inewtonur, iperched=m0.iperched, ivarcv=m0.ivarcv,
                    idewatcv=m0.idewatcv, insto=m0.insto, istor_coef=m0.istor_coef, iconf_ss=m0.iconf_ss,
                    iorig_ss=m0.iorig_ss, shape=None,
                    wetdry=cat("wetdry") if m0.irewet else None, irewet=m0.irewet, iwetit=m0.iwetit, ihdwet=m0.ihdwet,
                    wetfct=m0.wetfct), offs


def reduce_model(m, keep):
    """Remove the cells with keep == False from any GwfModel the way the reference numbers a grid with IDOMAIN == 0
    cells (reduced node numbers in user order, DiscretizationBase `nodereduced` / `nodeuser`; connections to a
    removed cell do not exist, Connections.f90:463-700, 803-960).  The kept connections keep their relative order, so
    the symmetric numbering stays the filljas one.  `meta` gets nodeuser / nodereduced like build_dis_model_idomain."""
    import dataclasses
    keep = np.asarray(keep, dtype=bool)
    n_old = m.nodes
    nodeuser = np.nonzero(keep)[0]
    n = nodeuser.size
    red = np.full(n_old, -1, dtype=np.int64)
    red[nodeuser] = np.arange(n)
    rows = np.repeat(np.arange(n_old), np.diff(m.ia))
    kc = keep[rows] & keep[m.ja]                       # CSR entries that survive
    keep_s = np.zeros(m.njas, dtype=bool)
    keep_s[m.jas[kc & (m.jas >= 0)]] = True
    jmap = np.full(m.njas, -1, dtype=np.int64)
    jmap[keep_s] = np.arange(int(keep_s.sum()))
    pos = np.full(m.nja, -1, dtype=np.int64)
    pos[kc] = np.arange(int(kc.sum()))
    ia = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(red[rows[kc]], minlength=n), out=ia[1:])
    ja = red[m.ja[kc]]
    jas = np.where(m.jas[kc] >= 0, jmap[np.maximum(m.jas[kc], 0)], -1)
    isym = pos[m.isym[kc]]
    # bottom of every vertical chain (gwf-npf.f90:1904-1940): walk down through the vertical connections
    r2 = red[rows[kc]]
    cand = (m.ihc[np.maximum(m.jas[kc], 0)] == 0) & (m.jas[kc] >= 0) & (ja > r2) & (m.bot[nodeuser][ja] < m.bot[nodeuser][r2])
    below = np.full(n, -1, dtype=np.int64)
    bb = np.full(n, np.inf)
    for r, c in zip(r2[cand], ja[cand]):
        if m.bot[nodeuser[c]] < bb[r]:
            bb[r], below[r] = m.bot[nodeuser[c]], c
    ibot = np.arange(n, dtype=np.int64)
    for c in range(n - 1, -1, -1):
        if below[c] >= 0:
            ibot[c] = ibot[below[c]]
    per_node = {f: getattr(m, f)[nodeuser] for f in ("top", "bot", "area", "k11", "k33", "icelltype", "strt", "ibound",
                                                      "ss", "sy", "iconvert")}
    for f in ("k22", "angle1", "angle2", "angle3", "wetdry"):
        per_node[f] = None if getattr(m, f) is None else getattr(m, f)[nodeuser]
    per_conn = {f: getattr(m, f)[keep_s] for f in ("ihc", "cl1", "cl2", "hwva")}
    for f in ("conn_nx", "conn_ny"):
        per_conn[f] = None if getattr(m, f) is None else getattr(m, f)[keep_s]
    out = dataclasses.replace(m, nodes=n, ia=ia, ja=ja, jas=jas, isym=isym, ibotnode=ibot, meta=dict(m.meta),
                              **per_node, **per_conn)
    out.meta["nodeuser"] = nodeuser
    out.meta["nodereduced"] = red
    return out


def build_disu_model(iac, ja, ihc, cl12, hwva, top, bot, area, k11, k33=None, icelltype=0, strt=0.0, ss=None,
                     sy=None, iconvert=None, **opts):
    """Unstructured (DISU) model from the CONNECTIONDATA block as the user writes it (gwf-disu.dfn): `iac` entries
    per cell, `ja` with the cell itself first (0-based here), and `ihc`, `cl12`, `hwva` per ja entry.
    Follows `disuconnections` + `con_finalize` (Connections.f90:803-960, 245-330): the symmetric arrays take
    ihc / hwva from the upper-triangle entry, cl1 = cl12 at (n, m), cl2 = cl12 at (m, n).  Rows are stored
    diagonal first, then ascending columns (the solution-matrix layout); ibotnode = the last cell of the chain of
    vertical (ihc == 0, m > n) connections below a cell."""
    iac = np.asarray(iac, dtype=np.int64)
    n = iac.size
    ia_in = np.concatenate([[0], np.cumsum(iac)])
    ja = np.abs(np.asarray(ja, dtype=np.int64))
    ihc, cl12, hwva = np.asarray(ihc, dtype=np.int32), np.asarray(cl12, dtype=np.float64), np.asarray(hwva, np.float64)
    rows = np.repeat(np.arange(n, dtype=np.int64), iac)
    if not np.array_equal(ja[ia_in[:-1]], np.arange(n)):
        raise ValueError("DISU: the first ja entry of every cell must be the cell itself")
    isdiag = np.zeros(ja.size, dtype=np.int64)
    isdiag[ia_in[:-1]] = -1                              # sorts the diagonal first
    order = np.lexsort((ja, isdiag, rows))
    ja_s, ihc_e, cl_e, hw_e = ja[order], ihc[order], cl12[order], hwva[order]
    key = rows * n + ja_s                                # rows is unchanged by a within-row sort
    ko = np.argsort(key, kind="stable")
    tkey = ja_s * n + rows
    loc = np.minimum(np.searchsorted(key[ko], tkey), key.size - 1)
    if not np.array_equal(key[ko][loc], tkey):
        raise ValueError("DISU: the connection list is not symmetric")
    isym = ko[loc]
    upper = ja_s > rows
    njas = int(upper.sum())
    jas = np.full(ja_s.size, -1, dtype=np.int64)
    jas[upper] = np.arange(njas)
    lower = ja_s < rows
    jas[lower] = jas[isym[lower]]
    if not np.allclose(hw_e[upper], hw_e[isym[upper]]) or not np.array_equal(ihc_e[upper], ihc_e[isym[upper]]):
        raise ValueError("DISU: ihc / hwva differ between the two directions of a connection")
    # bottom of every vertical chain
    below = np.full(n, -1, dtype=np.int64)
    vert = upper & (ihc_e == 0)
    below[rows[vert][::-1]] = ja_s[vert][::-1]           # the first (lowest-numbered) vertical neighbour below wins
    ibot = np.arange(n, dtype=np.int64)
    for c in range(n - 1, -1, -1):
        if below[c] >= 0:
            ibot[c] = ibot[below[c]]
    bc = lambda v, dt=np.float64: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=dt), (n,))).copy()  # noqa: E731
    m = GwfModel(nodes=n, ia=ia_in, ja=ja_s, jas=jas, isym=isym, ihc=ihc_e[upper], cl1=cl_e[upper],
                 cl2=cl_e[isym[upper]], hwva=hw_e[upper], top=bc(top), bot=bc(bot), area=bc(area), k11=bc(k11),
                 k33=bc(k33 if k33 is not None else k11), icelltype=bc(icelltype, np.int32), strt=bc(strt),
                 ibotnode=ibot.astype(np.int32), ss=None if ss is None else bc(ss), sy=None if sy is None else bc(sy),
                 iconvert=None if iconvert is None else bc(iconvert, np.int32), shape=(n,), **opts)
    if ss is not None or sy is not None:
        m.insto = 1
    return m


def build_dis_model_idomain(nlay, nrow, ncol, delr, delc, top, botm, idomain, k11, k33=None, icelltype=0, strt=0.0,
                            ss=None, sy=None, iconvert=None, **opts):
    """DIS model with an IDOMAIN array, numbered like the reference numbers it: only cells with idomain > 0 exist
    (reduced node numbers in user order), idomain == 0 removes a cell, idomain < 0 makes it a vertical pass-through
    -- the cells above and below it are connected directly with their own half thicknesses (`disconnections`,
    Connections.f90:463-700: the vertical search `do kk = k+1, nlay ... if (mr >= 0) exit`).
    `model.meta` carries nodeuser (reduced -> user) and nodereduced (user -> reduced, -1 where no cell exists)."""
    shp = (nlay, nrow, ncol)
    nrc = nrow * ncol
    idom = np.asarray(idomain).reshape(shp)
    delr = np.broadcast_to(np.asarray(delr, dtype=np.float64), (ncol,))
    delc = np.broadcast_to(np.asarray(delc, dtype=np.float64), (nrow,))
    botm = np.asarray(botm, dtype=np.float64)
    bot3 = np.broadcast_to(botm[:, None, None] if botm.ndim == 1 else botm.reshape(shp), shp)
    top3 = np.concatenate([np.broadcast_to(np.asarray(top, dtype=np.float64), (1, nrow, ncol)), bot3[:-1]])
    area3 = np.broadcast_to(delc[:, None] * delr[None, :], shp)
    active = idom > 0
    nodeuser = np.nonzero(active.reshape(-1))[0]
    n = nodeuser.size
    nodereduced = np.full(nlay * nrc, -1, dtype=np.int64)
    nodereduced[nodeuser] = np.arange(n)
    # layer of the next cell below / above that is not a pass-through (idomain >= 0), -1 if none
    below = np.full(shp, -1, dtype=np.int64)
    above = np.full(shp, -1, dtype=np.int64)
    for k in range(nlay - 2, -1, -1):
        below[k] = np.where(idom[k + 1] >= 0, k + 1, below[k + 1])
    for k in range(1, nlay):
        above[k] = np.where(idom[k - 1] >= 0, k - 1, above[k - 1])
    ku, iu, ju = np.unravel_index(nodeuser, shp)
    half = 0.5 * (top3 - bot3)
    ent = [[] for _ in range(7)]        # per direction: (valid, user node of the neighbour, ihc, cl12 of THIS side, hwva)
    def add(valid, mk, mi, mj, ihc, cl, hw):
        mk, mi, mj = np.where(valid, mk, 0), np.where(valid, mi, 0), np.where(valid, mj, 0)
        valid = valid & (idom[mk, mi, mj] > 0)
        return valid, (mk * nrow + mi) * ncol + mj, ihc, cl, hw
    dirs = [add(above[ku, iu, ju] >= 0, above[ku, iu, ju], iu, ju, 0, half[ku, iu, ju], area3[ku, iu, ju]),
            add(iu > 0, ku, iu - 1, ju, 1, 0.5 * delc[iu], delr[ju]),
            add(ju > 0, ku, iu, ju - 1, 1, 0.5 * delr[ju], delc[iu]),
            add(ju < ncol - 1, ku, iu, ju + 1, 1, 0.5 * delr[ju], delc[iu]),
            add(iu < nrow - 1, ku, iu + 1, ju, 1, 0.5 * delc[iu], delr[ju]),
            add(below[ku, iu, ju] >= 0, below[ku, iu, ju], iu, ju, 0, half[ku, iu, ju], area3[ku, iu, ju])]
    valid = np.stack([np.ones(n, bool)] + [d[0] for d in dirs], axis=1)
    cols = np.stack([np.arange(n)] + [nodereduced[d[1]] for d in dirs], axis=1)
    ihc = np.stack([np.zeros(n, np.int32)] + [np.full(n, d[2], np.int32) for d in dirs], axis=1)
    cl12 = np.stack([np.zeros(n)] + [np.broadcast_to(d[3], (n,)) for d in dirs], axis=1)
    hwva = np.stack([np.zeros(n)] + [np.broadcast_to(d[4], (n,)) for d in dirs], axis=1)
    pick = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a), shp)).reshape(-1)[nodeuser]   # noqa: E731
    m = build_disu_model(valid.sum(axis=1), cols[valid], ihc[valid], cl12[valid], hwva[valid],
                         top3.reshape(-1)[nodeuser], bot3.reshape(-1)[nodeuser], area3.reshape(-1)[nodeuser], pick(k11),
                         k33=None if k33 is None else pick(k33), icelltype=pick(icelltype), strt=pick(strt),
                         ss=None if ss is None else pick(ss), sy=None if sy is None else pick(sy),
                         iconvert=None if iconvert is None else pick(iconvert), **opts)
    m.shape = shp
    m.meta["nodeuser"] = nodeuser
    m.meta["nodereduced"] = nodereduced
    return m
