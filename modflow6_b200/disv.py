"""Synthetic DISV (vertex-discretised, layered unstructured) models: hexagonal and triangular cell2d
tilings extruded over `nlay` layers (BASELINE config 4).

The arrays follow what `disvconnections` / `vertexconnect` + `cprops` build in the reference
(src/Model/ModelUtilities/Connections.f90:702-800, 1272-1364; src/Model/ModelUtilities/DisvGeom.f90:136-202):
node = k*ncpl + icell2d; CSR rows diagonal first then ascending columns; horizontal connections have
ihc = 1, cl1/cl2 = distance from each cell centre to the shared edge, hwva = shared edge length;
vertical connections ihc = 0, cl = half thickness, hwva = cell area.  Upper-triangle connections are
numbered in (n, ascending m) order (`filljas`, Connections.f90:1235-1268).
"""
import numpy as np

from .grid import GwfModel


def hex_cell2d(nr, nc, s=50.0):
    """Regular pointy-top hexagons, 'odd-r' offset rows; edge length s.  6 neighbours in the interior."""
    r, c = np.meshgrid(np.arange(nr), np.arange(nc), indexing="ij")
    r, c = r.reshape(-1), c.reshape(-1)
    odd = (r % 2).astype(np.int64)
    dr = np.array([0, 0, -1, -1, 1, 1])
    dc_even = np.array([-1, 1, -1, 0, -1, 0])
    dc_odd = np.array([-1, 1, 0, 1, 0, 1])
    rr = r[:, None] + dr[None, :]
    cc = c[:, None] + np.where(odd[:, None] == 1, dc_odd[None, :], dc_even[None, :])
    ok = (rr >= 0) & (rr < nr) & (cc >= 0) & (cc < nc)
    nbr = np.where(ok, rr * nc + cc, -1)
    ncpl = nr * nc
    half = np.sqrt(3.0) / 2.0 * s
    return dict(ncpl=ncpl, nbr=nbr, nbr_cl=np.where(ok, half, 0.0), nbr_len=np.where(ok, s, 0.0),
                area=np.full(ncpl, 1.5 * np.sqrt(3.0) * s * s), kind="hexagonal")


def tri_cell2d(nr, nc2, s=50.0):
    """Rows of alternating up / down equilateral triangles (nc2 triangles per row); <= 3 neighbours."""
    r, c = np.meshgrid(np.arange(nr), np.arange(nc2), indexing="ij")
    r, c = r.reshape(-1), c.reshape(-1)
    up = ((r + c) % 2 == 0)
    rr = np.stack([r, r, np.where(up, r + 1, r - 1)], axis=1)
    cc = np.stack([c - 1, c + 1, c], axis=1)
    ok = (rr >= 0) & (rr < nr) & (cc >= 0) & (cc < nc2)
    nbr = np.where(ok, rr * nc2 + cc, -1)
    ncpl = nr * nc2
    return dict(ncpl=ncpl, nbr=nbr, nbr_cl=np.where(ok, s / (2.0 * np.sqrt(3.0)), 0.0),
                nbr_len=np.where(ok, s, 0.0), area=np.full(ncpl, np.sqrt(3.0) / 4.0 * s * s), kind="triangular")


def build_disv_model(nlay, cell2d, top, botm, k11, k33=None, icelltype=0, strt=0.0, ss=None, sy=None,
                     iconvert=None, **opts):
    ncpl = cell2d["ncpl"]
    nbr2 = cell2d["nbr"]
    n = nlay * ncpl
    maxnb = nbr2.shape[1]
    # sort each cell's 2-d neighbours ascending (missing ones last)
    key = np.where(nbr2 >= 0, nbr2, np.iinfo(np.int64).max)
    order = np.argsort(key, axis=1, kind="stable")
    nbr2s = np.take_along_axis(nbr2, order, axis=1)
    cl2s = np.take_along_axis(cell2d["nbr_cl"], order, axis=1)
    len2s = np.take_along_axis(cell2d["nbr_len"], order, axis=1)
    k_idx = np.repeat(np.arange(nlay, dtype=np.int64), ncpl)
    ic = np.tile(np.arange(ncpl, dtype=np.int64), nlay)
    node = k_idx * ncpl + ic
    ncand = 3 + maxnb
    col = np.empty((n, ncand), dtype=np.int64)
    mask = np.zeros((n, ncand), dtype=bool)
    cl_own = np.zeros((n, ncand))          # distance from THIS cell's centre to the shared face
    width = np.zeros((n, ncand))           # hwva of the connection
    ihc = np.ones((n, ncand), dtype=np.int32)
    botm = np.asarray(botm, dtype=np.float64)
    bot3 = np.broadcast_to(botm[:, None] if botm.ndim == 1 else botm.reshape(nlay, ncpl), (nlay, ncpl))
    top3 = np.empty((nlay, ncpl))
    top3[0] = np.broadcast_to(np.asarray(top, dtype=np.float64), (ncpl,))
    if nlay > 1:
        top3[1:] = bot3[:-1]
    topv, botv = top3.reshape(-1).copy(), np.ascontiguousarray(bot3).reshape(-1).copy()
    thick = topv - botv
    area = np.tile(cell2d["area"], nlay)
    col[:, 0], mask[:, 0] = node, True
    col[:, 1], mask[:, 1] = node - ncpl, k_idx > 0                      # up
    cl_own[:, 1], width[:, 1], ihc[:, 1] = 0.5 * thick, area, 0
    nb = nbr2s[ic]                                                       # same layer
    col[:, 2:2 + maxnb] = np.where(nb >= 0, k_idx[:, None] * ncpl + nb, 0)
    mask[:, 2:2 + maxnb] = nb >= 0
    cl_own[:, 2:2 + maxnb] = cl2s[ic]
    width[:, 2:2 + maxnb] = len2s[ic]
    col[:, -1], mask[:, -1] = node + ncpl, k_idx < nlay - 1             # down
    cl_own[:, -1], width[:, -1], ihc[:, -1] = 0.5 * thick, area, 0
    cnt = mask.sum(axis=1)
    ia = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=ia[1:])
    ja = col[mask]
    cl_e, w_e, ihc_e = cl_own[mask], width[mask], ihc[mask]
    rows = np.repeat(np.arange(n, dtype=np.int64), cnt)
    # isym through sorted (row, col) keys
    key = rows * n + ja
    order = np.argsort(key, kind="stable")
    tkey = ja * n + rows
    loc = np.searchsorted(key[order], tkey)
    isym = order[loc]
    assert np.array_equal(key[order][loc], tkey), "cell2d neighbour lists are not symmetric"
    upper = ja > rows
    njas = int(upper.sum())
    jas = np.full(ja.size, -1, dtype=np.int64)
    jas[upper] = np.arange(njas)
    lower = ja < rows
    jas[lower] = jas[isym[lower]]
    m = GwfModel(nodes=n, ia=ia, ja=ja, jas=jas, isym=isym, ihc=ihc_e[upper], cl1=cl_e[upper],
                 cl2=cl_e[isym[upper]], hwva=w_e[upper], top=topv, bot=botv, area=area,
                 k11=np.broadcast_to(np.asarray(k11, dtype=np.float64).reshape(-1) if np.ndim(k11) else k11, (n,)).copy(),
                 k33=np.broadcast_to(np.asarray(k33 if k33 is not None else k11, dtype=np.float64).reshape(-1)
                                     if np.ndim(k33 if k33 is not None else k11) else (k33 if k33 is not None else k11), (n,)).copy(),
                 icelltype=np.broadcast_to(np.asarray(icelltype, dtype=np.int32).reshape(-1) if np.ndim(icelltype) else icelltype, (n,)).copy(),
                 strt=np.broadcast_to(np.asarray(strt, dtype=np.float64).reshape(-1) if np.ndim(strt) else strt, (n,)).copy(),
                 ibotnode=((nlay - 1) * ncpl + ic).astype(np.int32),
                 ss=None if ss is None else np.broadcast_to(np.asarray(ss, dtype=np.float64), (n,)).copy(),
                 sy=None if sy is None else np.broadcast_to(np.asarray(sy, dtype=np.float64), (n,)).copy(),
                 iconvert=None if iconvert is None else np.broadcast_to(np.asarray(iconvert, np.int32).reshape(-1) if np.ndim(iconvert) else iconvert, (n,)).copy(),
                 shape=(nlay, ncpl, 1), **opts)
    if ss is not None or sy is not None:
        m.insto = 1
    m.meta["cell2d"] = cell2d["kind"]
    return m


def cell2d_from_vertices(vertices, cells):
    """cell2d dictionary of build_disv_model from a DISV package's VERTICES / CELL2D blocks.

    vertices: (nvert, 2) x, y; cells: list of (xc, yc, [0-based vertex numbers, clockwise]).
    Restates `disvconnections` + `vertexconnect` (Connections.f90:702-790, 1272-1361) and
    `DisvGeomType%cprops` (DisvGeom.f90:136-202): two cells are connected when they share an EDGE -- two
    consecutive vertices, traversed forward in one list and backward in the other (`shared_edge`, :300-333);
    hwva = edge length, cl = normal distance from the cell centre to the edge line (`distance_normal`,
    :432-447), area by the shoelace formula about the first vertex (`get_area`, :340-386)."""
    vertices = np.asarray(vertices, dtype=np.float64)
    ncpl = len(cells)
    closed = []
    for _, _, iv in cells:
        iv = list(iv)
        if iv[0] != iv[-1]:
            iv.append(iv[0])          # the reference closes the polygon when it loads CELL2D
        closed.append(iv)
    # directed edge (a, b) -> cell; the neighbour across it owns (b, a)
    owner = {}
    for j, iv in enumerate(closed):
        for a, b in zip(iv[:-1], iv[1:]):
            owner[(a, b)] = j
    nbr_l, cl_l, len_l = [], [], []
    area = np.zeros(ncpl)
    for j, ((xc, yc, _), iv) in enumerate(zip(cells, closed)):
        nb, cl, ln = [], [], []
        for a, b in zip(iv[:-1], iv[1:]):
            m = owner.get((b, a))
            if m is None or m == j or m in nb:
                continue
            x1, y1 = vertices[a]
            x2, y2 = vertices[b]
            d = np.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2)
            nb.append(m)
            ln.append(d)
            cl.append(abs((x2 - x1) * (y1 - yc) - (x1 - xc) * (y2 - y1)) / d)
        nbr_l.append(nb); cl_l.append(cl); len_l.append(ln)
        x = vertices[iv, 0]
        y = vertices[iv, 1]
        a1 = np.sum((x[:-1] - x[0]) * (y[1:] - y[0]))
        a2 = np.sum((x[1:] - x[0]) * (y[:-1] - y[0]))
        area[j] = 0.5 * abs(a1 - a2)
    maxnb = max(1, max(len(v) for v in nbr_l))
    nbr = np.full((ncpl, maxnb), -1, dtype=np.int64)
    nbr_cl = np.zeros((ncpl, maxnb))
    nbr_len = np.zeros((ncpl, maxnb))
    for j in range(ncpl):
        k = len(nbr_l[j])
        nbr[j, :k], nbr_cl[j, :k], nbr_len[j, :k] = nbr_l[j], cl_l[j], len_l[j]
    return dict(ncpl=ncpl, nbr=nbr, nbr_cl=nbr_cl, nbr_len=nbr_len, area=area, kind="vertices")
