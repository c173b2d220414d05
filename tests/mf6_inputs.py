"""Writers of small MODFLOW 6 input decks for the reader tests (what FloPy would write; the layout follows
the reference's `.mf6minsim/` example and doc/mf6io).  Only test infrastructure."""
import os

import numpy as np


def _w(path, text):
    with open(path, "w") as f:
        f.write("# written by tests/mf6_inputs.py\n" + text)


def _arr(name, a, layered=False):
    """READARRAY text: CONSTANT when uniform; LAYERED -> one control record per layer (a[k] scalar or 2-D)"""
    if layered:
        s = f"  {name} LAYERED\n"
        for ak in a:
            ak = np.asarray(ak, dtype=float)
            if ak.ndim == 0 or np.all(ak == ak.reshape(-1)[0]):
                s += f"    CONSTANT  {float(ak.reshape(-1)[0])!r}\n"
            else:
                s += "    INTERNAL FACTOR 1.0\n" + "\n".join("      " + " ".join(repr(float(v)) for v in row)
                                                           for row in ak) + "\n"
        return s
    a = np.asarray(a)
    if a.ndim == 0 or np.all(a == a.reshape(-1)[0]):
        return f"  {name}\n    CONSTANT  {a.reshape(-1)[0]}\n"
    return f"  {name}\n    INTERNAL FACTOR 1.0 IPRN 0\n" + "\n".join(
        "      " + " ".join(repr(float(v)) for v in row) for row in a.reshape(-1, a.shape[-1])) + "\n"


def _disv_text(shape, delr, delc, top, botm, xoff=0.0, yoff=0.0):
    """the rectangular grid as a DISV package: vertices row by row from the top-left corner, cells clockwise
    from their top-left vertex (what flopy's structured-to-vertex conversion writes)"""
    nlay, nrow, ncol = shape
    delr = np.broadcast_to(np.asarray(delr, dtype=float), (ncol,))
    delc = np.broadcast_to(np.asarray(delc, dtype=float), (nrow,))
    xe = xoff + np.concatenate([[0.0], np.cumsum(delr)])
    ye = yoff + delc.sum() - np.concatenate([[0.0], np.cumsum(delc)])
    s = (f"BEGIN options\nEND options\n\nBEGIN dimensions\n  NLAY {nlay}\n  NCPL {nrow * ncol}\n"
         f"  NVERT {(nrow + 1) * (ncol + 1)}\nEND dimensions\n\nBEGIN griddata\n" + _arr("top", top)
         + _arr("botm", botm, layered=np.ndim(botm) > 0) + "END griddata\n\nBEGIN vertices\n")
    for i in range(nrow + 1):
        for j in range(ncol + 1):
            s += f"  {i * (ncol + 1) + j + 1}  {float(xe[j])!r}  {float(ye[i])!r}\n"
    s += "END vertices\n\nBEGIN cell2d\n"
    for i in range(nrow):
        for j in range(ncol):
            v = [i * (ncol + 1) + j, i * (ncol + 1) + j + 1, (i + 1) * (ncol + 1) + j + 1, (i + 1) * (ncol + 1) + j]
            s += (f"  {i * ncol + j + 1}  {float(0.5 * (xe[j] + xe[j + 1]))!r}  {float(0.5 * (ye[i] + ye[i + 1]))!r}  4  "
                  + "  ".join(str(x + 1) for x in v) + "\n")
    return s + "END cell2d\n"


def _disu_text(shape, delr, delc, top, botm):
    """the rectangular grid as a DISU package: per-node top / bot / area and the CONNECTIONDATA arrays (iac, ja with
    the cell itself first, ihc, cl12, hwva) in the row order up, back, left, right, front, down"""
    nlay, nrow, ncol = shape
    delr = np.broadcast_to(np.asarray(delr, dtype=float), (ncol,))
    delc = np.broadcast_to(np.asarray(delc, dtype=float), (nrow,))
    bot = np.broadcast_to(np.asarray(botm, dtype=float)[:, None, None], shape)
    tp = np.concatenate([np.broadcast_to(np.asarray(top, dtype=float), (1, nrow, ncol)), bot[:-1]])
    node = lambda k, i, j: (k * nrow + i) * ncol + j   # noqa: E731
    iac, ja, ihc, cl12, hw = [], [], [], [], []
    for k in range(nlay):
        for i in range(nrow):
            for j in range(ncol):
                e = [(node(k, i, j), 0, 0.0, 0.0)]
                dz = 0.5 * (tp[k, i, j] - bot[k, i, j])
                if k > 0:
                    e.append((node(k - 1, i, j), 0, dz, delr[j] * delc[i]))
                if i > 0:
                    e.append((node(k, i - 1, j), 1, 0.5 * delc[i], delr[j]))
                if j > 0:
                    e.append((node(k, i, j - 1), 1, 0.5 * delr[j], delc[i]))
                if j < ncol - 1:
                    e.append((node(k, i, j + 1), 1, 0.5 * delr[j], delc[i]))
                if i < nrow - 1:
                    e.append((node(k, i + 1, j), 1, 0.5 * delc[i], delr[j]))
                if k < nlay - 1:
                    e.append((node(k + 1, i, j), 0, dz, delr[j] * delc[i]))
                iac.append(len(e))
                for m, h, c, w in e:
                    ja.append(m + 1); ihc.append(h); cl12.append(c); hw.append(w)
    n = nlay * nrow * ncol
    one = lambda name, a, f=repr: f"  {name}\n    INTERNAL FACTOR 1\n      " + " ".join(f(v) for v in a) + "\n"   # noqa: E731
    area = np.broadcast_to(delc[:, None] * delr[None, :], shape)
    return (f"BEGIN options\nEND options\n\nBEGIN dimensions\n  NODES {n}\n  NJA {len(ja)}\nEND dimensions\n\n"
            "BEGIN griddata\n" + one("top", [float(v) for v in tp.ravel()]) + one("bot", [float(v) for v in bot.ravel()])
            + one("area", [float(v) for v in area.ravel()]) + "END griddata\n\nBEGIN connectiondata\n"
            + one("iac", iac, str) + one("ja", ja, str) + one("ihc", ihc, str) + one("cl12", [float(v) for v in cl12])
            + one("hwva", [float(v) for v in hw]) + "END connectiondata\n")


def write_gwf(d, name, shape, delr, delc, top, botm, k, chd=None, wel=None, icelltype=0, strt=0.0, k33=None,
              sto=None, oc=True, newton=False, disv=False, extra_packages=(), disu=False, idomain=None):
    """chd / wel: dict iper -> list of ((k,i,j), value) with 1-based cellids.  disv=True writes the same
    rectangular grid as a DISV package (cellids become layer, icell2d)"""
    nlay, nrow, ncol = shape
    pk = f"  DIS{'V' if disv else ('U' if disu else '')}6  {name}.dis  dis\n  IC6  {name}.ic  ic\n  NPF6  {name}.npf  npf\n"
    if disu:
        _w(os.path.join(d, f"{name}.dis"), _disu_text(shape, delr, delc, top, botm))
    elif disv:
        _w(os.path.join(d, f"{name}.dis"), _disv_text(shape, delr, delc, top, botm))
    else:
        _w(os.path.join(d, f"{name}.dis"),
           f"BEGIN options\nEND options\n\nBEGIN dimensions\n  NLAY {nlay}\n  NROW {nrow}\n  NCOL {ncol}\nEND dimensions\n\n"
           "BEGIN griddata\n" + _arr("delr", delr) + _arr("delc", delc) + _arr("top", top)
           + _arr("botm", botm, layered=np.ndim(botm) > 0)
           + ("" if idomain is None else _arr("idomain", np.asarray(idomain, dtype=float), layered=True))
           + "END griddata\n")
    _w(os.path.join(d, f"{name}.ic"), "BEGIN griddata\n" + _arr("strt", strt) + "END griddata\n")
    npf = "BEGIN options\n  SAVE_FLOWS\nEND options\n\nBEGIN griddata\n" + _arr("icelltype", icelltype) \
        + _arr("k", np.asarray(k, dtype=float).reshape(shape) if np.ndim(k) else k, layered=np.ndim(k) > 0 and not disu)
    if k33 is not None:
        npf += _arr("k33", k33)
    _w(os.path.join(d, f"{name}.npf"), npf + "END griddata\n")
    if sto:
        pk += f"  STO6  {name}.sto  sto\n"
        s = "BEGIN options\nEND options\n\nBEGIN griddata\n" + _arr("iconvert", sto.get("iconvert", 0)) \
            + _arr("ss", sto["ss"]) + _arr("sy", sto.get("sy", 0.0)) + "END griddata\n\n"
        for iper, key in sorted(sto["periods"].items()):
            s += f"BEGIN period {iper}\n  {key}\nEND period {iper}\n\n"
        _w(os.path.join(d, f"{name}.sto"), s)
    for ft, spd in (("CHD", chd), ("WEL", wel)):
        if not spd:
            continue
        pk += f"  {ft}6  {name}.{ft.lower()}  {ft.lower()}_0\n"
        s = f"BEGIN options\nEND options\n\nBEGIN dimensions\n  MAXBOUND  {max(len(v) for v in spd.values())}\nEND dimensions\n\n"
        for iper, rows in sorted(spd.items()):
            cid = (lambda c: f"{c[0]} {(c[1] - 1) * ncol + c[2]}") if disv else (lambda c: f"{c[0]} {c[1]} {c[2]}")
            if disu:
                cid = lambda c: f"{((c[0] - 1) * nrow + c[1] - 1) * ncol + c[2]}"   # noqa: E731
            s += f"BEGIN period  {iper}\n" + "".join(f"  {cid(c)}  {v!r}\n" for c, v in rows) \
                + f"END period  {iper}\n\n"
        _w(os.path.join(d, f"{name}.{ft.lower()}"), s)
    for ft, ext, text in extra_packages:          # (ftype, file extension, file text)
        pk += f"  {ft}  {name}.{ext}  {ext}_0\n"
        _w(os.path.join(d, f"{name}.{ext}"), text)
    if oc:
        pk += f"  OC6  {name}.oc  oc\n"
        _w(os.path.join(d, f"{name}.oc"),
           f"BEGIN options\n  HEAD FILEOUT {name}.hds\n  BUDGET FILEOUT {name}.cbc\nEND options\n\n"
           "BEGIN period 1\n  SAVE HEAD ALL\n  SAVE BUDGET LAST\n  PRINT HEAD LAST\nEND period 1\n")
    _w(os.path.join(d, f"{name}.nam"),
       "BEGIN options\n" + ("  NEWTON UNDER_RELAXATION\n" if newton else "") + "END options\n\nBEGIN packages\n" + pk
       + "END packages\n")


def write_sim(d, models, perioddata, ims_text, exchanges=()):
    """models: [name]; exchanges: [(file stem, m1, m2, rows)] with rows = (cellid1, cellid2, ihc, cl1, cl2, hwva)"""
    _w(os.path.join(d, "sim.tdis"),
       f"BEGIN options\n  TIME_UNITS days\nEND options\n\nBEGIN dimensions\n  NPER {len(perioddata)}\nEND dimensions\n\n"
       "BEGIN perioddata\n" + "".join(f"  {p[0]!r}  {p[1]}  {p[2]!r}\n" for p in perioddata) + "END perioddata\n")
    _w(os.path.join(d, "sim.ims"), ims_text)
    ex = ""
    for stem, m1, m2, rows in exchanges:
        ex += f"  GWF6-GWF6  {stem}.gwfgwf  {m1}  {m2}\n"
        _w(os.path.join(d, f"{stem}.gwfgwf"),
           f"BEGIN options\n  AUXILIARY ANGLDEGX CDIST\n  SAVE_FLOWS\nEND options\n\nBEGIN dimensions\n  NEXG {len(rows)}\nEND dimensions\n\n"
           "BEGIN exchangedata\n" + "".join(
               f"  {a[0]} {a[1]} {a[2]}  {b[0]} {b[1]} {b[2]}  {ihc}  {c1!r}  {c2!r}  {hw!r}  0.0  {c1 + c2!r}\n"
               for a, b, ihc, c1, c2, hw in rows) + "END exchangedata\n")
    _w(os.path.join(d, "mfsim.nam"),
       "BEGIN options\nEND options\n\nBEGIN timing\n  TDIS6  sim.tdis\nEND timing\n\nBEGIN models\n"
       + "".join(f"  gwf6  {m}.nam  {m}\n" for m in models) + "END models\n\nBEGIN exchanges\n" + ex
       + "END exchanges\n\nBEGIN solutiongroup  1\n  ims6  sim.ims  " + "  ".join(models) + "\nEND solutiongroup  1\n")


IMS_PAR_GWF01 = """BEGIN options
  PRINT_OPTION  all
END options

BEGIN nonlinear
  OUTER_DVCLOSE  1.0E-08
  OUTER_MAXIMUM  100
  UNDER_RELAXATION  dbd
END nonlinear

BEGIN linear
  INNER_MAXIMUM  300
  INNER_DVCLOSE  1.0E-08
  inner_rclose   0.001
  LINEAR_ACCELERATION  bicgstab
  RELAXATION_FACTOR    0.97
END linear
"""


def write_par_gwf01(d, shape):
    """autotest/test_par_gwf01.py: two models of `shape` side by side, CHD 1.0 on the left edge, 10.0 on the
    right edge, K = 1, confined; known answer: heads 1, 2, ..., 10 along the columns"""
    nlay, nrow, ncol = shape
    botm = [-100.0 * (k + 1) for k in range(nlay)]
    left = [((k + 1, i + 1, 1), 1.0) for k in range(nlay) for i in range(nrow)]
    right = [((k + 1, i + 1, ncol), 10.0) for k in range(nlay) for i in range(nrow)]
    write_gwf(d, "leftmodel", shape, 100.0, 100.0, 0.0, botm, 1.0, chd={1: left})
    write_gwf(d, "rightmodel", shape, 100.0, 100.0, 0.0, botm, 1.0, chd={1: right})
    rows = [((k + 1, i + 1, ncol), (k + 1, i + 1, 1), 1, 50.0, 50.0, 100.0) for k in range(nlay) for i in range(nrow)]
    write_sim(d, ["leftmodel", "rightmodel"], [(1.0, 1, 1.0)], IMS_PAR_GWF01,
              exchanges=[("sim", "leftmodel", "rightmodel", rows)])
