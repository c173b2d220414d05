"""Binary head (.hds) and cell-by-cell budget (.cbc) files in the reference's formats (SURVEY.md section 8f,
rank 1: the step right after the hot path of every time step).

MODFLOW 6 opens its binary output with ACCESS='STREAM', FORM='UNFORMATTED' (src/Utilities/OpenSpec.f90):
plain little-endian values back to back, no record markers.  All integers are i32, all reals f64, all
strings 16 characters: right-justified literals for the flow terms, the left-justified variable name for heads.

Writers restate
  ulasav    src/Utilities/InputOutput.f90:924-940     one layer of a dependent variable
  ubdsv1    src/Utilities/InputOutput.f90:945-974     full array budget term (IMETH = 1)
  ubdsv06   src/Utilities/InputOutput.f90:981-1024    list budget term header (IMETH = 6)
  ubdsvd    src/Utilities/InputOutput.f90:1053-1071   one list entry (node, node2, q)
as they are driven by record_array (Dis.f90:1401-1485), record_connection_array
(DiscretizationBase.f90:955-969), sto_save_model_flows (gwf-sto.f90:620-640),
save_print_model_flows (BoundaryPackage.f90:1753-1900) and gwf_ot_flow (gwf.f90:911-960: STO-SS, STO-SY,
FLOW-JA-FACE, then the stress packages in package order).
Readers restate HeadFileReader.f90:88-130 and BudgetFileReader.f90:120-230 and are used by the tests to
check the writers (round trip + hand-packed records).
"""
import struct

import numpy as np

from . import ctypes_types as T

PKG_TEXT = {T.PKG_CHD: "CHD", T.PKG_WEL: "WEL", T.PKG_RIV: "RIV", T.PKG_RCH: "RCH", T.PKG_GHB: "GHB",
            T.PKG_DRN: "DRN"}


def _text16(s, right=True):
    s = s[:16]
    return (s.rjust(16) if right else s.ljust(16)).encode("ascii")


def grid_shape(shape):
    """(nlay, nrow, ncol) for DIS; (nlay, ncpl) for DISV is written as (nlay, 1, ncpl) (Disv.f90 mshape)"""
    if len(shape) == 3:
        return tuple(int(v) for v in shape)
    if len(shape) == 2:
        return int(shape[0]), 1, int(shape[1])
    return 1, 1, int(shape[0])


class HeadFileWriter:
    def __init__(self, path, shape, text="HEAD"):
        self.nlay, self.nrow, self.ncol = grid_shape(shape)
        # OutputControlData keeps the variable name in a character(len=16) (`this%cname = cname`,
        # OutputControlData.f90:22,162): LEFT-justified, unlike the right-justified literals of the budget terms
        # (checked against the reference's own autotest/data/ex-gwf-bump/results.hds.cmp)
        self.text = _text16(text, right=False)
        self.f = open(path, "wb")

    def write(self, kstp, kper, pertim, totim, head):
        """one ulasav record per layer"""
        head = np.ascontiguousarray(head, dtype="<f8").reshape(self.nlay, self.nrow * self.ncol)
        for k in range(self.nlay):
            self.f.write(struct.pack("<iidd", kstp, kper, pertim, totim) + self.text
                         + struct.pack("<iii", self.ncol, self.nrow, k + 1))
            self.f.write(head[k].tobytes())
        self.f.flush()

    def close(self):
        self.f.close()


class BudgetFileWriter:
    def __init__(self, path, shape, model_name):
        self.nlay, self.nrow, self.ncol = grid_shape(shape)
        self.model = _text16(model_name.upper(), right=False)
        self.f = open(path, "wb")

    def _header(self, kstp, kper, text, ncol, nrow, nlay, imeth, delt, pertim, totim):
        self.f.write(struct.pack("<ii", kstp, kper) + _text16(text) + struct.pack("<iii", ncol, nrow, -nlay))
        self.f.write(struct.pack("<iddd", imeth, delt, pertim, totim))

    def write_flowja(self, kstp, kper, delt, pertim, totim, flowja):
        """record_connection_array: ubdsv1 with ncol = nja, nrow = nlay = 1"""
        flowja = np.ascontiguousarray(flowja, dtype="<f8")
        self._header(kstp, kper, "FLOW-JA-FACE", flowja.size, 1, 1, 1, delt, pertim, totim)
        self.f.write(flowja.tobytes())

    def write_array(self, kstp, kper, delt, pertim, totim, text, values):
        """record_array with idataun < 0 (STO-SS, STO-SY): the whole grid as one IMETH = 1 record"""
        values = np.ascontiguousarray(values, dtype="<f8")
        assert values.size == self.nlay * self.nrow * self.ncol
        self._header(kstp, kper, text, self.ncol, self.nrow, self.nlay, 1, delt, pertim, totim)
        self.f.write(values.tobytes())

    def write_list(self, kstp, kper, delt, pertim, totim, text, package_name, nodes, q, auxname=(), aux=None,
                   bound_index=None):
        """save_print_model_flows (BoundaryPackage.f90 / BudgetObject ubdsv06): IMETH = 6 header, the auxiliary
        variable names, then (node, bound index, rate, auxiliary values) per boundary.
        `nodes` are 0-based cell numbers; entries with node < 0 (inactive) are skipped like node <= 0 there."""
        nodes = np.asarray(nodes)
        q = np.asarray(q, dtype=np.float64)
        keep = nodes >= 0
        naux = len(auxname)
        self._header(kstp, kper, text, self.ncol, self.nrow, self.nlay, 6, delt, pertim, totim)
        self.f.write(self.model + _text16(package_name.upper(), right=False) + self.model
                     + _text16(package_name.upper(), right=False))
        self.f.write(struct.pack("<i", naux + 1))
        for a in auxname:
            self.f.write(_text16(a.upper(), right=False))
        self.f.write(struct.pack("<i", int(keep.sum())))        # nlist
        rec = np.empty(int(keep.sum()), dtype=np.dtype([("n", "<i4"), ("n2", "<i4"), ("q", "<f8"),
                                                        ("aux", "<f8", (naux,))]))
        rec["n"] = nodes[keep] + 1
        rec["n2"] = (np.nonzero(keep)[0] if bound_index is None else np.asarray(bound_index)[keep]) + 1
        rec["q"] = q[keep]
        if naux:
            rec["aux"] = np.asarray(aux, dtype=np.float64).reshape(nodes.size, naux)[keep]
        self.f.write(rec.tobytes())

    def write_exchange(self, kstp, kper, delt, pertim, totim, exchange_name, other_model, nodes, other_nodes, q,
                       auxname=(), aux=None):
        """gwf_gwf_bdsav_model (exg-gwfgwf.f90:997-1153): FLOW-JA-FACE list between this model and `other_model`,
        source and destination package = the exchange name, one entry per exchange connection
        (node here, node there, rate into this model, auxiliary values); 0-based cell numbers in"""
        nodes, other_nodes = np.asarray(nodes), np.asarray(other_nodes)
        naux = len(auxname)
        self._header(kstp, kper, "FLOW-JA-FACE", self.ncol, self.nrow, self.nlay, 6, delt, pertim, totim)
        self.f.write(self.model + _text16(exchange_name.upper(), right=False)
                     + _text16(other_model.upper(), right=False) + _text16(exchange_name.upper(), right=False))
        self.f.write(struct.pack("<i", naux + 1))
        for a in auxname:
            self.f.write(_text16(a.upper(), right=False))
        self.f.write(struct.pack("<i", nodes.size))
        rec = np.empty(nodes.size, dtype=np.dtype([("n", "<i4"), ("n2", "<i4"), ("q", "<f8"), ("aux", "<f8", (naux,))]))
        rec["n"] = nodes + 1
        rec["n2"] = other_nodes + 1
        rec["q"] = q
        if naux:
            rec["aux"] = np.asarray(aux, dtype=np.float64).reshape(nodes.size, naux)
        self.f.write(rec.tobytes())
        self.f.flush()

    def write_step(self, kstp, kper, delt, pertim, totim, solution, packages, package_names=None, nodeuser=None):
        """everything gwf_ot_flow saves for one time step, taken from a solution object
        (GpuNumericalSolution or the oracle: flowja, storage_rates, simvals).  nodeuser: reduced -> user node of
        a grid with IDOMAIN holes -- list records carry USER node numbers, arrays are expanded to the user grid"""
        m = solution.model
        if getattr(m, "insto", 0):
            ss, sy = solution.storage_rates
            if nodeuser is not None:          # record_array fills the removed cells with 0 (dinact of sto_save_model_flows)
                full = np.zeros((2, self.nlay * self.nrow * self.ncol))
                full[0, nodeuser], full[1, nodeuser] = ss, sy
                ss, sy = full
            self.write_array(kstp, kper, delt, pertim, totim, "STO-SS", ss)
            if m.iconvert is not None and np.any(m.iconvert):      # iusesy (gwf-sto.f90:633-637)
                self.write_array(kstp, kper, delt, pertim, totim, "STO-SY", sy)
        self.write_flowja(kstp, kper, delt, pertim, totim, solution.flowja)
        sim = solution.simvals
        eff = getattr(solution, "effective_nodes", None)     # RCH: rch_cf resets nodelist to the highest active cell
        count = {}
        for i, p in enumerate(packages):
            t = PKG_TEXT[p.type]
            count[t] = count.get(t, 0) + 1
            name = package_names[i] if package_names else f"{t}-{count[t]}"   # default package names, e.g. CHD-1
            nodes = p.nodelist if eff is None else eff[i]
            self.write_list(kstp, kper, delt, pertim, totim, t, name, nodes if nodeuser is None else nodeuser[nodes],
                            sim[i], auxname=getattr(p, "auxnames", ()) or (), aux=getattr(p, "aux", None),
                            bound_index=getattr(p, "bound_index", None))
        self.f.flush()

    def close(self):
        self.f.close()


def fortran_g(x, w, d):
    """Fortran `1P, Gw.d` editing (what tdis_ot prints its times with): F editing with d significant digits and four
    trailing blanks while 0.1 <= |x| < 10**d, else 1PEw.d (one digit before the point, d after)"""
    ax = abs(x)
    if ax == 0.0:
        return f"{x:{w - 4}.{d - 1}f}" + "    "
    if 0.1 - 0.5 * 10.0 ** (-d - 1) <= ax < 10.0 ** d - 0.5:
        n = 0
        while ax >= 10.0 ** n - 0.5 * 10.0 ** (n - d):
            n += 1
        return f"{x:#{w - 4}.{d - n}f}" + "    "
    return f"{x:{w}.{d}E}"


def _budget_value(v, big):
    """Budget.f90 value_to_string :150-170"""
    a = abs(v)
    if v != 0.0 and (a >= big or a < 0.1):
        return f"{v:17.4E}"
    return f"{v:17.4f}"


class ListingFileWriter:
    """The part of a model listing file (.lst) that post-processors parse: the VOLUME BUDGET table of `budget_ot`
    (Budget.f90:178-311, labelled variant: formats 261 / 266 / 276 / 286 / 287 / 298 / 299 / 300) followed by the TIME
    SUMMARY of `tdis_ot` (tdis.f90:273-340), written for every time step whose budget OC asks to PRINT.  Entries are
    (text, rate in, rate out, package name); the cumulative volumes accumulate rate * delt like `addentry`."""

    _SECONDS = {"SECONDS": 1.0, "MINUTES": 60.0, "HOURS": 3600.0, "DAYS": 86400.0, "YEARS": 31557600.0}

    def __init__(self, path, model_name, time_units=None):
        self.f = open(path, "w")
        self.f.write(f"                                   MODFLOW 6 (mf6-b200 GPU path)\n\n"
                     f" GROUNDWATER FLOW MODEL ({model_name.upper()})\n")
        self.cnv = self._SECONDS.get((time_units or "").upper(), 0.0)
        self.vol = {}

    def write_budget(self, kstp, kper, delt, pertim, totim, entries):
        w = self.f.write
        rows = []
        for i, (text, rin, rout, label) in enumerate(entries):
            vin, vout = self.vol.get((i, text), (0.0, 0.0))
            vin, vout = vin + rin * delt, vout + rout * delt
            self.vol[(i, text)] = (vin, vout)
            rows.append((f"{text.upper():>16s}"[:16], vin, vout, rin, rout, label))
        totrin, totrot = sum(r[3] for r in rows), sum(r[4] for r in rows)
        totvin, totvot = sum(r[1] for r in rows), sum(r[2] for r in rows)
        big1, big2 = 9.99999e11, 9.99999e10
        w(f"\n\n  VOLUME BUDGET FOR ENTIRE MODEL AT END OF TIME STEP{kstp:5d}, STRESS PERIOD{kper:4d}\n  " + 99 * "-" + "\n")
        w(" \n     CUMULATIVE VOLUME      L**3       RATES FOR THIS TIME STEP      L**3/T          "
          + f"{'PACKAGE NAME':<16s}\n     " + 18 * "-" + 17 * " " + 24 * "-" + 21 * " " + 16 * "-" + "\n\n"
          + 11 * " " + "IN:" + 38 * " " + "IN:\n" + 11 * " " + "---" + 38 * " " + "---\n")
        for nm, vin, _, rin, _, label in rows:
            w(f"    {nm} ={_budget_value(vin, big1)}      {nm} ={_budget_value(rin, big1)}     {label}\n")
        w(" \n" + 12 * " " + f"TOTAL IN ={_budget_value(totvin, big1)}" + 14 * " " + f"TOTAL IN ={_budget_value(totrin, big1)}\n")
        w(" \n" + 10 * " " + "OUT:" + 37 * " " + "OUT:\n" + 10 * " " + "----" + 37 * " " + "----\n")
        for nm, _, vout, _, rout, label in rows:
            w(f"    {nm} ={_budget_value(vout, big1)}      {nm} ={_budget_value(rout, big1)}     {label}\n")
        w(" \n" + 11 * " " + f"TOTAL OUT ={_budget_value(totvot, big1)}" + 13 * " " + f"TOTAL OUT ={_budget_value(totrot, big1)}\n")
        diffr, diffv = totrin - totrot, totvin - totvot
        pdiffr = 100.0 * diffr / ((totrin + totrot) / 2.0) if (totrin + totrot) != 0.0 else 0.0
        pdiffv = 100.0 * diffv / ((totvin + totvot) / 2.0) if (totvin + totvot) != 0.0 else 0.0
        w(" \n" + 12 * " " + f"IN - OUT ={_budget_value(diffv, big2)}" + 14 * " " + f"IN - OUT ={_budget_value(diffr, big2)}\n")
        w(f" \n PERCENT DISCREPANCY ={pdiffv:15.2f}     PERCENT DISCREPANCY ={pdiffr:15.2f}\n\n")
        # tdis_ot
        w(" \n\n\n" + 9 * " " + f"TIME SUMMARY AT END OF TIME STEP{kstp:5d} IN STRESS PERIOD {kper:4d}\n")
        if self.cnv == 0.0:
            w(21 * " " + f"     TIME STEP LENGTH ={delt:15.6G}\n" + 21 * " " + f"   STRESS PERIOD TIME ={pertim:15.6G}\n"
              + 21 * " " + f"TOTAL SIMULATION TIME ={totim:15.6G}\n")
        else:
            w(19 * " " + " SECONDS     MINUTES      HOURS" + 7 * " " + "DAYS        YEARS\n" + 20 * " " + 59 * "-" + "\n")
            for label, t in (("  TIME STEP LENGTH", delt), ("STRESS PERIOD TIME", pertim), ("        TOTAL TIME", totim)):
                sec = self.cnv * t
                vals = (sec, sec / 60.0, sec / 3600.0, sec / 86400.0, sec / 86400.0 / 365.25)
                w(" " + label + "".join(fortran_g(v, 12, 5) for v in vals) + "\n")
            w("\n")
        self.f.flush()
        return pdiffr

    def close(self):
        self.f.close()


class BudgetCsvWriter:
    """OC `BUDGETCSV FILEOUT <file>`: one line per time step (Budget.f90 writecsv :671-718, header :726-757):
    time, every entry's inflow rate, every entry's outflow rate, TOTAL_IN, TOTAL_OUT, PERCENT_DIFFERENCE"""

    def __init__(self, path):
        self.f = open(path, "w")
        self.header = False

    def write(self, totim, entries):
        if not self.header:
            names = [f"{t.strip().upper()}({lab.strip()})" for t, _, _, lab in entries]
            self.f.write("time," + "".join(n + "_IN," for n in names) + "".join(n + "_OUT," for n in names)
                         + "TOTAL_IN,TOTAL_OUT,PERCENT_DIFFERENCE\n")
            self.header = True
        rin, rout = [e[1] for e in entries], [e[2] for e in entries]
        tin, tout = sum(rin), sum(rout)
        pd = 100.0 * (tin - tout) / ((tin + tout) / 2.0) if (tin + tout) != 0.0 else 0.0
        self.f.write(",".join(repr(float(v)) for v in [totim] + rin + rout + [tin, tout, pd]) + "\n")
        self.f.flush()

    def close(self):
        self.f.close()


def read_listing_budgets(path):
    """the budget tables of a listing file -> list of dicts(kstp, kper, totim, IN / OUT rates and cumulative volumes
    per (text, package), totals, percent discrepancy): the parsing a list-budget post-processor does"""
    out, cur, side = [], None, None
    with open(path) as f:
        for line in f:
            if "BUDGET FOR ENTIRE MODEL AT END OF TIME STEP" in line:
                t = line.replace(",", " ").split()
                cur = dict(kstp=int(t[t.index("STEP") + 1]), kper=int(t[t.index("PERIOD") + 1]), rates_in={},
                           rates_out={}, volumes_in={}, volumes_out={})
                out.append(cur)
                side = None
            elif cur is None:
                continue
            elif line.strip().startswith("IN:"):
                side = "in"
            elif line.strip().startswith("OUT:"):
                side = "out"
            elif "TOTAL IN =" in line or "TOTAL OUT =" in line:
                v = line.split("=")
                key = "total_in" if "TOTAL IN" in line else "total_out"
                cur[key + "_volume"], cur[key] = float(v[1].split()[0]), float(v[2].split()[0])
            elif "PERCENT DISCREPANCY =" in line:
                cur["pdiffr"] = float(line.split("=")[2].split()[0])
            elif "TOTAL TIME" in line and "totim" not in cur:
                cur["totim_seconds"] = float(line.split("TIME")[1].split()[0])
            elif "TOTAL SIMULATION TIME =" in line:
                cur["totim"] = float(line.split("=")[1])
            elif side and line.count("=") == 2:
                a, b, c = line.split("=")
                name, vol, rate, label = a.strip(), float(b.split()[0]), float(c.split()[0]), " ".join(c.split()[1:])
                cur["volumes_" + side][(name, label)] = vol
                cur["rates_" + side][(name, label)] = rate
    return out


def _grb_line(text, n):
    return (text[:n - 1].ljust(n - 1) + "\n").encode("ascii")


def write_grb(path, grid, model, nodeuser=None):
    """Binary grid file, what post-processors need to interpret FLOW-JA-FACE (`write_grb` of Dis.f90:547-659,
    Disv.f90:709-850, Disu.f90:908-1030): four 50-byte header lines (GRID <type>, VERSION 1, NTXT, LENTXT 100), one
    100-byte definition line per variable, then the data in that order.  IA / JA are the connectivity in USER
    numbering (`iajausr`, Connections.f90:1112-1160: a removed cell has an empty row), 1-based, every row = the cell
    itself followed by its neighbours.  `grid` is GwfInput.grid, `model` the (reduced) GwfModel."""
    kind = grid["kind"]
    nodesuser = int(grid["idomain"].size)
    nu = np.arange(nodesuser) if nodeuser is None else np.asarray(nodeuser)
    cnt = np.zeros(nodesuser, dtype=np.int64)
    cnt[nu] = np.diff(model.ia)
    iausr = (1 + np.concatenate([[0], np.cumsum(cnt)])).astype("<i4")
    jausr = (nu[model.ja] + 1).astype("<i4")
    nja = int(model.nja)
    f8 = lambda a: np.ascontiguousarray(a, dtype="<f8").tobytes()      # noqa: E731
    i4 = lambda a: np.ascontiguousarray(a, dtype="<i4").tobytes()      # noqa: E731
    g25 = lambda v: f"{v:24.15E}"                                      # noqa: E731
    defs, data = [], []

    def scalar_i(name, v):
        defs.append(f"{name} INTEGER NDIM 0 # {int(v)}")
        data.append(struct.pack("<i", int(v)))

    def scalar_d(name, v):
        defs.append(f"{name} DOUBLE NDIM 0 # {g25(float(v))}")
        data.append(struct.pack("<d", float(v)))

    def arr_d(name, a, dims=None):
        a = np.asarray(a, dtype=np.float64)
        defs.append(f"{name} DOUBLE NDIM {dims or ('1 %d' % a.size)}")
        data.append(f8(a))

    def arr_i(name, a):
        a = np.asarray(a)
        defs.append(f"{name} INTEGER NDIM 1 {a.size}")
        data.append(i4(a))

    if kind == "DIS":
        scalar_i("NCELLS", nodesuser); scalar_i("NLAY", grid["nlay"]); scalar_i("NROW", grid["nrow"])
        scalar_i("NCOL", grid["ncol"]); scalar_i("NJA", nja)
        scalar_d("XORIGIN", grid["xorigin"]); scalar_d("YORIGIN", grid["yorigin"]); scalar_d("ANGROT", grid["angrot"])
        arr_d("DELR", grid["delr"]); arr_d("DELC", grid["delc"]); arr_d("TOP", grid["top"]); arr_d("BOTM", grid["botm"])
    elif kind == "DISV":
        verts = np.asarray(grid["vertices"], dtype=np.float64)
        iav, jav = [1], []
        for _, _, iv in grid["cells"]:
            iv = list(iv)
            if iv[0] != iv[-1]:
                iv.append(iv[0])             # the polygon is stored closed (Disv.f90 source_cell2d)
            jav += [v + 1 for v in iv]
            iav.append(len(jav) + 1)
        scalar_i("NCELLS", nodesuser); scalar_i("NLAY", grid["nlay"]); scalar_i("NCPL", grid["ncpl"])
        scalar_i("NVERT", verts.shape[0]); scalar_i("NJAVERT", len(jav)); scalar_i("NJA", nja)
        scalar_d("XORIGIN", grid["xorigin"]); scalar_d("YORIGIN", grid["yorigin"]); scalar_d("ANGROT", grid["angrot"])
        arr_d("TOP", grid["top"]); arr_d("BOTM", grid["botm"])
        arr_d("VERTICES", verts.reshape(-1), dims=f"2 2 {verts.shape[0]}")
        arr_d("CELLX", [c[0] for c in grid["cells"]]); arr_d("CELLY", [c[1] for c in grid["cells"]])
        arr_i("IAVERT", iav); arr_i("JAVERT", jav)
    elif kind == "DISU":
        scalar_i("NODES", nodesuser); scalar_i("NJA", nja)
        scalar_d("XORIGIN", grid["xorigin"]); scalar_d("YORIGIN", grid["yorigin"]); scalar_d("ANGROT", grid["angrot"])
        arr_d("TOP", grid["top"]); arr_d("BOT", grid["bot"])
    else:
        raise ValueError(f"write_grb: unknown grid kind {kind}")
    arr_i("IA", iausr); arr_i("JA", jausr); arr_i("IDOMAIN", grid["idomain"]); arr_i("ICELLTYPE", grid["icelltype"])
    with open(path, "wb") as f:
        for t in (f"GRID {kind}", "VERSION 1", f"NTXT {len(defs)}", "LENTXT 100"):
            f.write(_grb_line(t, 50))
        for t in defs:
            f.write(_grb_line(t, 100))
        for b in data:
            f.write(b)


def read_grb(path):
    """the self-describing binary grid file -> dict(GRID=..., NAME=value or array)"""
    out = {}
    with open(path, "rb") as f:
        hdr = [f.read(50).decode("ascii").strip() for _ in range(4)]
        out["GRID"] = hdr[0].split()[1]
        ntxt, lentxt = int(hdr[2].split()[1]), int(hdr[3].split()[1])
        defs = [f.read(lentxt).decode("ascii").split("#")[0].split() for _ in range(ntxt)]
        for d in defs:
            name, typ, ndim = d[0], d[1], int(d[3])
            n = int(np.prod([int(v) for v in d[4:4 + ndim]])) if ndim else 1
            dt = "<i4" if typ == "INTEGER" else "<f8"
            a = np.frombuffer(f.read(n * int(dt[-1])), dtype=dt)
            out[name] = a[0].item() if ndim == 0 else a.copy()
    return out


def read_head_file(path):
    """list of dicts (kstp, kper, pertim, totim, text, ncol, nrow, ilay, data[nrow, ncol])"""
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(52)
            if len(h) < 52:
                break
            kstp, kper, pertim, totim = struct.unpack("<iidd", h[:24])
            text = h[24:40].decode("ascii")
            ncol, nrow, ilay = struct.unpack("<iii", h[40:52])
            data = np.frombuffer(f.read(8 * ncol * nrow), dtype="<f8").reshape(nrow, ncol)
            out.append(dict(kstp=kstp, kper=kper, pertim=pertim, totim=totim, text=text, ncol=ncol, nrow=nrow,
                            ilay=ilay, data=data))
    return out


def read_budget_file(path):
    """list of dicts; IMETH 1: `flow` array (FLOW-JA-FACE: nval = ncol; else ncol*nrow*|nlay|),
    IMETH 6: names, auxtxt, `node`, `node2`, `q`, `aux`"""
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(36)
            if len(h) < 36:
                break
            kstp, kper = struct.unpack("<ii", h[:8])
            text = h[8:24].decode("ascii")
            nval, idum1, idum2 = struct.unpack("<iii", h[24:36])
            imeth, delt, pertim, totim = struct.unpack("<iddd", f.read(28))
            r = dict(kstp=kstp, kper=kper, text=text, ncol=nval, nrow=idum1, nlay=idum2, imeth=imeth, delt=delt,
                     pertim=pertim, totim=totim)
            if imeth == 1:
                n = nval if text.strip() == "FLOW-JA-FACE" else nval * idum1 * abs(idum2)
                r["flow"] = np.frombuffer(f.read(8 * n), dtype="<f8")
            elif imeth == 6:
                names = [f.read(16).decode("ascii") for _ in range(4)]
                r.update(srcmodel=names[0], srcpackage=names[1], dstmodel=names[2], dstpackage=names[3])
                ndat, = struct.unpack("<i", f.read(4))
                naux = ndat - 1
                r["auxtxt"] = [f.read(16).decode("ascii") for _ in range(naux)]
                nlist, = struct.unpack("<i", f.read(4))
                dt = np.dtype([("n", "<i4"), ("n2", "<i4"), ("q", "<f8"), ("aux", "<f8", (naux,))])
                rec = np.frombuffer(f.read(dt.itemsize * nlist), dtype=dt)
                r.update(node=rec["n"].copy(), node2=rec["n2"].copy(), q=rec["q"].copy(), aux=rec["aux"].copy())
            else:
                raise ValueError(f"budget file: IMETH {imeth} is not written by MODFLOW 6")
            out.append(r)
    return out
