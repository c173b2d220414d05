"""Reader for the subset of MODFLOW 6 input files that feeds the accelerated path (SURVEY.md section 8f,
rank 3): mfsim.nam, TDIS, IMS, GWF name file, DIS, DISV, DISU, IC, NPF, STO, CHD/WEL/DRN/RIV/GHB/RCH lists, array-based
RCH (READASARRAYS), OC and
GWF-GWF exchanges -- enough to run FloPy-written models such as the reference's `.mf6minsim/` example through
`mf6gpu_solution_*` without the Fortran host (which cannot be built in this image).

Format facts restated here (doc/mf6io/mf6ivar/dfn/*.dfn; src/Utilities/BlockParser.f90,
src/Utilities/ArrayReaders.f90; src/Utilities/Idm/mf6blockfile/*):
  * free format, case-insensitive keywords, tokens separated by blanks or commas, quotes group a token;
    `#`, `!` and `//` start a comment line; BEGIN <name> [<number>] ... END <name>
  * READARRAY control lines: CONSTANT v | INTERNAL [FACTOR f] [IPRN n] + values | OPEN/CLOSE file [FACTOR f]
    [(BINARY)] [IPRN n]; with LAYERED after the array name there is one control line per layer
  * list packages: PERIOD <iper> blocks hold `cellid  bound values...`; a period's list stays in force until
    the next PERIOD block (BoundaryPackage bnd_rp)
Anything outside the supported subset raises Mf6InputError naming the keyword -- nothing is ignored silently
except print/format options that do not change results.
"""
import os
import shlex
from dataclasses import dataclass, field

import numpy as np

from . import ctypes_types as T
from .disv import build_disv_model, cell2d_from_vertices
from .timeseries import TimeSeriesError, read_ts_file  # noqa: F401
from .grid import Package, build_dis_model, build_dis_model_idomain, build_disu_model, reduce_model


class Mf6InputError(ValueError):
    pass


# ---- block-file tokeniser -----------------------------------------------------------------------
def _tokens(line):
    s = line.strip()
    if not s or s[0] in "#!" or s.startswith("//"):
        return []
    lex = shlex.shlex(s.replace(",", " "), posix=True)
    lex.whitespace_split = True
    lex.commenters = "#"
    return list(lex)


def read_blocks(path):
    """-> list of (NAME, number or None, [token lists])"""
    if not os.path.exists(path):
        raise Mf6InputError(f"input file not found: {path}")
    blocks, cur = [], None
    with open(path) as f:
        for raw in f:
            t = _tokens(raw)
            if not t:
                continue
            k = t[0].upper()
            if k == "BEGIN":
                if cur is not None:
                    raise Mf6InputError(f"{path}: BEGIN {t[1]} inside block {cur[0]}")
                num = int(t[2]) if len(t) > 2 and t[2].lstrip("-").isdigit() else None
                cur = (t[1].upper(), num, [])
            elif k == "END":
                if cur is None or cur[0] != t[1].upper():
                    raise Mf6InputError(f"{path}: END {t[1]} does not close a block")
                blocks.append(cur)
                cur = None
            elif cur is not None:
                cur[2].append(t)
            else:
                raise Mf6InputError(f"{path}: text outside a block: {raw.strip()}")
    if cur is not None:
        raise Mf6InputError(f"{path}: block {cur[0]} is not closed")
    return blocks


def _block(blocks, name, required=True):
    for b in blocks:
        if b[0] == name:
            return b[2]
    if required:
        raise Mf6InputError(f"block {name} is missing")
    return []


def _options(lines):
    return {ln[0].upper(): ln[1:] for ln in lines}


# ---- READARRAY ----------------------------------------------------------------------------------
class _ArrayReader:
    def __init__(self, lines, base_dir):
        self.lines, self.pos, self.dir = lines, 0, base_dir

    def more(self):
        return self.pos < len(self.lines)

    def name_line(self):
        ln = self.lines[self.pos]
        self.pos += 1
        return ln[0].upper(), (len(ln) > 1 and ln[1].upper() == "LAYERED")

    def _one(self, n, dtype):
        ctl = self.lines[self.pos]
        self.pos += 1
        key = ctl[0].upper()
        conv = float if dtype == np.float64 else (lambda s: int(float(s)))
        if key == "CONSTANT":
            return np.full(n, conv(ctl[1]), dtype=dtype)
        factor = 1.0
        up = [c.upper() for c in ctl]
        if "FACTOR" in up:
            factor = float(ctl[up.index("FACTOR") + 1])
        if key == "INTERNAL":
            vals = []
            while len(vals) < n:
                if self.pos >= len(self.lines):
                    raise Mf6InputError("INTERNAL array is shorter than the grid")
                vals += self.lines[self.pos]
                self.pos += 1
            a = np.array([float(v) for v in vals[:n]])
        elif key == "OPEN/CLOSE":
            fn = os.path.join(self.dir, ctl[1])
            if "(BINARY)" in up:
                # records of a 52-byte header (kstp, kper, pertim, totim, text, m1, m2, m3) followed by m1 * m2
                # values, one record per layer of a 3-D array (ArrayReaders.f90 read_binary_header :1029-1067)
                item = 8 if dtype == np.float64 else 4
                raw = open(fn, "rb").read()
                vals, off = [], 0
                while sum(v.size for v in vals) < n:
                    if off + 52 > len(raw):
                        raise Mf6InputError(f"{fn}: binary array file is shorter than the grid")
                    m1, m2, _ = np.frombuffer(raw, dtype="<i4", count=3, offset=off + 40)
                    cnt = int(m1) * int(m2)
                    vals.append(np.frombuffer(raw, dtype="<f8" if item == 8 else "<i4", count=cnt, offset=off + 52))
                    off += 52 + cnt * item
                a = np.concatenate(vals)[:n].astype(np.float64)
            else:
                a = np.loadtxt(fn).reshape(-1)[:n].astype(np.float64)
        else:
            raise Mf6InputError(f"array control record {ctl[0]} is not supported")
        a = a * factor
        return a.astype(dtype) if dtype != np.float64 else a

    def read(self, shape, layered, dtype=np.float64):
        n = int(np.prod(shape))
        if layered and len(shape) == 3:
            per = shape[1] * shape[2]
            return np.concatenate([self._one(per, dtype) for _ in range(shape[0])])
        return self._one(n, dtype)


def read_griddata(lines, base_dir, spec):
    """spec: NAME -> (shape, dtype).  -> dict NAME -> flat array"""
    rd, out = _ArrayReader(lines, base_dir), {}
    while rd.more():
        name, layered = rd.name_line()
        if name not in spec:
            raise Mf6InputError(f"GRIDDATA array {name} is not supported on the GPU path")
        shape, dtype = spec[name]
        out[name] = rd.read(shape, layered, dtype)
    return out


# ---- simulation objects ---------------------------------------------------------------------------
@dataclass
class StressPackage:
    ftype: str
    name: str
    periods: dict            # iper (1-based) -> Package or None (empty period block)
    iflowred: int = 0
    flowred: float = 0.1
    from_arrays: bool = False     # READASARRAYS input: one boundary per 2-D cell, also over cells that do not exist


@dataclass
class GwfInput:
    name: str
    model: object            # grid.GwfModel
    shape: tuple
    packages: list = field(default_factory=list)      # [StressPackage]
    sto_transient: dict = field(default_factory=dict)  # iper -> bool (True = TRANSIENT)
    head_file: str = None
    budget_file: str = None
    save: dict = field(default_factory=dict)          # iper -> list of (rtype, ocsetting tokens)
    printrec: dict = field(default_factory=dict)      # the same for the OC PRINT records
    budgetcsv_file: str = None                        # OC BUDGETCSV FILEOUT
    list_file: str = None                             # model listing file (name file LIST option or <name file>.lst)
    grid: dict = None                                 # what the binary grid file records (output.write_grb)
    gnc: tuple = None                                 # GNC6: (noden, nodem, nodesj, alphasj), reduced 0-based nodes
    hfb: dict = field(default_factory=dict)           # iper -> (noden, nodem, hydchr), HFB6 barriers (0-based nodes)
    nodeuser: np.ndarray = None       # DIS with IDOMAIN <= 0 cells: reduced -> user node (model.nodes entries)
    nodereduced: np.ndarray = None    # user -> reduced node, -1 where no cell exists

    @property
    def nodesuser(self):
        return int(np.prod(self.shape))


@dataclass
class Simulation:
    base_dir: str
    nper: int
    perioddata: list         # [(perlen, nstp, tsmult)]
    models: list             # [GwfInput]
    exchanges: list          # dicts for grid.merge_models
    sln: object
    ims: object
    warnings: list = field(default_factory=list)
    time_units: str = None
    continue_: bool = False      # mfsim.nam CONTINUE: go on after a time step that did not converge


_PKG_TYPE = {"CHD6": T.PKG_CHD, "WEL6": T.PKG_WEL, "RIV6": T.PKG_RIV, "RCH6": T.PKG_RCH, "GHB6": T.PKG_GHB,
             "DRN6": T.PKG_DRN}
_PKG_NCOL = {"CHD6": 1, "WEL6": 1, "RIV6": 3, "RCH6": 1, "GHB6": 2, "DRN6": 2}


def read_tdis(path):
    b = read_blocks(path)
    if "ATS6" in _options(_block(b, "OPTIONS", required=False)):
        raise Mf6InputError(f"{path}: adaptive time stepping (ATS6) is not supported on the GPU path")
    nper = int(_options(_block(b, "DIMENSIONS"))["NPER"][0])
    pd = [(float(t[0]), int(t[1]), float(t[2])) for t in _block(b, "PERIODDATA")]
    if len(pd) != nper:
        raise Mf6InputError(f"{path}: PERIODDATA has {len(pd)} rows, NPER = {nper}")
    return nper, pd


def read_time_units(path):
    opt = _options(_block(read_blocks(path), "OPTIONS", required=False))
    return opt["TIME_UNITS"][0].upper() if opt.get("TIME_UNITS") else None


def read_ims(path, warnings):
    """IMS options -> (SlnSettings, ImsSettings).  Defaults and the COMPLEXITY presets follow
    NumericalSolution.f90:568-866 / sln_set_defaults and ImsLinearSettings.f90:120-253 (preset_config)."""
    b = read_blocks(path)
    opt = _options(_block(b, "OPTIONS", required=False))
    nl = _options(_block(b, "NONLINEAR", required=False))
    li = _options(_block(b, "LINEAR", required=False))
    cx = (opt.get("COMPLEXITY", ["SIMPLE"])[0]).upper()
    if cx not in ("SIMPLE", "MODERATE", "COMPLEX"):
        raise Mf6InputError(f"{path}: unknown COMPLEXITY {cx}")
    # presets: sln_setouter (NumericalSolution.f90:2623-2671), preset_config (ImsLinearSettings.f90:74-116)
    ci = ("SIMPLE", "MODERATE", "COMPLEX").index(cx)
    sln = dict(dvclose=(1e-3, 1e-2, 1e-1)[ci], mxiter=(25, 50, 100)[ci], nonmeth=(0, 3, 3)[ci],
               theta=(1.0, 0.9, 0.8)[ci], akappa=(0.0, 1e-4, 1e-4)[ci], gamma=(1.0, 0.0, 0.0)[ci], amomentum=0.0,
               numtrack=(0, 0, 20)[ci], btol=(0.0, 0.0, 1.05)[ci], breduc=(0.0, 0.0, 0.1)[ci],
               res_lim=(0.0, 0.0, 0.002)[ci])
    ims = dict(iter1=(50, 100, 500)[ci], ilinmeth=(1, 2, 2)[ci], dvclose=(1e-3, 1e-2, 1e-1)[ci], rclose=1e-1,
               relax=(0.0, 0.97, 0.0)[ci], level=(0, 0, 5)[ci], droptol=(0.0, 0.0, 1e-4)[ci], north=(0, 0, 2)[ci],
               iscl=0, iord=0, icnvgopt=0)
    iallowptc = 1
    if "NO_PTC" in opt:
        v = opt["NO_PTC"]
        iallowptc = -1 if (v and v[0].upper() == "FIRST") else 0
    ur = {"NONE": 0, "SIMPLE": 1, "COOLEY": 2, "DBD": 3}
    for k, v in nl.items():
        if k in ("OUTER_DVCLOSE", "OUTER_HCLOSE"):
            sln["dvclose"] = float(v[0])
        elif k == "OUTER_MAXIMUM":
            sln["mxiter"] = int(v[0])
        elif k == "UNDER_RELAXATION":
            sln["nonmeth"] = ur[v[0].upper()]
        elif k == "UNDER_RELAXATION_THETA":
            sln["theta"] = float(v[0])
        elif k == "UNDER_RELAXATION_KAPPA":
            sln["akappa"] = float(v[0])
        elif k == "UNDER_RELAXATION_GAMMA":
            sln["gamma"] = float(v[0])
        elif k == "UNDER_RELAXATION_MOMENTUM":
            sln["amomentum"] = float(v[0])
        elif k == "BACKTRACKING_NUMBER":
            sln["numtrack"] = int(v[0])
        elif k == "BACKTRACKING_TOLERANCE":
            sln["btol"] = float(v[0])
        elif k == "BACKTRACKING_REDUCTION_FACTOR":
            sln["breduc"] = float(v[0])
        elif k == "BACKTRACKING_RESIDUAL_LIMIT":
            sln["res_lim"] = float(v[0])
        elif k == "OUTER_RCLOSEBND":
            warnings.append("OUTER_RCLOSEBND is deprecated and ignored (as in the reference)")
        else:
            raise Mf6InputError(f"{path}: NONLINEAR option {k} is not supported")
    rc = {"STRICT": 1, "L2NORM_RCLOSE": 2, "RELATIVE_RCLOSE": 3, "L2NORM_RELATIVE_RCLOSE": 4}
    for k, v in li.items():
        if k == "INNER_MAXIMUM":
            ims["iter1"] = int(v[0])
        elif k in ("INNER_DVCLOSE", "INNER_HCLOSE"):
            ims["dvclose"] = float(v[0])
        elif k == "INNER_RCLOSE":
            ims["rclose"] = float(v[0])
            if len(v) > 1:
                ims["icnvgopt"] = rc[v[1].upper()]
        elif k == "LINEAR_ACCELERATION":
            ims["ilinmeth"] = {"CG": 1, "BICGSTAB": 2}[v[0].upper()]
        elif k == "RELAXATION_FACTOR":
            ims["relax"] = float(v[0])
        elif k == "PRECONDITIONER_LEVELS":
            ims["level"] = int(v[0])
        elif k == "PRECONDITIONER_DROP_TOLERANCE":
            ims["droptol"] = float(v[0])
        elif k == "NUMBER_ORTHOGONALIZATIONS":
            ims["north"] = int(v[0])
        elif k == "SCALING_METHOD":
            ims["iscl"] = {"NONE": 0, "DIAGONAL": 1, "L2NORM": 2}[v[0].upper()]
        elif k == "REORDERING_METHOD":
            ims["iord"] = {"NONE": 0, "RCM": 1, "MD": 2}[v[0].upper()]
        else:
            raise Mf6InputError(f"{path}: LINEAR option {k} is not supported")
    # the policy of petsc_check_settings (PetscSolver.F90:123-154): what the backend cannot honour is
    # downgraded with a warning the caller can print
    if ims["iord"] != 0:
        warnings.append("REORDERING_METHOD is replaced by the GPU level ordering")
        ims["iord"] = 0
    return T.SlnSettings.make(iallowptc=iallowptc, **sln), T.ImsSettings.make(**ims)


def _cellid(tokens, shape):
    if len(shape) == 1:       # DISU: node
        nd = int(tokens[0]) - 1
        if not 0 <= nd < shape[0]:
            raise Mf6InputError(f"cellid {tokens[0]} outside the grid of {shape[0]} nodes")
        return nd, 1
    if len(shape) == 2:       # DISV: (layer, icell2d)
        k, j = int(tokens[0]) - 1, int(tokens[1]) - 1
        if not (0 <= k < shape[0] and 0 <= j < shape[1]):
            raise Mf6InputError(f"cellid {tokens[:2]} outside the grid {shape}")
        return k * shape[1] + j, 2
    if len(shape) == 3:
        k, i, j = int(tokens[0]) - 1, int(tokens[1]) - 1, int(tokens[2]) - 1
        if not (0 <= k < shape[0] and 0 <= i < shape[1] and 0 <= j < shape[2]):
            raise Mf6InputError(f"cellid {tokens[:3]} outside the grid {shape}")
        return (k * shape[1] + i) * shape[2] + j, 3
    raise Mf6InputError("only DIS cellids are supported")


def _read_rcha(blocks, name, shape, fixed_cell=0, base_dir=""):
    """array-based recharge (gwf-rcha.dfn): PERIOD blocks hold IRCH (layer of every 2-D cell, default 1) and
    RECHARGE arrays; an array that a block omits keeps its previous values (rch_rp / RchType read_initial_attr).
    Turned into the equivalent list: one boundary per 2-D cell at (irch, cell)."""
    ncpl = int(np.prod(shape[1:]))
    a2 = (1,) + tuple(shape[1:]) if len(shape) == 3 else (1, 1, shape[1])
    irch = np.ones(ncpl, dtype=np.int32)
    rech = np.zeros(ncpl)
    periods = {}
    for nm, num, lines in blocks:
        if nm != "PERIOD":
            continue
        g = read_griddata(lines, base_dir, {"IRCH": (a2, np.int32), "RECHARGE": (a2, np.float64)})
        irch = g.get("IRCH", irch)
        rech = g.get("RECHARGE", rech)
        if irch.min() < 1 or irch.max() > shape[0]:
            raise Mf6InputError(f"RCHA {name}: IRCH outside 1..{shape[0]}")
        nodes = (irch.astype(np.int64) - 1) * ncpl + np.arange(ncpl)
        periods[num] = Package(T.PKG_RCH, nodes, rech.copy(), iflowred=fixed_cell)
    return StressPackage("RCH", name, periods, fixed_cell, from_arrays=True)


def read_stress_package(path, ftype, name, shape, inewton=0):
    b = read_blocks(path)
    opt = _options(_block(b, "OPTIONS", required=False))
    auxnames = [a.upper() for a in opt.get("AUXILIARY", opt.get("AUX", []))]
    naux = len(auxnames)
    for k in opt:
        if k in ("TAS6", "MOVER") or (k == "READASARRAYS" and ftype != "RCH6"):
            raise Mf6InputError(f"{path}: option {k} is not supported on the GPU path")
    fixed_cell = 1 if (ftype == "RCH6" and "FIXED_CELL" in opt) else 0    # carried in Package.iflowred for RCH
    if "READASARRAYS" in opt:
        return _read_rcha(b, name, shape, fixed_cell, os.path.dirname(path)), naux
    iflowred, flowred = fixed_cell, 0.1
    if "AUTO_FLOW_REDUCE" in opt:
        iflowred, flowred = 1, float(opt["AUTO_FLOW_REDUCE"][0]) if opt["AUTO_FLOW_REDUCE"] else 0.1
    ncol = _PKG_NCOL[ftype]
    # DRN: AUXDEPTHNAME names the auxiliary column that holds the drainage depth (carried in b3); the discharge
    # scaling is cubic under NEWTON or with DEV_CUBIC_SCALING (carried in Package.iflowred), gwf-drn.f90:127-135, 203-231
    depth_col = None
    if ftype == "DRN6":
        iflowred = 1 if (inewton or "DEV_CUBIC_SCALING" in opt) else 0
        if "AUXDEPTHNAME" in opt:
            nm_d = opt["AUXDEPTHNAME"][0].upper()
            if nm_d not in auxnames:
                raise Mf6InputError(f"{path}: AUXDEPTHNAME {nm_d} is not one of the AUXILIARY variables")
            depth_col = ncol + auxnames.index(nm_d)
    nread = ncol + naux
    # TS6 FILEIN <file> (any number of them): entries of the list may name a time series instead of a number;
    # AUXMULTNAME: the auxiliary variable that multiplies the rate / conductance / head column
    series = {}
    for ln in _block(b, "OPTIONS", required=False):
        if ln[0].upper() == "TS6":
            series.update(read_ts_file(os.path.join(os.path.dirname(path), ln[-1]), read_blocks))
    mult = None
    if "AUXMULTNAME" in opt:
        nm_m = opt["AUXMULTNAME"][0].upper()
        if nm_m not in auxnames:
            raise Mf6InputError(f"{path}: AUXMULTNAME {nm_m} is not one of the AUXILIARY variables")
        mult = ({"CHD6": "b1", "WEL6": "b1", "RCH6": "b1", "RIV6": "b2", "GHB6": "b2", "DRN6": "b2"}[ftype],
                auxnames.index(nm_m))

    def numbers(tokens, row, links):
        out = []
        for c, tok in enumerate(tokens):
            try:
                out.append(float(tok))
            except ValueError:
                if tok.upper() not in series:
                    raise Mf6InputError(f"{path}: '{tok}' is neither a number nor a time series of this package") from None
                if depth_col is not None and c == depth_col:
                    target = "b3"
                else:
                    target = ("b1", "b2", "b3")[c] if c < ncol else ("aux", c - ncol)
                links.append((target, row, tok.upper()))
                out.append(0.0)
        return out

    periods = {}
    for nm, num, lines in b:
        if nm != "PERIOD":
            continue
        nodes, vals, links = [], [], []
        for t in lines:
            if t[0].upper() == "OPEN/CLOSE":
                # the list in an external file (ListReader.f90): text rows like the inline ones, or (BINARY)
                # records of cellid as i32 followed by the bound and auxiliary columns as f64
                fn = os.path.join(os.path.dirname(path), t[1])
                if "(BINARY)" in [x.upper() for x in t[2:]]:
                    nd = len(shape)
                    dt = np.dtype([("cellid", "<i4", (nd,)), ("v", "<f8", (ncol + naux,))])
                    rec = np.fromfile(fn, dtype=dt)
                    for r in rec:
                        node, _ = _cellid([str(int(c)) for c in r["cellid"]], shape)
                        nodes.append(node)
                        vals.append([float(x) for x in r["v"][:nread]])
                else:
                    with open(fn) as f:
                        for raw in f:
                            tt = _tokens(raw)
                            if tt:
                                node, w = _cellid(tt, shape)
                                nodes.append(node)
                                vals.append(numbers(tt[w:w + nread], len(nodes) - 1, links))
                continue
            node, w = _cellid(t, shape)
            nodes.append(node)
            vals.append(numbers(t[w:w + nread], len(nodes) - 1, links))
        if nodes:
            v = np.array(vals)
            cols = [v[:, c] if c < ncol else None for c in range(3)]
            if depth_col is not None:
                cols[2] = v[:, depth_col]
            periods[num] = Package(_PKG_TYPE[ftype], np.array(nodes), cols[0], cols[1], cols[2], iflowred=iflowred,
                                   flowred=flowred, auxnames=tuple(auxnames),
                                   aux=v[:, ncol:ncol + naux].copy() if naux else None,
                                   ts_links=links or None, series=series or None, mult=mult)
        else:
            periods[num] = None
    return StressPackage(ftype[:-1], name, periods, iflowred, flowred), naux


def read_hfb(path, shape, nodereduced, model):
    """HFB6 (gwf-hfb.dfn; hfb_rp / read_data / check_data, gwf-hfb.f90:149-201, 596-760): PERIOD blocks of
    `cellid1 cellid2 hydchr`; a block replaces the whole barrier list, which then stays in force.  The two cells of
    a barrier must be horizontally connected."""
    b = read_blocks(path)
    dim = _options(_block(b, "DIMENSIONS"))
    maxhfb = int(dim["MAXHFB"][0])
    row_of = np.repeat(np.arange(model.nodes), np.diff(model.ia))
    conn = {(int(r), int(c)): int(j) for r, c, j in zip(row_of, model.ja, model.jas) if j >= 0}
    periods = {}
    for nm, num, lines in b:
        if nm != "PERIOD":
            continue
        n1, n2, hc = [], [], []
        for t in lines:
            a, w = _cellid(t, shape)
            c, w2 = _cellid(t[w:], shape)
            if nodereduced is not None:
                a, c = int(nodereduced[a]), int(nodereduced[c])
                if a < 0 or c < 0:
                    raise Mf6InputError(f"{path}: period {num}: a barrier lies in a cell that IDOMAIN removes")
            j = conn.get((a, c))
            if j is None:
                raise Mf6InputError(f"{path}: period {num}: HFB cells {t[:w]} and {t[w:w + w2]} are not connected")
            if model.ihc[j] == 0:
                raise Mf6InputError(f"{path}: period {num}: HFB between vertically connected cells "
                                    f"{t[:w]} and {t[w:w + w2]}")
            n1.append(a)
            n2.append(c)
            hc.append(float(t[w + w2]))
        if len(n1) > maxhfb:
            raise Mf6InputError(f"{path}: period {num}: {len(n1)} barriers, MAXHFB is {maxhfb}")
        periods[num] = (np.array(n1, dtype=np.int32), np.array(n2, dtype=np.int32), np.array(hc, dtype=np.float64))
    return periods


def read_gwf_model(name, nam_path, base_dir, warnings):
    b = read_blocks(nam_path)
    opt = _options(_block(b, "OPTIONS", required=False))
    inewton = 1 if "NEWTON" in opt else 0
    inewtonur = 1 if inewton and opt["NEWTON"] and opt["NEWTON"][0].upper() == "UNDER_RELAXATION" else 0
    list_file = os.path.join(base_dir, opt["LIST"][0]) if opt.get("LIST") else os.path.splitext(nam_path)[0] + ".lst"
    files = {}
    stress = []
    for t in _block(b, "PACKAGES"):
        ft = t[0].upper()
        fn = os.path.join(base_dir, t[1])
        pn = t[2] if len(t) > 2 else None
        if ft in _PKG_TYPE:
            stress.append((ft, fn, pn))
        elif ft in ("DIS6", "DISV6", "DISU6", "IC6", "NPF6", "STO6", "OC6", "HFB6", "GNC6"):
            files[ft] = fn
        else:
            raise Mf6InputError(f"{nam_path}: package {ft} is outside the GPU path (SURVEY.md section 8)")
    for need in ("IC6", "NPF6"):
        if need not in files:
            raise Mf6InputError(f"{nam_path}: {need} package is required")
    if sum(k in files for k in ("DIS6", "DISV6", "DISU6")) != 1:
        raise Mf6InputError(f"{nam_path}: exactly one of DIS6 / DISV6 / DISU6 is required")
    cell2d = None
    disu = None
    if "DISU6" in files:
        d = read_blocks(files["DISU6"])
        dim = _options(_block(d, "DIMENSIONS"))
        nodes, nja = int(dim["NODES"][0]), int(dim["NJA"][0])
        nlay, shape = 1, (nodes,)
        g = read_griddata(_block(d, "GRIDDATA"), base_dir,
                          {"TOP": ((nodes,), np.float64), "BOT": ((nodes,), np.float64), "AREA": ((nodes,), np.float64),
                           "IDOMAIN": ((nodes,), np.int32)})
        disu = read_griddata(_block(d, "CONNECTIONDATA"), base_dir,
                             {"IAC": ((nodes,), np.int32), "JA": ((nja,), np.int32), "IHC": ((nja,), np.int32),
                              "CL12": ((nja,), np.float64), "HWVA": ((nja,), np.float64),
                              "ANGLDEGX": ((nja,), np.float64)})
        if int(disu["IAC"].sum()) != nja:
            raise Mf6InputError(f"{files['DISU6']}: sum(IAC) != NJA")
    elif "DIS6" in files:
        d = read_blocks(files["DIS6"])
        dim = _options(_block(d, "DIMENSIONS"))
        nlay, nrow, ncol = int(dim["NLAY"][0]), int(dim["NROW"][0]), int(dim["NCOL"][0])
        shape = (nlay, nrow, ncol)
        g = read_griddata(_block(d, "GRIDDATA"), base_dir,
                          {"DELR": ((ncol,), np.float64), "DELC": ((nrow,), np.float64),
                           "TOP": ((nrow, ncol), np.float64), "BOTM": (shape, np.float64),
                           "IDOMAIN": (shape, np.int32)})
    else:
        d = read_blocks(files["DISV6"])
        dim = _options(_block(d, "DIMENSIONS"))
        nlay, ncpl, nvert = int(dim["NLAY"][0]), int(dim["NCPL"][0]), int(dim["NVERT"][0])
        shape = (nlay, ncpl)
        g = read_griddata(_block(d, "GRIDDATA"), base_dir,
                          {"TOP": ((1, 1, ncpl), np.float64), "BOTM": ((nlay, 1, ncpl), np.float64),
                           "IDOMAIN": ((nlay, 1, ncpl), np.int32)})
        verts = np.zeros((nvert, 2))
        for t in _block(d, "VERTICES"):
            verts[int(t[0]) - 1] = float(t[1]), float(t[2])
        cells = [None] * ncpl
        for t in _block(d, "CELL2D"):
            nv = int(t[3])
            cells[int(t[0]) - 1] = (float(t[1]), float(t[2]), [int(v) - 1 for v in t[4:4 + nv]])
        if any(c is None for c in cells):
            raise Mf6InputError(f"{files['DISV6']}: CELL2D does not list every cell")
        cell2d = cell2d_from_vertices(verts, cells)
    # what the binary grid file (.grb) records about the grid, in USER numbering
    dopt = _options(_block(d, "OPTIONS", required=False))
    grid = dict(xorigin=float(dopt["XORIGIN"][0]) if "XORIGIN" in dopt else 0.0,
                yorigin=float(dopt["YORIGIN"][0]) if "YORIGIN" in dopt else 0.0,
                angrot=float(dopt["ANGROT"][0]) if "ANGROT" in dopt else 0.0,
                nogrb="NOGRB" in dopt,
                file=os.path.join(base_dir, dopt["GRB6"][-1]) if "GRB6" in dopt else None)
    if disu is not None:
        grid.update(kind="DISU", nodes=nodes, top=g["TOP"].reshape(-1), bot=g["BOT"].reshape(-1),
                    default=files["DISU6"] + ".grb")
    elif cell2d is None:
        grid.update(kind="DIS", nlay=nlay, nrow=nrow, ncol=ncol, delr=g["DELR"].reshape(-1), delc=g["DELC"].reshape(-1),
                    top=g["TOP"].reshape(-1), botm=g["BOTM"].reshape(-1), default=files["DIS6"] + ".grb")
    else:
        grid.update(kind="DISV", nlay=nlay, ncpl=ncpl, vertices=verts, cells=cells, top=g["TOP"].reshape(-1),
                    botm=g["BOTM"].reshape(-1), default=files["DISV6"] + ".grb")
    grid["idomain"] = (g["IDOMAIN"].reshape(-1).astype(np.int32) if "IDOMAIN" in g
                       else np.ones(int(np.prod(shape)), dtype=np.int32))
    # READARRAY layout (LAYERED = one control record per layer)
    ashape = shape if len(shape) == 3 else ((nlay, 1, shape[1]) if len(shape) == 2 else shape)
    # IC / NPF / STO
    ic = read_griddata(_block(read_blocks(files["IC6"]), "GRIDDATA"), base_dir, {"STRT": (ashape, np.float64)})
    nb = read_blocks(files["NPF6"])
    nopt = _options(_block(nb, "OPTIONS", required=False))
    for k in nopt:
        if k in ("XT3D", "TVK6"):
            raise Mf6InputError(f"NPF option {k} is not supported on the GPU path")
        if k in ("SAVE_SPECIFIC_DISCHARGE", "SAVE_SATURATION"):
            warnings.append(f"{name}: NPF {k}: the DATA-{'SPDIS' if 'SPEC' in k else 'SAT'} record is not written to the "
                            "budget file by this path")
    np_ = read_griddata(_block(nb, "GRIDDATA"), base_dir,
                        {"ICELLTYPE": (ashape, np.int32), "K": (ashape, np.float64), "K22": (ashape, np.float64),
                         "K33": (ashape, np.float64), "ANGLE1": (ashape, np.float64), "ANGLE2": (ashape, np.float64),
                         "ANGLE3": (ashape, np.float64), "WETDRY": (ashape, np.float64)})
    # K22OVERK / K33OVERK: the arrays hold ratios (gwf-npf.f90 prepcheck: k22 = k22 * k11, k33 = k33 * k11)
    if "K22OVERK" in nopt and "K22" in np_:
        np_["K22"] = np_["K22"] * np_["K"]
    if "K33OVERK" in nopt and "K33" in np_:
        np_["K33"] = np_["K33"] * np_["K"]
    kw = {}
    if "REWET" in nopt:
        # REWET WETFCT <wetfct> IWETIT <iwetit> IHDWET <ihdwet> (gwf-npf.dfn); needs the WETDRY array (:1677-1684)
        r = [t.upper() for t in nopt["REWET"]]
        if "WETDRY" not in np_:
            raise Mf6InputError("NPF: REWET needs the WETDRY array")
        if inewton:
            raise Mf6InputError("NPF: REWET cannot be used with NEWTON")
        if "IDOMAIN" in g and (g["IDOMAIN"] <= 0).any():
            raise Mf6InputError("NPF: REWET on a grid with removed cells (IDOMAIN <= 0) is not supported on the GPU path")
        kw.update(irewet=1, wetdry=np_["WETDRY"].reshape(-1),
                  wetfct=float(r[r.index("WETFCT") + 1]) if "WETFCT" in r else 1.0,
                  iwetit=int(r[r.index("IWETIT") + 1]) if "IWETIT" in r else 1,
                  ihdwet=int(r[r.index("IHDWET") + 1]) if "IHDWET" in r else 0)
    aniso = ("K22" in np_ and not np.array_equal(np_["K22"], np_["K"])) or any(a in np_ for a in ("ANGLE1", "ANGLE2", "ANGLE3"))
    if aniso:
        # hy_eff (gwf-npf.f90:2280-2355); angles are given in degrees and stored in radians (:1213-1239)
        if disu is not None or cell2d is not None or ("IDOMAIN" in g and (g["IDOMAIN"] <= 0).any()):
            raise Mf6InputError("NPF K22 / ANGLE anisotropy is supported on full DIS grids only on the GPU path")
        if "ANGLE2" in np_ and "ANGLE1" not in np_:
            raise Mf6InputError("NPF: ANGLE2 needs ANGLE1 (gwf-npf.f90 prepcheck)")
        if "ANGLE3" in np_ and "ANGLE2" not in np_:
            raise Mf6InputError("NPF: ANGLE3 needs ANGLE1 and ANGLE2 (gwf-npf.f90 prepcheck)")
        kw["k22"] = (np_["K22"] if "K22" in np_ else np_["K"]).reshape(shape)
        for a in ("ANGLE1", "ANGLE2", "ANGLE3"):
            if a in np_:
                kw[a.lower()] = (np_[a] * (np.arctan(1.0) / 45.0)).reshape(shape)   # DPIO180, Constants.f90:130
    if "THICKSTRT" in nopt:
        kw["ithickstrt"] = 1
    avg = {"LOGARITHMIC": 1, "AMT-LMK": 2, "AMT-HMK": 3}
    if "ALTERNATIVE_CELL_AVERAGING" in nopt:
        kw["icellavg"] = avg[nopt["ALTERNATIVE_CELL_AVERAGING"][0].upper()]
    if "PERCHED" in nopt:
        kw["iperched"] = 1
    if "VARIABLECV" in nopt:
        kw["ivarcv"] = 1
        if nopt["VARIABLECV"] and nopt["VARIABLECV"][0].upper() == "DEWATERED":
            kw["idewatcv"] = 1
    sto = {}
    sto_tr = {}
    if "STO6" in files:
        sb = read_blocks(files["STO6"])
        sopt = _options(_block(sb, "OPTIONS", required=False))
        sto = read_griddata(_block(sb, "GRIDDATA"), base_dir,
                            {"ICONVERT": (ashape, np.int32), "SS": (ashape, np.float64), "SY": (ashape, np.float64)})
        if "STORAGECOEFFICIENT" in sopt:
            kw["istor_coef"] = 1
        if "SS_CONFINED_ONLY" in sopt:
            kw["iconf_ss"] = 1
        for nm, num, lines in sb:
            if nm == "PERIOD":
                key = lines[0][0].upper() if lines else "STEADY-STATE"
                sto_tr[num] = (key == "TRANSIENT")
    common = dict(k33=np_["K33"] if "K33" in np_ else None, icelltype=np_["ICELLTYPE"], strt=ic["STRT"],
                  ss=sto.get("SS"), sy=sto.get("SY"), iconvert=sto.get("ICONVERT"), inewton=inewton,
                  inewtonur=inewtonur, **kw)
    if disu is not None:
        m = build_disu_model(disu["IAC"], np.abs(disu["JA"]) - 1, disu["IHC"], disu["CL12"], disu["HWVA"], g["TOP"],
                             g["BOT"], g["AREA"], np_["K"], **common)
    elif cell2d is None:
        r3 = lambda a: None if a is None else a.reshape(shape)   # noqa: E731
        common = {k: (r3(v) if isinstance(v, np.ndarray) else v) for k, v in common.items()}
        if "IDOMAIN" in g and (g["IDOMAIN"] <= 0).any():
            # reduced node numbering + vertical pass-through cells, like the reference numbers the grid
            m = build_dis_model_idomain(nlay, nrow, ncol, g["DELR"], g["DELC"], g["TOP"].reshape(nrow, ncol),
                                        g["BOTM"].reshape(shape), g.pop("IDOMAIN").reshape(shape),
                                        np_["K"].reshape(shape), **common)
        else:
            m = build_dis_model(nlay, nrow, ncol, g["DELR"], g["DELC"], g["TOP"].reshape(nrow, ncol),
                                g["BOTM"].reshape(shape), np_["K"].reshape(shape), **common)
    else:
        m = build_disv_model(nlay, cell2d, g["TOP"], g["BOTM"].reshape(nlay, shape[1]), np_["K"], **common)
    if "IDOMAIN" in g:       # DISV / DISU (the DIS branch consumed its IDOMAIN above)
        if (g["IDOMAIN"] < 0).any():
            raise Mf6InputError("IDOMAIN < 0 (vertical pass-through cells) is supported on DIS grids only on the GPU path")
        if (g["IDOMAIN"] == 0).any():
            # reduced node numbering, like the reference: the removed cells and their connections do not exist
            m = reduce_model(m, g["IDOMAIN"].reshape(-1) > 0)
    grid["icelltype"] = np_["ICELLTYPE"].reshape(-1).astype(np.int32)      # as read, user numbering (dis_ar)
    gi = GwfInput(name=name, model=m, shape=shape, grid=grid, sto_transient=sto_tr, nodeuser=m.meta.get("nodeuser"),
                  nodereduced=m.meta.get("nodereduced"))
    count = {}
    for ft, fn, pn in stress:
        count[ft] = count.get(ft, 0) + 1
        sp, naux = read_stress_package(fn, ft, pn or f"{ft[:-1]}-{count[ft]}", shape, inewton)
        if gi.nodereduced is not None:          # user cellids -> reduced nodes; boundaries in removed cells are dropped
            for iper, p in sp.periods.items():
                if p is None:
                    continue
                red = gi.nodereduced[p.nodelist]
                keep = red >= 0
                if not keep.all() and not sp.from_arrays:
                    # the reference stops with an error for a LIST boundary in a cell that IDOMAIN removes
                    # (DiscretizationBase noder / "cell is outside active grid domain")
                    raise Mf6InputError(f"{name}: {sp.name} period {iper}: {int((~keep).sum())} boundaries lie in "
                                        "cells that IDOMAIN removes")
                # array input (nlarray_to_nodelist): such an entry gets node 0 -- no recharge, no hand-down to the
                # layer below, no budget record; the others keep their position in the list as the bound number
                # (autotest/test_gwf_rch02.py, test_gwf_rch03.py)
                sp.periods[iper] = p.with_nodes(red, keep=None if keep.all() else keep)
        gi.packages.append(sp)
    gi.list_file = list_file
    if "HFB6" in files:
        gi.hfb = read_hfb(files["HFB6"], shape, gi.nodereduced, m)
    if "GNC6" in files:
        gi.gnc = _reduce_gnc(read_gnc(files["GNC6"], shape, shape, warnings), gi.nodereduced, gi.nodereduced, name)
    if "OC6" in files:
        ob = read_blocks(files["OC6"])
        oopt = _block(ob, "OPTIONS", required=False)
        for t in oopt:
            if t[0].upper() == "HEAD" and t[1].upper() == "FILEOUT":
                gi.head_file = os.path.join(base_dir, t[2])
            elif t[0].upper() == "BUDGET" and t[1].upper() == "FILEOUT":
                gi.budget_file = os.path.join(base_dir, t[2])
            elif t[0].upper() == "BUDGETCSV" and t[1].upper() == "FILEOUT":
                gi.budgetcsv_file = os.path.join(base_dir, t[2])
        for nm, num, lines in ob:
            if nm == "PERIOD":
                gi.save[num] = [(t[1].upper(), [x.upper() for x in t[2:]]) for t in lines if t[0].upper() == "SAVE"]
                gi.printrec[num] = [(t[1].upper(), [x.upper() for x in t[2:]]) for t in lines
                                    if t[0].upper() == "PRINT"]
    return gi


def read_gnc(path, shape_n, shape_m, warnings):
    """GNC6 (gwf-gnc.dfn; GhostNode.f90 read_options / read_dimensions / read_data :637-864): GNCDATA rows of
    cellidn cellidm cellidsj(numalphaj) alphasj(numalphaj); cellidn and the contributing cells belong to the first
    grid, cellidm to the second (the same grid for the single-model package); an all-zero cellid = no cell.
    Returns user node numbers (0-based): (noden, nodem, nodesj[ngnc, numj] with -1 = none, alphasj)."""
    b = read_blocks(path)
    opt = _options(_block(b, "OPTIONS", required=False))
    if "EXPLICIT" not in opt:
        warnings.append(f"{os.path.basename(path)}: the ghost node correction is applied EXPLICITLY on the GPU path "
                        "(right-hand side; the same converged heads as the implicit variant, more outer iterations)")
    dim = _options(_block(b, "DIMENSIONS"))
    ngnc, numj = int(dim["NUMGNC"][0]), int(dim["NUMALPHAJ"][0])
    wn = len(shape_n) if len(shape_n) > 1 else 1
    noden, nodem, nodesj, alphasj = [], [], [], []
    for t in _block(b, "GNCDATA"):
        a, w = _cellid(t, shape_n)
        c, w2 = _cellid(t[w:], shape_m)
        pos = w + w2
        js = []
        for _ in range(numj):
            tok = t[pos:pos + wn]
            js.append(-1 if all(int(v) == 0 for v in tok) else _cellid(tok, shape_n)[0])
            pos += wn
        noden.append(a)
        nodem.append(c)
        nodesj.append(js)
        alphasj.append([float(v) for v in t[pos:pos + numj]])
    if len(noden) != ngnc:
        raise Mf6InputError(f"{path}: NUMGNC is {ngnc}, GNCDATA holds {len(noden)} entries")
    return (np.array(noden, dtype=np.int64), np.array(nodem, dtype=np.int64),
            np.array(nodesj, dtype=np.int64).reshape(ngnc, numj), np.array(alphasj).reshape(ngnc, numj))


def _reduce_gnc(gnc, red_n, red_m, what):
    """user -> reduced node numbers of a grid with removed cells"""
    noden, nodem, nodesj, alphasj = gnc
    if red_n is not None:
        noden = red_n[noden]
        nodesj = np.where(nodesj >= 0, red_n[np.maximum(nodesj, 0)], -1)
    if red_m is not None:
        nodem = red_m[nodem]
    if (noden < 0).any() or (nodem < 0).any():
        raise Mf6InputError(f"{what}: a ghost node correction names a cell that IDOMAIN removes")
    return noden, nodem, nodesj, alphasj


def read_exchange(path, m1, m2, shape1, shape2, exg_id=1):
    b = read_blocks(path)
    opt = _options(_block(b, "OPTIONS", required=False))
    for k in opt:
        if k in ("MVR6", "XT3D", "CELL_AVERAGING", "VARIABLECV", "DEWATERED"):
            raise Mf6InputError(f"{path}: exchange option {k} is not supported on the GPU path")
    auxname = [a.upper() for a in opt.get("AUXILIARY", opt.get("AUX", []))]
    n1, n2, ihc, cl1, cl2, hw, aux = [], [], [], [], [], [], []
    for t in _block(b, "EXCHANGEDATA"):
        a, w = _cellid(t, shape1)
        c, w2 = _cellid(t[w:], shape2)
        r = t[w + w2:]
        n1.append(a); n2.append(c)
        ihc.append(int(r[0])); cl1.append(float(r[1])); cl2.append(float(r[2])); hw.append(float(r[3]))
        aux.append([float(v) for v in r[4:4 + len(auxname)]])
    # name as simulation_cr builds it (SimulationCreate.f90:455); SAVE_FLOWS = ipakcb (DisConnExchange.f90:142-146)
    return dict(m1=m1, m2=m2, nodem1=np.array(n1), nodem2=np.array(n2), ihc=np.array(ihc, dtype=np.int32),
                cl1=np.array(cl1), cl2=np.array(cl2), hwva=np.array(hw), name=f"GWF-GWF_{exg_id}",
                save_flows="SAVE_FLOWS" in opt, auxname=auxname,
                aux=np.array(aux, dtype=np.float64).reshape(len(n1), len(auxname)),
                gnc_file=os.path.join(os.path.dirname(path), opt["GNC6"][-1]) if "GNC6" in opt else None)


def read_simulation(sim_dir):
    """mfsim.nam and everything it names (src/SimulationCreate.f90 order: timing, models, exchanges, solutions)"""
    sim_dir = os.path.abspath(sim_dir)
    b = read_blocks(os.path.join(sim_dir, "mfsim.nam"))
    warnings = []
    tim = _block(b, "TIMING")
    nper, pd = read_tdis(os.path.join(sim_dir, tim[0][1]))
    models, index = [], {}
    for t in _block(b, "MODELS"):
        if t[0].upper() != "GWF6":
            raise Mf6InputError(f"model type {t[0]} is outside the GPU path (GWF6 only)")
        index[t[2].upper()] = len(models)
        models.append(read_gwf_model(t[2], os.path.join(sim_dir, t[1]), sim_dir, warnings))
    exchanges = []
    for t in _block(b, "EXCHANGES", required=False):
        if t[0].upper() != "GWF6-GWF6":
            raise Mf6InputError(f"exchange type {t[0]} is outside the GPU path")
        i1, i2 = index[t[2].upper()], index[t[3].upper()]
        e = read_exchange(os.path.join(sim_dir, t[1]), i1, i2, models[i1].shape, models[i2].shape,
                          exg_id=len(exchanges) + 1)
        e["usernodem1"], e["usernodem2"] = e["nodem1"], e["nodem2"]
        e["gnc"] = None
        if e["gnc_file"]:
            g = read_gnc(e["gnc_file"], models[i1].shape, models[i2].shape, warnings)
            e["gnc"] = _reduce_gnc(g, models[i1].nodereduced, models[i2].nodereduced, t[1])
        for key, gi in (("nodem1", models[i1]), ("nodem2", models[i2])):
            if gi.nodereduced is not None:
                e[key] = gi.nodereduced[e[key]]
                if (e[key] < 0).any():
                    raise Mf6InputError(f"{t[1]}: an exchange connects a cell that IDOMAIN removes from {gi.name}")
        exchanges.append(e)
    sg = [x for x in b if x[0] == "SOLUTIONGROUP"]
    if len(sg) != 1 or len(sg[0][2]) != 1 or sg[0][2][0][0].upper() != "IMS6":
        raise Mf6InputError("exactly one solution group with one IMS6 solution is supported")
    sl = sg[0][2][0]
    if sorted(x.upper() for x in sl[2:]) != sorted(index):
        raise Mf6InputError("every model must belong to the one IMS solution")
    sln, ims = read_ims(os.path.join(sim_dir, sl[1]), warnings)
    return Simulation(sim_dir, nper, pd, models, exchanges, sln, ims, warnings,
                      time_units=read_time_units(os.path.join(sim_dir, tim[0][1])),
                      continue_="CONTINUE" in _options(_block(b, "OPTIONS", required=False)))
