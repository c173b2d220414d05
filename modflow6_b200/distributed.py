"""Split-model (multi-GPU) host side: one GWF submodel per rank / GPU.

Mirrors what the reference does with FloPy's Mf6Splitter + GWF-GWF exchanges +
`mf6 -p` (src/Exchange/exg-gwfgwf.f90:363-661, src/Model/Connection/*,
src/Distributed/*): a DIS grid is cut into pr x pc blocks of (row, col); every
rank owns one block (all layers), sees the face-adjacent cells of its
neighbours as HALO cells, and the cross-block connections become ordinary
connections of the rank's extended model (cl1/cl2/hwva/ihc = the EXCHANGEDATA
columns).  Heads/ibound of halo cells move by NCCL send/recv, Krylov scalars by
all-gathers, the preconditioner is ILU0 of the rank's diagonal block.

Every rank builds only its own block (+ a one-cell ring), so the 128M-cell C5
configuration never exists in one piece on the host.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import ctypes_types as T
from .grid import GwfModel, Package, build_dis_model, package_array
from .lib import check, ensure_init, load


# ----------------------------------------------------------------------------------------------
# decomposition-independent seeded fields (same value for a global cell id on every rank)
def _mix(a):
    a = (a ^ (a >> np.uint64(33))) * np.uint64(0xFF51AFD7ED558CCD)
    a = (a ^ (a >> np.uint64(33))) * np.uint64(0xC4CEB9FE1A85EC53)
    return a ^ (a >> np.uint64(33))


def hash_uniform(gid, seed):
    """U(0,1) per global cell id (splitmix-style integer hash): reproducible on any sub-box"""
    with np.errstate(over="ignore"):
        h = _mix(gid.astype(np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) + 1.1102230246251565e-16


def hash_normal(gid, seed):
    u1, u2 = hash_uniform(gid, seed), hash_uniform(gid, seed + 7919)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


@dataclass
class GridSpec:
    """Global DIS grid + property recipe (functions of the global cell id)."""
    nlay: int
    nrow: int
    ncol: int
    delr: float = 100.0
    delc: float = 100.0
    top: float = 0.0
    dz: float = 10.0
    kmean: float = 10.0
    ksigma: float = 1.0
    k33_ratio: float = 0.1
    seed: int = 20260101
    strt: float = 44.0
    icelltype: int = 0

    def gid(self, k, i, j):
        return (k.astype(np.int64) * self.nrow + i) * self.ncol + j

    def box_model(self, i0, i1, j0, j1, **opts):
        """DIS model of rows [i0,i1) x cols [j0,j1), all layers, with the global recipe."""
        nl, nr, nc = self.nlay, i1 - i0, j1 - j0
        kk, ii, jj = np.meshgrid(np.arange(nl), np.arange(i0, i1), np.arange(j0, j1), indexing="ij")
        gid = self.gid(kk, ii, jj)
        k = np.exp(np.log(self.kmean) + self.ksigma * hash_normal(gid, self.seed))
        botm = self.top - self.dz * np.arange(1, nl + 1)
        m = build_dis_model(nl, nr, nc, self.delr, self.delc, self.top, botm, k, k33=self.k33_ratio * k,
                            icelltype=self.icelltype, strt=self.strt, **opts)
        return m, gid.reshape(-1)


def block_ranges(n, parts):
    edges = [(n * p) // parts for p in range(parts + 1)]
    return [(edges[p], edges[p + 1]) for p in range(parts)]


@dataclass
class SubModel:
    rank: int
    nranks: int
    model: GwfModel           # extended: owned cells first, then halo cells
    n_own: int
    global_id: np.ndarray     # [n_ext]
    nbr_rank: np.ndarray
    send_ptr: np.ndarray
    send_idx: np.ndarray
    recv_ptr: np.ndarray
    block: tuple              # (i0, i1, j0, j1)
    spec: GridSpec = None
    meta: dict = field(default_factory=dict)

    def local_nodes(self, gnodes):
        """global node ids -> (mask of the ones owned here, their local indices)"""
        s = self.spec
        i0, i1, j0, j1 = self.block
        g = np.asarray(gnodes, dtype=np.int64)
        k = g // (s.nrow * s.ncol)
        rem = g - k * (s.nrow * s.ncol)
        i = rem // s.ncol
        j = rem - i * s.ncol
        mask = (i >= i0) & (i < i1) & (j >= j0) & (j < j1)
        loc = (k * (i1 - i0) + (i - i0)) * (j1 - j0) + (j - j0)
        return mask, loc[mask].astype(np.int32)

    def localize_packages(self, pkgs):
        """same list of package types on every rank, each restricted to the owned cells"""
        out = []
        for p in pkgs:
            mask, loc = self.local_nodes(p.nodelist)
            out.append(Package(p.type, loc, p.b1[mask], p.b2[mask], p.b3[mask], p.iflowred, p.flowred))
        return out


def build_dis_block(spec, pr, pc, rank, **opts):
    """Submodel of `rank` in a pr x pc block decomposition of (row, col)."""
    nranks = pr * pc
    rb, cb = block_ranges(spec.nrow, pr), block_ranges(spec.ncol, pc)
    bi, bj = divmod(rank, pc)
    i0, i1 = rb[bi]
    j0, j1 = cb[bj]
    ie0, ie1 = max(i0 - 1, 0), min(i1 + 1, spec.nrow)
    je0, je1 = max(j0 - 1, 0), min(j1 + 1, spec.ncol)
    box, gid_box = spec.box_model(ie0, ie1, je0, je1, **opts)
    nl, nre, nce = box.shape
    nb = box.nodes
    b = np.arange(nb, dtype=np.int64)
    kb = b // (nre * nce)
    rem = b - kb * (nre * nce)
    ib = rem // nce + ie0
    jb = rem - (rem // nce) * nce + je0
    in_i, in_j = (ib >= i0) & (ib < i1), (jb >= j0) & (jb < j1)
    owned = in_i & in_j
    halo = (in_i ^ in_j) & (in_i | in_j)        # ring cells sharing a face with the block (no corners)
    # owner rank of every halo cell
    own_bi = np.searchsorted(np.array([r[1] for r in rb]), ib, side="right")
    own_bj = np.searchsorted(np.array([c[1] for c in cb]), jb, side="right")
    owner = own_bi * pc + own_bj
    oidx = np.nonzero(owned)[0]                      # ascending box order == ascending global id
    hidx = np.nonzero(halo)[0]
    horder = np.lexsort((gid_box[hidx], owner[hidx]))
    hidx = hidx[horder]
    n_own, n_halo = oidx.size, hidx.size
    n_ext = n_own + n_halo
    box2loc = np.full(nb, -1, dtype=np.int64)
    box2loc[oidx] = np.arange(n_own)
    box2loc[hidx] = n_own + np.arange(n_halo)
    # CSR of the owned rows
    ia_b, ja_b, jas_b = box.ia.astype(np.int64), box.ja, box.jas
    cnt = (ia_b[oidx + 1] - ia_b[oidx])
    ia = np.zeros(n_ext + 1, dtype=np.int64)
    np.cumsum(cnt, out=ia[1:n_own + 1])
    nja_own = int(ia[n_own])
    # positions of the owned rows' entries in the box arrays
    starts = np.repeat(ia_b[oidx], cnt)
    within = np.arange(nja_own) - np.repeat(ia[:n_own], cnt)
    pos = starts + within
    ja = box2loc[ja_b[pos]]
    assert (ja >= 0).all(), "an owned cell touches a corner cell"
    jas_old = jas_b[pos].astype(np.int64)
    diag = within == 0
    uniq, inv = np.unique(jas_old[~diag], return_inverse=True)
    jas = np.full(nja_own, -1, dtype=np.int64)
    jas[~diag] = inv
    # halo rows: diagonal only
    ia[n_own + 1:] = nja_own + np.arange(1, n_halo + 1)
    ja_full = np.concatenate([ja, n_own + np.arange(n_halo)])
    jas_full = np.concatenate([jas, np.full(n_halo, -1, dtype=np.int64)])
    # isym (position of the transposed entry; only rows of owned cells have one)
    rows = np.concatenate([np.repeat(np.arange(n_own), cnt), n_own + np.arange(n_halo)])
    key = rows.astype(np.int64) * n_ext + ja_full
    order = np.argsort(key, kind="stable")
    tkey = ja_full.astype(np.int64) * n_ext + rows
    loc = np.searchsorted(key[order], tkey)
    loc = np.minimum(loc, key.size - 1)
    isym = np.where(key[order][loc] == tkey, order[loc], 0)
    sel = np.concatenate([oidx, hidx])
    ibot = box2loc[box.ibotnode[sel]]
    ibot = np.where(ibot < 0, np.arange(n_ext), ibot)    # halo columns: not used
    m = GwfModel(nodes=n_ext, ia=ia, ja=ja_full, jas=jas_full, isym=isym,
                 ihc=box.ihc[uniq], cl1=box.cl1[uniq], cl2=box.cl2[uniq], hwva=box.hwva[uniq],
                 top=box.top[sel], bot=box.bot[sel], area=box.area[sel], k11=box.k11[sel], k33=box.k33[sel],
                 icelltype=box.icelltype[sel], strt=box.strt[sel], ibound=box.ibound[sel], ibotnode=ibot,
                 ss=box.ss[sel], sy=box.sy[sel], iconvert=box.iconvert[sel],
                 icellavg=box.icellavg, inewton=box.inewton, inewtonur=box.inewtonur, iperched=box.iperched,
                 ivarcv=box.ivarcv, idewatcv=box.idewatcv, insto=box.insto, istor_coef=box.istor_coef,
                 iconf_ss=box.iconf_ss, iorig_ss=box.iorig_ss, shape=(nl, i1 - i0, j1 - j0))
    gid = gid_box[sel].astype(np.int32)
    # neighbours: recv ranges follow the (owner, gid) order of the halo; send lists are the owned cells on
    # the face towards each neighbour in ascending global id (the neighbour's halo order)
    howner = owner[hidx]
    nbrs = np.unique(howner)
    recv_ptr = np.concatenate([[0], np.cumsum([(howner == q).sum() for q in nbrs])]).astype(np.int32)
    send_idx, send_ptr = [], [0]
    kk, ii, jj = kb[oidx], ib[oidx], jb[oidx]
    for q in nbrs:
        qi, qj = divmod(int(q), pc)
        if qi < bi:
            face = ii == i0
        elif qi > bi:
            face = ii == i1 - 1
        elif qj < bj:
            face = jj == j0
        else:
            face = jj == j1 - 1
        send_idx.append(np.nonzero(face)[0])
        send_ptr.append(send_ptr[-1] + send_idx[-1].size)
    send_idx = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, np.int32)
    return SubModel(rank=rank, nranks=nranks, model=m, n_own=n_own, global_id=gid,
                    nbr_rank=nbrs.astype(np.int32), send_ptr=np.asarray(send_ptr, np.int32), send_idx=send_idx,
                    recv_ptr=recv_ptr, block=(i0, i1, j0, j1), spec=spec)


@dataclass
class GenericSubModel(SubModel):
    """submodel cut out of an arbitrary global model by an owner map (extract_submodel)"""
    g2l_own: np.ndarray = None      # [global nodes] local index of the owned cells, -1 elsewhere

    def local_nodes(self, gnodes):
        loc = self.g2l_own[np.asarray(gnodes, dtype=np.int64)]
        mask = loc >= 0
        return mask, loc[mask].astype(np.int32)


def extract_submodel(g, owner, rank, nranks):
    """The submodel of `rank` from a global (e.g. merge_models) GwfModel and the owner rank of every cell --
    what `mf6 -p` gets from GWF-GWF exchanges: the rank's own model plus the exchange partners' cells as a
    halo (SpatialModelConnection.f90:306-510, the interface model), with the synchronisation lists of
    VirtualGwfModel / MpiRouter.  Same layout rules as build_dis_block: owned cells first in ascending global
    id, halo cells grouped by owner rank in ascending global id, halo rows diagonal only, per-connection
    arrays (cl1 = the lower GLOBAL id's side) taken unchanged from the global model."""
    owner = np.asarray(owner)
    n = g.nodes
    ia_g = g.ia.astype(np.int64)
    own = np.nonzero(owner == rank)[0]
    n_own = own.size
    cnt = (ia_g[own + 1] - ia_g[own])
    ia = np.zeros(n_own + 1, dtype=np.int64)
    np.cumsum(cnt, out=ia[1:])
    nja_own = int(ia[-1])
    within = np.arange(nja_own) - np.repeat(ia[:-1], cnt)
    pos = np.repeat(ia_g[own], cnt) + within
    cols = g.ja[pos].astype(np.int64)
    rows_l = np.repeat(np.arange(n_own), cnt)
    foreign = owner[cols] != rank
    halo_g = np.unique(cols[foreign])
    halo_g = halo_g[np.lexsort((halo_g, owner[halo_g]))]
    n_halo = halo_g.size
    n_ext = n_own + n_halo
    g2l = np.full(n, -1, dtype=np.int64)
    g2l[own] = np.arange(n_own)
    g2l_own = g2l.copy()
    g2l[halo_g] = n_own + np.arange(n_halo)
    ja = g2l[cols]
    diag = within == 0
    jas_old = g.jas[pos].astype(np.int64)
    uniq, inv = np.unique(jas_old[~diag], return_inverse=True)
    jas = np.full(nja_own, -1, dtype=np.int64)
    jas[~diag] = inv
    ia_full = np.concatenate([ia, nja_own + np.arange(1, n_halo + 1)])
    ja_full = np.concatenate([ja, n_own + np.arange(n_halo)])
    jas_full = np.concatenate([jas, np.full(n_halo, -1, dtype=np.int64)])
    rows = np.concatenate([rows_l, n_own + np.arange(n_halo)]).astype(np.int64)
    key = rows * n_ext + ja_full
    order = np.argsort(key, kind="stable")
    tkey = ja_full * n_ext + rows
    loc = np.minimum(np.searchsorted(key[order], tkey), key.size - 1)
    isym = np.where(key[order][loc] == tkey, order[loc], 0)
    sel = np.concatenate([own, halo_g])
    ibot = g2l[g.ibotnode[sel]]
    ibot = np.where(ibot < 0, np.arange(n_ext), ibot)
    m = GwfModel(nodes=n_ext, ia=ia_full, ja=ja_full, jas=jas_full, isym=isym,
                 ihc=g.ihc[uniq], cl1=g.cl1[uniq], cl2=g.cl2[uniq], hwva=g.hwva[uniq],
                 top=g.top[sel], bot=g.bot[sel], area=g.area[sel], k11=g.k11[sel], k33=g.k33[sel],
                 icelltype=g.icelltype[sel], strt=g.strt[sel], ibound=g.ibound[sel], ibotnode=ibot,
                 ss=g.ss[sel], sy=g.sy[sel], iconvert=g.iconvert[sel],
                 icellavg=g.icellavg, inewton=g.inewton, inewtonur=g.inewtonur, iperched=g.iperched,
                 ivarcv=g.ivarcv, idewatcv=g.idewatcv, insto=g.insto, istor_coef=g.istor_coef,
                 iconf_ss=g.iconf_ss, iorig_ss=g.iorig_ss, ithickstrt=g.ithickstrt, shape=None)
    howner = owner[halo_g]
    nbrs = np.unique(howner)
    recv_ptr = np.concatenate([[0], np.cumsum([(howner == q).sum() for q in nbrs])]).astype(np.int32)
    send_idx, send_ptr = [], [0]
    for q in nbrs:
        face = np.unique(rows_l[owner[cols] == q])       # owned cells next to q's cells, ascending global id
        send_idx.append(face)
        send_ptr.append(send_ptr[-1] + face.size)
    send_idx = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, np.int32)
    return GenericSubModel(rank=rank, nranks=nranks, model=m, n_own=n_own, global_id=sel.astype(np.int32),
                           nbr_rank=nbrs.astype(np.int32), send_ptr=np.asarray(send_ptr, np.int32),
                           send_idx=send_idx, recv_ptr=recv_ptr, block=None, spec=None, g2l_own=g2l_own)


def global_packages_c2(spec):
    """CHD 48 / 40 on the first / last column, WEL -1000 in the centre of the middle layer (C2 recipe)."""
    kk, ii = np.meshgrid(np.arange(spec.nlay), np.arange(spec.nrow), indexing="ij")
    west = ((kk * spec.nrow + ii) * spec.ncol).reshape(-1)
    east = west + spec.ncol - 1
    chd = Package(T.PKG_CHD, np.concatenate([west, east]),
                  np.concatenate([np.full(west.size, 48.0), np.full(east.size, 40.0)]))
    wn = ((spec.nlay // 2) * spec.nrow + spec.nrow // 2) * spec.ncol + spec.ncol // 2
    wel = Package(T.PKG_WEL, [wn], [-1000.0])
    return [chd, wel]


# ----------------------------------------------------------------------------------------------
class GpuComm:
    """NCCL communicator of libmf6gpu; the 128-byte unique id is broadcast with torch.distributed."""

    def __init__(self, rank=0, nranks=1):
        ensure_init()
        self._L = load()
        self.rank, self.nranks = rank, nranks
        self.p2p = False
        self.h = C.c_void_p()
        if nranks > 1:
            import torch
            import torch.distributed as dist
            buf = (C.c_ubyte * 128)()
            if rank == 0:
                check(self._L.mf6gpu_comm_unique_id(buf))
            t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, 0)
            raw = bytes(t.cpu().tolist())
            buf = (C.c_ubyte * 128).from_buffer_copy(raw)
            check(self._L.mf6gpu_comm_create(nranks, rank, buf, C.byref(self.h)))
        else:
            check(self._L.mf6gpu_comm_create(1, 0, None, C.byref(self.h)))

    def enable_p2p(self, halo_doubles):
        """Map every rank's mailbox with CUDA IPC so that halo messages and the small all-gathers move by
        direct peer stores over NVLink instead of NCCL calls (MF6GPU_P2P=0 keeps NCCL).  `halo_doubles`
        = this rank's largest halo message; the maximum over ranks sizes the mailboxes."""
        import os
        if self.nranks == 1 or os.environ.get("MF6GPU_P2P", "1") == "0" or self.p2p:
            return self.p2p
        import torch
        import torch.distributed as dist
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([int(halo_doubles)], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cap = int(t.item())
        buf = (C.c_ubyte * 64)()
        ok = self._L.mf6gpu_comm_p2p_export(self.h, cap, buf) == 0
        handles = [None] * self.nranks
        dist.all_gather_object(handles, bytes(buf) if ok else None)
        if any(h is None for h in handles):
            return False
        allh = (C.c_ubyte * (64 * self.nranks)).from_buffer_copy(b"".join(handles))
        ok = self._L.mf6gpu_comm_p2p_import(self.h, allh) == 0
        flags = [None] * self.nranks
        dist.all_gather_object(flags, bool(ok))
        self.p2p = all(flags)
        if not self.p2p:
            self._L.mf6gpu_comm_p2p_disable(self.h)
        if not self.p2p and self.rank == 0:
            print("modflow6_b200: CUDA IPC peer mapping unavailable (" +
                  self._L.mf6gpu_last_error().decode("utf-8", "replace") + "); using NCCL send/recv + all-gather",
                  flush=True)
        return self.p2p

    def destroy(self):
        if self.h:
            self._L.mf6gpu_comm_destroy(self.h)
            self.h = C.c_void_p()


class GpuDistributedSolution:
    """One rank of the split-model solution (same interface as GpuNumericalSolution; `x` holds the
    heads of the OWNED cells in local order, reports carry GLOBAL iteration counts and budgets)."""

    def __init__(self, sub, sln_settings, ims_settings, comm):
        ensure_init()
        self._L = load()
        self.sub = sub
        self.comm = comm
        self.model = sub.model
        self._ms = sub.model.struct()
        self.n = sub.n_own
        self.h = C.c_void_p()
        self._keep = [T.as_i32(sub.nbr_rank), T.as_i32(sub.send_ptr), T.as_i32(sub.send_idx),
                      T.as_i32(sub.recv_ptr), T.as_i32(sub.global_id)]
        k = self._keep
        msg = max([0] + list(np.diff(sub.send_ptr)) + list(np.diff(sub.recv_ptr)))
        comm.enable_p2p(int(msg))
        check(self._L.mf6gpu_solution_create_dist(C.byref(self._ms), C.byref(sln_settings), C.byref(ims_settings),
                                                  comm.h, sub.n_own, k[0].size, T.ptr_i32(k[0]), T.ptr_i32(k[1]),
                                                  T.ptr_i32(k[2]), T.ptr_i32(k[3]), T.ptr_i32(k[4]),
                                                  C.byref(self.h)))

    def set_packages(self, global_pkgs):
        self._pkgs = self.sub.localize_packages(global_pkgs)
        check(self._L.mf6gpu_solution_set_packages(self.h, len(self._pkgs), package_array(self._pkgs)))

    def timestep(self, kper=1, kstp=1, delt=1.0, iss=1):
        rep = T.StepReport()
        check(self._L.mf6gpu_solution_timestep(self.h, int(kper), int(kstp), float(delt), int(iss), C.byref(rep)))
        return rep

    @property
    def x(self):
        a = np.empty(self.n)
        check(self._L.mf6gpu_solution_get_x(self.h, T.ptr_f64(a)))
        return a

    def reset_x(self):
        check(self._L.mf6gpu_solution_reset_x(self.h))

    def stat(self, what):
        return self._L.mf6gpu_solution_stat(self.h, what)

    def solver_stat(self, what):
        return self._L.mf6gpu_solver_stat(self._L.mf6gpu_solution_solver(self.h), what)

    def profile(self, enable=True):
        check(self._L.mf6gpu_solver_profile(self._L.mf6gpu_solution_solver(self.h), 1 if enable else 0))

    def profile_result(self):
        out = {}
        sv = self._L.mf6gpu_solution_solver(self.h)
        for i, name in enumerate(("spmv", "ilu0_apply", "update", "dot", "direction", "factor")):
            ms, cnt = C.c_double(), C.c_int64()
            check(self._L.mf6gpu_solver_profile_get(sv, i, C.byref(ms), C.byref(cnt)))
            out[name] = (ms.value, cnt.value)
        return out

    def destroy(self):
        if self.h:
            self._L.mf6gpu_solution_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
