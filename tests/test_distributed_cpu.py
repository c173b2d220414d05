"""Host-side logic of the split-model (multi-GPU) path on CPU: block decomposition, halo plans and a
world_size-2 gloo emulation of the halo exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modflow6_b200 import ctypes_types as T
from modflow6_b200.distributed import GridSpec, build_dis_block, global_packages_c2, hash_normal

SPEC = GridSpec(nlay=3, nrow=11, ncol=13, ksigma=0.8, seed=5)


def _global():
    sub = build_dis_block(SPEC, 1, 1, 0)
    assert sub.n_own == sub.model.nodes == 3 * 11 * 13 and sub.nbr_rank.size == 0
    return sub


def test_hash_fields_do_not_depend_on_the_box():
    g = np.arange(1000, dtype=np.int64)
    a = hash_normal(g, 3)
    b = np.concatenate([hash_normal(g[:400], 3), hash_normal(g[400:], 3)])
    assert np.array_equal(a, b) and abs(a.mean()) < 0.15 and 0.85 < a.std() < 1.15


def test_single_block_equals_plain_dis_model():
    from modflow6_b200.grid import build_dis_model
    sub = _global()
    m = sub.model
    kk, ii, jj = np.meshgrid(np.arange(3), np.arange(11), np.arange(13), indexing="ij")
    k = np.exp(np.log(10.0) + 0.8 * hash_normal(SPEC.gid(kk, ii, jj), 5))
    ref = build_dis_model(3, 11, 13, 100.0, 100.0, 0.0, [-10.0, -20.0, -30.0], k, k33=0.1 * k, strt=44.0)
    assert np.array_equal(m.ia, ref.ia) and np.array_equal(m.ja, ref.ja)
    off = np.ones(m.nja, bool)
    off[m.ia[:-1]] = False
    assert np.array_equal(m.jas[off], ref.jas[off]) and np.array_equal(m.isym, ref.isym)
    for name in ("ihc", "cl1", "cl2", "hwva", "top", "bot", "area", "k11", "k33"):
        assert np.array_equal(getattr(m, name), getattr(ref, name)), name
    assert np.array_equal(sub.global_id, np.arange(m.nodes))


def _conn_table(sub):
    """{(gid_row, gid_col): (ihc, cl_row_side, cl_col_side, hwva)} for the owned rows"""
    m, gid = sub.model, sub.global_id
    out = {}
    for r in range(sub.n_own):
        for p in range(m.ia[r] + 1, m.ia[r + 1]):
            c, jj = m.ja[p], m.jas[p]
            lo_is_row = gid[r] < gid[c]
            cl_r, cl_c = (m.cl1[jj], m.cl2[jj]) if lo_is_row else (m.cl2[jj], m.cl1[jj])
            out[(int(gid[r]), int(gid[c]))] = (int(m.ihc[jj]), cl_r, cl_c, m.hwva[jj])
    return out


@pytest.mark.parametrize("pr,pc", [(1, 2), (2, 1), (2, 2), (3, 2)])
def test_block_decomposition_reproduces_the_global_connectivity(pr, pc):
    ref = _conn_table(_global())
    subs = [build_dis_block(SPEC, pr, pc, r) for r in range(pr * pc)]
    merged = {}
    owned = []
    for s in subs:
        t = _conn_table(s)
        assert not (set(t) & set(merged))
        merged.update(t)
        owned.append(s.global_id[:s.n_own])
        m = s.model
        # owned rows: diagonal first; halo rows: diagonal only; per-cell data follow the global recipe
        assert np.array_equal(m.ja[m.ia[:-1]], np.arange(m.nodes))
        assert np.all(np.diff(m.ia[s.n_own:]) == 1)
        g = _global().model
        for name in ("top", "bot", "k11", "k33", "area"):
            assert np.array_equal(getattr(m, name), getattr(g, name)[s.global_id]), name
        assert np.all(np.diff(s.global_id[:s.n_own]) > 0)
    assert merged == ref
    allowned = np.sort(np.concatenate(owned))
    assert np.array_equal(allowned, np.arange(SPEC.nlay * SPEC.nrow * SPEC.ncol))
    # halo plans are symmetric: what a sends to b is what b expects from a, in the same order
    for a in subs:
        for ka, b_rank in enumerate(a.nbr_rank):
            b = subs[b_rank]
            kb = int(np.nonzero(b.nbr_rank == a.rank)[0][0])
            sent = a.global_id[a.send_idx[a.send_ptr[ka]:a.send_ptr[ka + 1]]]
            expected = b.global_id[b.n_own + b.recv_ptr[kb]:b.n_own + b.recv_ptr[kb + 1]]
            assert np.array_equal(sent, expected)


def test_packages_are_localised_with_identical_structure():
    pk = global_packages_c2(SPEC)
    subs = [build_dis_block(SPEC, 2, 2, r) for r in range(4)]
    tot = [0, 0]
    for s in subs:
        loc = s.localize_packages(pk)
        assert [p.type for p in loc] == [T.PKG_CHD, T.PKG_WEL]
        for i, p in enumerate(loc):
            tot[i] += p.nodelist.size
            assert np.all(p.nodelist < s.n_own)
            mask, _ = s.local_nodes(pk[i].nodelist)
            assert np.array_equal(s.global_id[p.nodelist], pk[i].nodelist[mask])
    assert tot == [pk[0].nodelist.size, 1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sub = build_dis_block(SPEC, 1, world, rank)
        # emulate HaloPlan::exchange: vec[n_own + recv range] <- neighbour's vec[send_idx]
        f = lambda g: np.sin(0.37 * g) + 0.01 * g          # noqa: E731  value of the global field
        vec = np.zeros(sub.model.nodes)
        vec[:sub.n_own] = f(sub.global_id[:sub.n_own].astype(np.float64))
        reqs, bufs = [], []
        for k, nb in enumerate(sub.nbr_rank):
            out = torch.from_numpy(vec[sub.send_idx[sub.send_ptr[k]:sub.send_ptr[k + 1]]].copy())
            inn = torch.empty(int(sub.recv_ptr[k + 1] - sub.recv_ptr[k]), dtype=torch.float64)
            reqs += [dist.isend(out, int(nb)), dist.irecv(inn, int(nb))]
            bufs.append((k, inn))
        for r in reqs:
            r.wait()
        for k, inn in bufs:
            vec[sub.n_own + sub.recv_ptr[k]:sub.n_own + sub.recv_ptr[k + 1]] = inn.numpy()
        ok = np.array_equal(vec, f(sub.global_id.astype(np.float64)))
        # the packed all-gather of reduction records: rank order, identical result on every rank
        rec = torch.tensor([float(rank + 1), float(sub.n_own)], dtype=torch.float64)
        allrec = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allrec, rec)
        total = sum(float(t[1]) for t in allrec)
        q.put((rank, bool(ok), total))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_halo_exchange_and_allgather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == res[1][2] == SPEC.nlay * SPEC.nrow * SPEC.ncol


@pytest.mark.parametrize("pr,pc", [(1, 2), (2, 2), (3, 2)])
def test_generic_extraction_equals_the_block_builder(pr, pc):
    """extract_submodel (any global model + owner map: the GWF-GWF exchange route) must cut the same
    submodels out of the global DIS model as the dedicated block builder"""
    from modflow6_b200.distributed import block_ranges, extract_submodel
    g = _global().model
    rb, cb = block_ranges(SPEC.nrow, pr), block_ranges(SPEC.ncol, pc)
    k, i, j = np.unravel_index(np.arange(g.nodes), (SPEC.nlay, SPEC.nrow, SPEC.ncol))
    bi = np.searchsorted(np.array([r[1] for r in rb]), i, side="right")
    bj = np.searchsorted(np.array([c[1] for c in cb]), j, side="right")
    owner = bi * pc + bj
    pk = global_packages_c2(SPEC)
    for rank in range(pr * pc):
        a, b = build_dis_block(SPEC, pr, pc, rank), extract_submodel(g, owner, rank, pr * pc)
        assert a.n_own == b.n_own and np.array_equal(a.global_id, b.global_id)
        assert np.array_equal(a.nbr_rank, b.nbr_rank) and np.array_equal(a.recv_ptr, b.recv_ptr)
        assert np.array_equal(a.send_ptr, b.send_ptr) and np.array_equal(a.send_idx, b.send_idx)
        ta, tb = _conn_table(a), _conn_table(b)
        assert ta.keys() == tb.keys()
        for key in ta:
            assert ta[key] == tb[key], key
        for name in ("top", "bot", "area", "k11", "k33", "strt", "ibound", "icelltype"):
            assert np.array_equal(getattr(a.model, name), getattr(b.model, name)), name
        for pa, pb in zip(a.localize_packages(pk), b.localize_packages(pk)):
            assert np.array_equal(pa.nodelist, pb.nodelist) and np.array_equal(pa.b1, pb.b1)
