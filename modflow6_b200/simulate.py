"""Run a MODFLOW 6 simulation directory through the accelerated path: the time loop of
`Mf6DoTimestep` (src/mf6core.f90:620-660) around `mf6gpu_solution_timestep`, with the input subset read by
`mf6io.py` and heads / budgets written in the reference's binary formats by `output.py`.

    python -m modflow6_b200.simulate <simulation directory> [--ordering block|multicolor|natural]

There is no CPU fallback: without a GPU the solution object cannot be created and the run fails.  (`run` takes
the solution class as a parameter so that the tests can drive the same reader / time loop / writers with the
CPU oracle and check them against the reference's known answers.)
"""
import argparse
import json
import sys

import numpy as np

from . import ctypes_types as T
from .grid import merge_models, tdis_steps
from .mf6io import Mf6InputError, TimeSeriesError, read_simulation
from .output import PKG_TEXT, BudgetCsvWriter, BudgetFileWriter, HeadFileWriter, ListingFileWriter, write_grb

DHNOFLO = 1.0e30   # Constants.f90: head written for cells outside the active domain


def _should_save(settings, kstp, nstp):
    """OC `SAVE <rtype> <ocsetting>`: ALL | FIRST | LAST | FREQUENCY n | STEPS n1 n2 ... (gwf-oc.dfn)"""
    for s in settings:
        if not s or s[0] == "ALL":
            return True
        if s[0] == "FIRST" and kstp == 1:
            return True
        if s[0] == "LAST" and kstp == nstp:
            return True
        if s[0] == "FREQUENCY" and int(s[1]) > 0 and kstp % int(s[1]) == 0:
            return True
        if s[0] == "STEPS" and kstp in [int(v) for v in s[1:]]:
            return True
    return False


class _ModelView:
    """what output.write_step needs for ONE model of a multi-model solution: the model's own FLOW-JA-FACE (the
    entries of the merged matrix whose row and column both belong to it -- same relative order as the model's
    own CSR, the diagonal carrying the residual that includes the exchange flows, like gwf_gwf_add_to_flowja),
    its storage rates and the rates of its own packages"""

    def __init__(self, solution, merged, model, lo, pkg_index):
        self.s, self.model, self.lo, self.pkg_index = solution, model, int(lo), pkg_index
        rows = np.repeat(np.arange(merged.nodes), np.diff(merged.ia))
        hi = self.lo + model.nodes
        self.pick = np.nonzero((rows >= self.lo) & (rows < hi) & (merged.ja >= self.lo) & (merged.ja < hi))[0]
        assert self.pick.size == model.nja

    @property
    def flowja(self):
        return self.s.flowja[self.pick]

    @property
    def storage_rates(self):
        ss, sy = self.s.storage_rates
        return ss[self.lo:self.lo + self.model.nodes], sy[self.lo:self.lo + self.model.nodes]

    @property
    def simvals(self):
        sv = self.s.simvals
        return [sv[i] for i in self.pkg_index]


def _exchange_rates(merged, flowja, offs, e):
    """simvals of a GWF-GWF exchange = flow into the model-1 cell from the model-2 cell (gwf_gwf_calc_simvals):
    the FLOW-JA-FACE entry (n1, n2) of the merged system"""
    a = e["nodem1"] + int(offs[e["m1"]])
    b = e["nodem2"] + int(offs[e["m2"]])
    q = np.empty(a.size)
    for i, (r, c) in enumerate(zip(a, b)):
        p0, p1 = merged.ia[r], merged.ia[r + 1]
        q[i] = flowja[p0 + int(np.nonzero(merged.ja[p0:p1] == c)[0][0])]
    return q


def _user_grid(gi, h):
    """heads of the model's cells on the user grid: cells that IDOMAIN removes carry 1e30 (record_array, dinact)"""
    if gi.nodeuser is None:
        return h
    full = np.full(gi.nodesuser, DHNOFLO)
    full[gi.nodeuser] = h
    return full


def run(sim_dir, ordering=T.ORDER_BLOCK_MULTICOLOR, write_output=True, log=None, solution_class=None, comm=None):
    """solution_class(model, sln_settings, ims_settings) -> object with set_packages / timestep / x / flowja /
    simvals / storage_rates; default GpuNumericalSolution.

    comm (distributed.GpuComm, one process per GPU): the split-model run of `mf6 -p` -- model k of the
    simulation lives on rank k, its exchange partners' cells are the halo (distributed.extract_submodel);
    every rank writes its own model's head file.  The deck must hold exactly one model per rank."""
    sim = read_simulation(sim_dir)
    log = log or (lambda *a: None)
    for w in sim.warnings:
        log("warning:", w)
    models = [gi.model for gi in sim.models]
    if len(models) == 1 and not sim.exchanges:
        model, offs = models[0], np.array([0, models[0].nodes])
    else:
        try:
            model, offs = merge_models(models, sim.exchanges)
        except ValueError as e:
            raise Mf6InputError(str(e)) from None
    sim.ims.gpu_ordering = ordering
    rank = None
    if comm is not None and comm.nranks > 1:
        from .distributed import GpuDistributedSolution, extract_submodel
        if len(models) != comm.nranks:
            raise ValueError(f"{len(models)} models for {comm.nranks} ranks: the split-model run needs one model per rank")
        rank = comm.rank
        owner = np.repeat(np.arange(len(models)), [m.nodes for m in models])
        S = GpuDistributedSolution(extract_submodel(model, owner, rank, comm.nranks), sim.sln, sim.ims, comm)
    else:
        if solution_class is None:
            from .solution import GpuNumericalSolution as solution_class
        S = solution_class(model, sim.sln, sim.ims)
    # ghost node corrections of the models (GNC6 in a name file) and of the exchanges (GNC6 FILEIN in a GWF-GWF
    # exchange: cellidn and the contributing cells in model 1, cellidm in model 2), in merged node numbers
    gncs = []
    for k, gi in enumerate(sim.models):
        if gi.gnc is not None:
            n_, m_, j_, a_ = gi.gnc
            gncs.append((n_ + offs[k], m_ + offs[k], np.where(j_ >= 0, j_ + offs[k], -1), a_))
    for e in sim.exchanges:
        if e.get("gnc") is not None:
            n_, m_, j_, a_ = e["gnc"]
            gncs.append((n_ + offs[e["m1"]], m_ + offs[e["m2"]], np.where(j_ >= 0, j_ + offs[e["m1"]], -1), a_))
    if gncs:
        if rank is not None:
            raise Mf6InputError("GNC6 is not available in the split-model run")
        numj = max(g[2].shape[1] for g in gncs)
        pad = lambda a, fill: np.pad(a, ((0, 0), (0, numj - a.shape[1])), constant_values=fill)   # noqa: E731
        S.set_gnc(np.concatenate([g[0] for g in gncs]), np.concatenate([g[1] for g in gncs]),
                  np.concatenate([pad(g[2], -1) for g in gncs]), np.concatenate([pad(g[3], 0.0) for g in gncs]))
    writers = []
    for k, gi in enumerate(sim.models):
        mine = rank is None or rank == k
        if write_output and mine and gi.grid is not None and not gi.grid["nogrb"]:
            write_grb(gi.grid["file"] or gi.grid["default"], gi.grid, gi.model, gi.nodeuser)   # dis_ar
        hw = HeadFileWriter(gi.head_file, gi.shape) if (write_output and gi.head_file and mine) else None
        bw = None
        if write_output and gi.budget_file and mine:
            if rank is not None:
                log(f"warning: {gi.name}: budget files are not written by the split-model run")
            else:
                bw = BudgetFileWriter(gi.budget_file, gi.shape, gi.name)
        lw = None
        if write_output and gi.list_file and mine and rank is None and any(
                r[0] == "BUDGET" for recs in gi.printrec.values() for r in recs):
            lw = ListingFileWriter(gi.list_file, gi.name, sim.time_units)
        cw = BudgetCsvWriter(gi.budgetcsv_file) if (write_output and gi.budgetcsv_file and mine and rank is None) else None
        writers.append((hw, bw, lw, cw))
    current = [[None] * len(gi.packages) for gi in sim.models]     # list in force per package
    saving = [dict() for _ in sim.models]                          # rtype -> settings in force
    printing = [dict() for _ in sim.models]
    reports, totim = [], 0.0
    hfb_now = {}                                                   # model -> barrier list in force
    stopped = False
    for kper in range(1, sim.nper + 1):
        perlen, nstp, tsmult = sim.perioddata[kper - 1]
        pkgs, owner = [], []
        for k, gi in enumerate(sim.models):
            for ip, sp in enumerate(gi.packages):
                if kper in sp.periods:
                    current[k][ip] = sp.periods[kper]
                p = current[k][ip]
                if p is not None:
                    pkgs.append(p.with_nodes(p.nodelist + int(offs[k])))
                    owner.append((k, ip))
            if kper in gi.save:
                saving[k] = {}
                for rtype, st in gi.save[kper]:
                    saving[k].setdefault(rtype, []).append(st)
            if kper in gi.printrec:
                printing[k] = {}
                for rtype, st in gi.printrec[kper]:
                    printing[k].setdefault(rtype, []).append(st)
        # a model without STO is steady; with STO the period keeps the last STEADY-STATE / TRANSIENT keyword,
        # TRANSIENT before any PERIOD block (gwf-sto.f90:170-182, 756)
        iss_of = []
        for gi in sim.models:
            if gi.model.insto:
                upto = [p for p in gi.sto_transient if p <= kper]
                iss_of.append(0 if (not upto or gi.sto_transient[max(upto)]) else 1)
        if len(set(iss_of)) > 1:     # one solution matrix carries one steady / transient state (see merge_models)
            raise Mf6InputError(f"period {kper}: the models of the solution disagree on STEADY-STATE / TRANSIENT")
        iss = iss_of[0] if iss_of else 1
        # AUXMULTNAME without a time series is a constant factor; with time series (TS6) the lists are re-evaluated
        # every time step (tsmgr_ad at the start of the step: begin = totim, end = totim + delt)
        pkgs = [p if p.time_dependent else p.at_time(totim, totim) for p in pkgs]
        timed = any(p.time_dependent for p in pkgs)
        templates = pkgs
        if not timed:
            S.set_packages(pkgs)
        if any(kper in gi.hfb for gi in sim.models):   # hfb_rp: a PERIOD block replaces the model's barrier list
            for k, gi in enumerate(sim.models):
                if kper in gi.hfb:
                    hfb_now[k] = gi.hfb[kper]
            if rank is not None:
                raise Mf6InputError("HFB6 is not available in the split-model run")
            lists = [(a + int(offs[k]), c + int(offs[k]), h) for k, (a, c, h) in sorted(hfb_now.items())]
            S.set_hfb(*(np.concatenate([l[i] for l in lists]) for i in range(3)))
        pertim = 0.0
        for kstp, delt in enumerate(tdis_steps(perlen, nstp, tsmult), start=1):
            if timed:
                try:
                    pkgs = [p.at_time(totim, totim + delt) for p in templates]
                except TimeSeriesError as e:         # e.g. a series that ends before the simulation does
                    raise Mf6InputError(f"period {kper} step {kstp}: {e}") from None
                S.set_packages(pkgs)
            rep = S.timestep(kper, kstp, delt, iss)
            pertim += delt
            totim += delt
            d = rep.as_dict()
            d.update(kper=kper, kstp=kstp, delt=delt, totim=totim)
            reports.append(d)
            log(f"period {kper} step {kstp}: outer {d['outer_iterations']} inner {d['inner_iterations']} "
                f"converged {d['converged']} budget discrepancy {d['pdiffr']:.3e} %")
            x = S.x                      # split-model run: the owned cells = this rank's model
            for k, gi in enumerate(sim.models):
                hw, bw = writers[k][:2]
                if hw and _should_save(saving[k].get("HEAD", []), kstp, nstp):
                    h = (x if rank is not None else x[offs[k]:offs[k] + gi.model.nodes]).copy()
                    h[gi.model.ibound == 0] = DHNOFLO
                    hw.write(kstp, kper, pertim, totim, _user_grid(gi, h))
                if bw and _should_save(saving[k].get("BUDGET", []), kstp, nstp):
                    mine = [i for i, (kk, _) in enumerate(owner) if kk == k]
                    view = S if len(models) == 1 else _ModelView(S, model, gi.model, offs[k], mine)
                    local = [pkgs[i].with_nodes(pkgs[i].nodelist - int(offs[k])) for i in mine]
                    bw.write_step(kstp, kper, delt, pertim, totim, view, local,
                                  [gi.packages[owner[i][1]].name for i in mine], nodeuser=gi.nodeuser)
            # the model budget table of the listing file (gwf_ot_bdsummary -> budget_ot, then tdis_ot)
            exq = {}
            for k, gi in enumerate(sim.models):
                lw, cw = writers[k][2:]
                if not (lw and _should_save(printing[k].get("BUDGET", []), kstp, nstp)):
                    lw = None
                if not (lw or cw):
                    continue
                mine = [i for i, (kk, _) in enumerate(owner) if kk == k]
                view = S if len(models) == 1 else _ModelView(S, model, gi.model, offs[k], mine)
                acc = lambda v: (float(v[v > 0].sum()), float(-v[v < 0].sum()))     # noqa: E731  rate_accumulator
                entries = []
                if gi.model.insto:
                    ss_, sy_ = view.storage_rates
                    entries.append(("STO-SS",) + acc(ss_) + ("STORAGE",))
                    if gi.model.iconvert is not None and np.any(gi.model.iconvert):
                        entries.append(("STO-SY",) + acc(sy_) + ("STORAGE",))
                sv = view.simvals
                for n_, i in enumerate(mine):
                    entries.append((PKG_TEXT[pkgs[i].type],) + acc(np.asarray(sv[n_])) + (gi.packages[owner[i][1]].name.upper(),))
                for e in sim.exchanges:            # gwf_gwf_bd: the exchange is a FLOW-JA-FACE entry of both budgets
                    if k in (e["m1"], e["m2"]):
                        if id(e) not in exq:
                            exq[id(e)] = _exchange_rates(model, S.flowja, offs, e)
                        q = exq[id(e)] if k == e["m1"] else -exq[id(e)]
                        entries.append(("FLOW-JA-FACE",) + acc(q) + (e["name"].upper(),))
                if lw:
                    lw.write_budget(kstp, kper, delt, pertim, totim, entries)
                if cw:
                    cw.write(totim, entries)             # every time step (gwf_ot_bdsummary)
            # exchange flows follow the models' own records (exg_ot after model_ot, mf6core.f90:755-771)
            for e in sim.exchanges:
                if not e["save_flows"]:
                    continue
                q = None
                for k, other, mine, theirs, sign in ((e["m1"], e["m2"], "nodem1", "nodem2", 1.0),
                                                     (e["m2"], e["m1"], "nodem2", "nodem1", -1.0)):
                    bw = writers[k][1]
                    if bw and _should_save(saving[k].get("BUDGET", []), kstp, nstp):
                        if q is None:
                            q = _exchange_rates(model, S.flowja, offs, e)
                        bw.write_exchange(kstp, kper, delt, pertim, totim, e["name"], sim.models[other].name,
                                          e["user" + mine], e["user" + theirs], sign * q, e["auxname"], e["aux"])
            # converge_check (Sim.f90:401-433): without CONTINUE in mfsim.nam a time step that did not converge ends
            # the simulation after its output has been written
            if not d["converged"] and not sim.continue_:
                log("Simulation convergence failure. Simulation will terminate after output and deallocation.")
                stopped = True
                break
        if stopped:
            break
    for ws in writers:
        for w_ in ws:
            if w_:
                w_.close()
    xf = np.array(S.x, copy=True)      # (a solution class may hand out a view of memory it owns)
    if rank is not None:
        heads = [_user_grid(sim.models[rank], xf).reshape(sim.models[rank].shape)]
    else:
        heads = [_user_grid(gi, xf[offs[k]:offs[k] + gi.model.nodes]).reshape(gi.shape)
                 for k, gi in enumerate(sim.models)]
    return dict(simulation=sim, reports=reports, heads=heads, solution=S)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("sim_dir")
    ap.add_argument("--ordering", default="block", choices=["natural", "multicolor", "block"])
    a = ap.parse_args(argv)
    o = {"natural": T.ORDER_NATURAL, "multicolor": T.ORDER_MULTICOLOR, "block": T.ORDER_BLOCK_MULTICOLOR}[a.ordering]
    out = run(a.sim_dir, o, log=lambda *s: print(*s, file=sys.stderr))
    ok = all(r["converged"] for r in out["reports"])
    print(json.dumps({"steps": len(out["reports"]), "converged": bool(ok),
                      "inner_iterations": int(sum(r["inner_iterations"] for r in out["reports"])),
                      "head_min": float(min(h.min() for h in out["heads"])),
                      "head_max": float(max(h.max() for h in out["heads"]))}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
