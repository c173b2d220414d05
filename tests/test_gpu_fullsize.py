"""Parity at the FULL size of the BASELINE configs: the device solve against the committed oracle fixtures
(tests/golden/*_full_*.npz, made on the CPU by tests/golden/make_golden_full.py with the elimination order the
device uses).  north_star bar: max |dhead| <= 0.1 x OUTER_DVCLOSE, budget percent discrepancy within 1e-3."""
import numpy as np
import pytest

from modflow6_b200 import configs, ctypes_types as T, lib

pytestmark = pytest.mark.gpu


def _run(cfg, max_steps=None):
    from modflow6_b200.solution import GpuNumericalSolution
    G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
    assert np.array_equal(G.elimination_order(), lib.model_elimination_order(cfg.model, cfg.ims.gpu_ordering))
    reps = configs.run_simulation(G, cfg, max_steps=max_steps)
    x = G.x
    G.destroy()
    return reps, x


@pytest.mark.parametrize("closure", ["tight2", "tight"])
def test_c2_full_size_against_oracle_fixture(gpu, closure):
    """BASELINE config 2 (10 x 1000 x 1000, CG + ILU0, block ordering): 1e7 heads against the oracle's on the same
    permuted system, north-star bar (max |dhead| <= 0.1 x OUTER_DVCLOSE, budget within 1e-3), at the closure
    bench.py times ("tight2") and one decade looser ("tight"); see configs.C2_CLOSURE"""
    from oracle import golden
    tag = "c2_full_block_" + closure
    if golden.load(tag) is None:
        pytest.skip("fixture missing")
    cfg = configs.c2_confined(closure=closure)
    reps, x = _run(cfg)
    c = golden.compare_heads(tag, x, cfg.sln.dvclose)
    assert reps[0]["converged"] == 1
    assert c["max_abs_dhead"] <= 0.1 * cfg.sln.dvclose, c
    if "max_abs_dblocksum" in c:
        assert c["max_abs_dblocksum"] <= 1000 * 0.1 * cfg.sln.dvclose
    assert abs(reps[0]["pdiffr"] - c["oracle"]["pdiffr"]) <= 1e-3
    assert reps[0]["outer_iterations"] == c["oracle"]["outer_iterations"]
    assert abs(reps[0]["inner_iterations"] - c["oracle"]["inner_iterations"]) <= c["oracle"]["inner_iterations"] // 20
    # the reference's own (natural) ordering at the same closure: a different convergence path to the same answer
    n = golden.compare_heads("c2_full_natural_" + closure, x, cfg.sln.dvclose)
    if n is not None:
        assert abs(reps[0]["pdiffr"] - n["oracle"]["pdiffr"]) <= 1e-3
        if closure == "tight2":      # the orderings themselves agree to 5.4e-7 here: the bar holds across them
            assert n["max_abs_dhead"] <= 0.1 * cfg.sln.dvclose, n
        else:                        # 9.6e-6 between the oracle's own two orderings at this closure
            assert n["max_abs_dhead"] <= cfg.sln.dvclose, n


def test_c2_full_size_survey_closure_slack(gpu):
    """the same model with the closure SURVEY.md names (inner 1e-6 / 1e-2, one decade below OUTER_DVCLOSE): the CG
    step-size test stops ~1e-4 short of the converged heads, so this is a characterisation of closure slack, not
    the parity claim: iteration counts and budget agree with the oracle, heads within OUTER_DVCLOSE (measured
    8.7e-6; the oracle's own two orderings are 5.7e-5 apart at this closure)"""
    from oracle import golden
    if golden.load("c2_full_block") is None:
        pytest.skip("fixture missing")
    cfg = configs.c2_confined(closure="survey")
    reps, x = _run(cfg)
    c = golden.compare_heads("c2_full_block", x, cfg.sln.dvclose)
    assert reps[0]["converged"] == 1
    assert c["max_abs_dhead"] <= cfg.sln.dvclose, c
    assert abs(reps[0]["pdiffr"] - c["oracle"]["pdiffr"]) <= 1e-3
    assert reps[0]["outer_iterations"] == c["oracle"]["outer_iterations"]
    assert abs(reps[0]["inner_iterations"] - c["oracle"]["inner_iterations"]) <= c["oracle"]["inner_iterations"] // 20


def test_c3_full_size_first_step_against_oracle_fixture(gpu, capsys):
    """BASELINE config 3 (5 x 2000 x 2000 Newton, BiCGSTAB, DBD under-relaxation, pseudo-transient continuation): the
    steady first step, 2e7 heads, against the oracle's 3.3-hour run on the same permuted system.

    At the closure SURVEY.md names (INNER_DVCLOSE 1e-6 / INNER_RCLOSE 1e-5, OUTER_DVCLOSE 1e-4) the damped Newton
    iteration stops well short of the converged heads -- the oracle's own budget is 0.028 % off and its two
    orderings are 2e-4 apart already on a 5 x 500 x 500 grid (profiles/r02_c3_closure_study.json) -- so head
    differences here measure closure slack (measured 1.5e-3; 10 outer / 8242 inner iterations against the oracle's
    10 / 7653), not parity: the parity bar is held by
    test_c3_tight_closure_against_oracle_fixture.  Asserted at this closure: convergence, the outer iteration count
    near the oracle's, heads within the characterised slack, a budget as good as the oracle's."""
    import json
    from oracle import golden
    if golden.load("c3_full_block") is None:
        pytest.skip("fixture missing")
    cfg = configs.c3_newton()
    reps, x = _run(cfg, max_steps=1)
    c = golden.compare_heads("c3_full_block", x, cfg.sln.dvclose)
    with capsys.disabled():
        print("\nC3_FULL " + json.dumps({"device": {k: reps[0][k] for k in ("converged", "outer_iterations",
                                                                            "inner_iterations", "pdiffr")},
                                         "compare": c}))
    assert reps[0]["converged"] == 1
    assert c["max_abs_dhead"] <= 50 * cfg.sln.dvclose, c
    assert abs(reps[0]["outer_iterations"] - c["oracle"]["outer_iterations"]) <= 2
    # budget percent discrepancy of the converged step: 0.16 % here, 0.028 % in the oracle run -- how far OUTER_DVCLOSE
    # 1e-4 leaves the damped Newton iterate from the balanced solution (one decade tighter: 1.0e-4 % vs 0.9e-4 %)
    assert abs(reps[0]["pdiffr"]) <= 0.5


@pytest.mark.parametrize("size", [(5, 500, 500), (5, 1000, 1000)])
def test_c3_tight_closure_against_oracle_fixture(gpu, size, capsys):
    """config 3 (Newton, BiCGSTAB + ILU0, DBD, pseudo-transient continuation), steady first step, at 1.25e6 and 5e6
    cells (measured 1.2e-7 and 1.4e-6 from the oracle: the distance grows ~12 x per 4 x cells, so at the full 2e7
    cells this closure would no longer hold the 1e-5 bar and a tighter one runs BiCGSTAB into stagnation -- see
    DESIGN.md section 5), with the inner closure one decade
    tighter (configs.tighten_inner_closure level 1: 1e-7 / 1e-7; the oracle's own two orderings then agree to 3.6e-7,
    profiles/r02_c3_closure_study.json).  North-star bar: max |dhead| <= 0.1 x OUTER_DVCLOSE against the oracle on the
    same permuted system AND against the reference's own natural-order solve, budget within 1e-3."""
    import json
    from oracle import golden
    tag = "c3_%dx%dx%d_block_tight" % size
    if golden.load(tag) is None:
        pytest.skip("fixture missing")
    cfg = configs.tighten_inner_closure(configs.c3_newton(*size), 1)
    reps, x = _run(cfg, max_steps=1)
    c = golden.compare_heads(tag, x, cfg.sln.dvclose)
    n = golden.compare_heads(tag.replace("block", "natural"), x, cfg.sln.dvclose)
    with capsys.disabled():
        print("\nC3_TIGHT " + json.dumps({"size": size, "device": {k: reps[0][k] for k in (
            "converged", "outer_iterations", "inner_iterations", "pdiffr")}, "compare": c, "natural": n}))
    assert reps[0]["converged"] == 1
    assert c["max_abs_dhead"] <= 0.1 * cfg.sln.dvclose, c
    assert abs(reps[0]["pdiffr"] - c["oracle"]["pdiffr"]) <= 1e-3
    assert reps[0]["outer_iterations"] == c["oracle"]["outer_iterations"]
    if n is not None:
        # across orderings: the oracle's own two orderings are 3.6e-7 apart at 1.25e6 cells and 5.4e-6 at 5e6 (the
        # conditioning grows with the grid), so the 0.1 x bar is asserted across orderings at the smallest size only
        assert n["max_abs_dhead"] <= (0.1 if size == (5, 500, 500) else 1.0) * cfg.sln.dvclose, n
        assert abs(reps[0]["pdiffr"] - n["oracle"]["pdiffr"]) <= 1e-3
    # size-independent property: the device heads satisfy the REFERENCE's discrete equations (the oracle formulates
    # the system at these heads; every cell's flow imbalance) as well as the oracle's own heads do
    ores = golden.load(tag)["meta"].get("residual")
    if ores and x.size <= 5000000:       # (the oracle-side formulate of 2e7 cells takes a minute)
        res = golden.nonlinear_residual(cfg, x)
        with capsys.disabled():
            print("C3_TIGHT_RESIDUAL " + json.dumps({"size": size, "device_heads": res, "oracle_heads": ores}))
        assert res["max_abs"] <= max(3 * ores["max_abs"], cfg.ims.rclose), (res, ores)
        assert res["l2"] <= 3 * ores["l2"] + 1e-12, (res, ores)
