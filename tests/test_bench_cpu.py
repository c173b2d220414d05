"""bench.py contract pieces that run without a GPU: the reference arm (the CPU oracle timed on a bounded
sample) prints one JSON line with the keys the driver reads; the product arm refuses to run without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "3,20,30",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ims_cell_iterations_per_second"
    assert line["unit"] == "cell-iter/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["steps"] == 2 and line["n_gpus"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["config"]["cells"] == 1800


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--size", "3,20,30", "--steps", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0            # no CPU fallback: the measurement cannot silently run on the host
