"""bench.py contract checks that need no GPU: the reference arm (the UNMODIFIED reference from
baseline/_ref/code on the host cores) prints one JSON line with the agreed keys, uses all host
threads even when the launcher exported OMP_NUM_THREADS=1 (torchrun does), and non-zero ranks stay
silent; our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "code", "spinn2d.py"))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True,
                          env=e, cwd=ROOT)


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref/code is created by __graft_entry__.build() where /root/reference exists")
def test_reference_arm_times_the_reference_itself_with_all_threads():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-chunk", "1024", "--samples", "65536"],
             env={"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["value"] > 0
    assert line["metric"].startswith("point-node evals/sec") and line["unit"] == "point-node evals/s"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and "baseline/_ref/code" in cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0)) > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["config"]["nodes"] == 4096 and line["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
