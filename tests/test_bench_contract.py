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


def test_roofline_helpers_read_the_committed_ncu_capture():
    """roofline.traffic comes from the committed `ncu --set full` capture of the BASELINE workload
    (profiles/r02_ncu_kernels.json, tools/ncu_summary.py json), never from a source constant; the
    per-kernel report turns a duration into fractions of the nominal FP32 / MUFU pipes."""
    sys.path.insert(0, ROOT)
    import bench
    t = bench.ncu_traffic("g2k1_bwd_kernel", 4194304, 4096)
    assert t is not None and t["kernel"].startswith("g2k1_bwd_kernel<4, 256, 2, 256, 1>")
    assert 1.3e8 < t["bytes"] < 1.6e8                      # algorithmic: 134 MB of point records + partials
    assert bench.ncu_traffic("g2k1_bwd_kernel", 1048576, 4096) is None      # other workloads: no number
    r = bench.kernel_report("gaussian", "bwd_structured", 4194304.0 * 4096, 9.4074, 1965.0)
    assert abs(r["frac"] - 0.6132) < 1e-3 and abs(r["fp32_slot_frac"] - 0.7849) < 1e-3
    assert abs(r["mufu_frac"] - 0.3925) < 1e-3 and r["binding_pipe"] == "fp32"
    s = bench.kernel_report("softplus", "fwd", 4194304.0 * 4096, 21.0817, 1965.0)
    assert s["binding_pipe"] == "mufu" and abs(s["binding_pipe_frac"] - 0.8757) < 1e-3
