"""bench.py's driver contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the agreed keys
(it times the reference's own shader text compiled for the host, or the oracle port if no object for the option set exists), ranks
other than 0 stay silent, and our arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json, os, subprocess, sys
import pytest
from conftest import ROOT


def run(args, env=None):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_json_line():
    r = run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "cornell_box_orig"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpath-segments/s" and d["unit"] == "Mpath-segments/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    from oracle.glsl_ref import binding as gb
    if gb.reference_available():
        assert cb["kind"] == "reference"          # the reference's own shader text is what gets timed where it can be built


def test_reference_arm_other_ranks_are_silent():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "cornell_box_orig"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(["--steps", "1", "--warmup", "1", "--workload", "cornell_box_orig"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_roofline_figures_come_from_the_latest_committed_capture():
    """bench.py cannot count instructions or DRAM bytes without a profiler: it quotes them from the ncu capture committed under profiles/ (labelled with the
    capture's commit and file).  The capture it picks must be the latest round's, must belong to the headline workload and must carry the figures the line uses."""
    import glob, importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    cap = b.capture("hyperion_rect_lights")
    latest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_capture.json")))[-1]
    assert cap is not None and os.path.join(ROOT, cap["file"]) == latest
    for k in ("commit", "k_trace_dram_bytes_per_step", "k_trace_thread_inst_per_step", "k_trace_lane_issue_frac", "k_shadow_lane_issue_frac", "k_trace_l2_gbs"):
        assert cap.get(k), k
    assert 0.0 < cap["k_trace_lane_issue_frac"] < 1.0 and 0.0 < cap["k_shadow_lane_issue_frac"] < 1.0
    assert b.capture("cornell_box_orig") is None            # no capture of that workload: nothing is quoted
