"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-seq", "2"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    r = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in r, k
    assert r["impl"] == "reference" and r["metric"] == "seconds_of_audio_matched_per_second"
    assert r["unit"] == "s_audio/s" and r["higher_is_better"] is True and r["vs_baseline"] is None
    assert r["config"]["workload"] == "speaker10_24s" and r["config"]["windows"] == 13312
    assert r["e2e"] == {"value": r["value"], "unit": "s_audio/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = r["cpu_baseline"]
    # "reference" when the reference tree is present (this container), "port" on the GPU box
    assert cb["kind"] in ("port", "reference") and cb["cores"] == 1 and cb["value"] == r["value"] and "sample" in cb
    assert r["extrapolated"] is True and r["scale"] == 256.0 and cb["extrapolated"] is True
    assert 0 < r["value"] < 10


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
