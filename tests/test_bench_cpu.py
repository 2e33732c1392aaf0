"""bench.py's reference arm (`--impl reference`) on a small grid, on CPU: the line carries the keys the driver reads, its
`config` is the dict our own arm prints (the reference arm runs "on your arm's config"), ranks other than 0 exit 0 without
work.  The arm times the unmodified reference (oracle/_ref) — bench.py's cpu_baseline leg is one of the three places that
may execute anything under oracle/."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_bindings as rb  # noqa: E402

needs_ref = pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (make -C oracle)")


def run_arm(extra_env=None, gpus=1):
    env = dict(os.environ)
    env.update(extra_env or {})
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", str(gpus), "--grid", "32", "--nev", "8",
           "--steps", "3", "--warmup", "1", "--ref-budget", "20"]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)


@needs_ref
def test_reference_arm_line_and_shared_config():
    import bench
    r = run_arm()
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lobpcg_iters_per_s" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["config"] == bench.bench_config(32, 8, 16, 1)          # what run_ours prints for the same arguments
    assert d["extrapolated"] is False and d["reference_run"]["grid_timed"] == 32
    assert d["steps"] >= 3 and d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    r = run_arm({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, gpus=2)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_share_the_config_builder():
    import bench
    c1, c8 = bench.bench_config(160, 150, 300, 1), bench.bench_config(160, 150, 300, 8)
    assert set(c1) == {"workload", "parallelism", "l2"}
    assert "160^3" in c1["workload"] and "nev=150" in c1["workload"] and "sizeSub=300" in c1["workload"]
    assert c1["parallelism"] == "rows in 1 z-slab(s)" and c8["parallelism"] == "rows in 8 z-slab(s)"
    assert "29.5 GB" in c1["l2"] and "3.7 GB" in c8["l2"]
