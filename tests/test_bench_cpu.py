"""CPU: bench.py's named workloads are BASELINE.json's configurations, and the reference arm runs (tiny shape) with every
host thread whatever OMP_NUM_THREADS says (under torchrun it ran on one thread)."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_named_configs_match_baseline_json():
    sys.path.insert(0, ROOT)
    import bench

    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    c = bench.CONFIGS
    assert sorted(c) == ["c2", "c3", "c4", "c5"]
    assert (c["c2"]["users"], c["c2"]["items"], c["c2"]["dim"], c["c2"]["k"], c["c2"]["distance"]) == (10**6, 10**6, 128, 10, "dot")
    assert "n_factors=128" in base[1] and "K=10" in base[1]
    assert (c["c3"]["distance"], c["c3"]["k"]) == ("cosine", 100) and "COSINE" in base[2] and "K=100" in base[2]
    assert (c["c4"]["items"], c["c4"]["k"]) == (10**7, 10) and "10M items" in base[3]
    assert (c["c5"]["items"], c["c5"]["dim"], c["c5"]["k"], c["c5"]["tc"]) == (5 * 10**6, 256, 20, "bf16") and "n_factors=256" in base[4]

    class A:
        config, users, items, dim, k, viewed, distance, tc = "c3", None, 1234, None, None, None, None, None

    a = A()
    bench.resolve_config(a)
    assert (a.users, a.items, a.k, a.distance) == (10**6, 1234, 100, "cosine")  # explicit arguments win
    assert bench.workload_name(a).startswith("config3:") and "Distance.COSINE" in bench.workload_name(a)


def test_reference_arm_uses_all_threads_under_torchrun_env():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
         "--users", "512", "--items", "4000", "--ref-users", "128"],
        capture_output=True, text=True, env=env, timeout=300, cwd=ROOT,
    )
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["e2e"]["h2d_bytes_per_step"] == 0
    n = len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["cores"] == n and line["cpu_baseline"]["kind"] == "port"
    # the other ranks of a torchrun launch print nothing and exit 0
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         env=env, timeout=120, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
