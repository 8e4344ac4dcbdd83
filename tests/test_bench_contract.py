"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints exactly one
JSON line on stdout with the agreed keys, and the own arm refuses to run without a CUDA device (no
CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    p = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-budget", "2")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "dqn_transitions_trained_per_sec"
    assert j["unit"] == "transitions/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["steps"] == 2 and j["warmup"] == 1 and j["n_gpus"] == 1 and j["ms_per_step"] > 0
    assert j["vs_baseline"] is None and j["data"] == "synthetic" and "workload" in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_own_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return          # on a GPU box the own arm is exercised by the driver
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0
    assert not [l for l in p.stdout.splitlines() if l.strip().startswith("{")]
    assert "CUDA" in (p.stderr + p.stdout) or "GPU" in (p.stderr + p.stdout)


def test_reference_arm_under_torchrun_env_uses_all_cores_and_the_stated_global_batch():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use the box's cores (round 1's
    N>1 ratios were void because it did not), rank 0 alone reports, and N>1 is configs[3] at its
    stated global batch of 65536 (strong scaling)."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "1", "--cpu-budget", "2"], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    j = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
    assert j["cpu_baseline"]["cores"] == min(len(os.sched_getaffinity(0)), 64)
    assert j["scaling"] == "strong" and j["config"]["global_batch"] == 65536 and j["config"]["batch_per_gpu"] == 32768
    env["RANK"] = "1"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and not p.stdout.strip()
