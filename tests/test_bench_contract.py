"""bench.py's reference arm (the CPU leg the driver runs beside the GPU arm): one JSON line with the contract's keys;
under torchrun only rank 0 runs it — with all host threads although torchrun exports OMP_NUM_THREADS=1. Needs no GPU; the
workload is shrunk to 2^12 lookups through B200_BENCH_MU so that the CPU suite stays short."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ)
    env["B200_BENCH_MU"] = "12"
    env["OMP_NUM_THREADS"] = "1"  # what torchrun sets for its workers
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                           "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900, env=env)


def test_reference_arm_prints_one_contract_line():
    out = run()
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["value"] > 0 and d["ms_per_step"] == d["value"] and d["vs_baseline"] is None
    assert "2^12" in d["metric"] and list(d["config"]) == ["workload"] and d["scaling"] == "strong"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))  # not the 1 thread OMP_NUM_THREADS asked for
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["parity"]["proof_bytes"] == 54112 and len(d["parity"]["sha256"]) == 64


def test_reference_arm_other_ranks_exit_quietly():
    out = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""
