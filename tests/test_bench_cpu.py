"""bench.py's reference arm runs without a GPU: its contract (ONE JSON line on stdout, the keys the driver reads, all host
cores whatever the launcher's OMP_NUM_THREADS says, rank != 0 silent) is checked here on a small --dr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--dr", "1.5e-2", "--steps", "2",
                        "--warmup", "1", *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run({"OMP_NUM_THREADS": "1"})      # what torch.distributed.run exports: must not pin the CPU arm to one core
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["dtype"] == "f64" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["unit"] == line["unit"]
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and abs(line["value"] - line["config"]["particles_timed"] * 2 / (2 * line["ms_per_step"] * 1e-3)) \
        <= 1e-6 * line["value"]


def test_reference_arm_other_ranks_exit_silently():
    out = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"},
               args=("--gpus", "2"))
    assert out.strip() == ""
