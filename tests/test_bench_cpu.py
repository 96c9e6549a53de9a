"""bench.py contract checks that need no GPU: the reference arm (CPU port of the path) prints one
JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--rows", "20000", "--batch", "8"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["value"] > 0 and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    assert "workload" in j["config"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                          "--rows", "1000"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_watchdog_prints_the_line_collected_so_far_and_ends_the_process():
    """If a secondary measurement of bench.py stalls, the headline must still come out: the watchdog dumps the stacks to
    stderr, prints the JSON collected so far with `secondary_incomplete` and exits 0 (the output is redirected to a file
    here, like the driver does: the line must not stay in a buffer)."""
    code = ("import sys, time; sys.path.insert(0, %r); import bench\n"
            "out = {'metric': 'm', 'value': 1.5}\n"
            "bench._arm_watchdog(out, 0, 0.3)\n"
            "out['sweep'] = [1, 2]\n"
            "time.sleep(30)\n") % ROOT
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        so, se = os.path.join(td, "o"), os.path.join(td, "e")
        with open(so, "w") as fo, open(se, "w") as fe:
            rc = subprocess.run([sys.executable, "-c", code], stdout=fo, stderr=fe, timeout=60, cwd=ROOT).returncode
        assert rc == 0
        j = json.loads(open(so).read().strip())
        assert j["value"] == 1.5 and j["sweep"] == [1, 2] and "secondary_incomplete" in j
        assert "time.sleep" in open(se).read() or "File" in open(se).read()      # the stack of the stalled main thread
