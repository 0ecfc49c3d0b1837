"""bench.py's output contract (CPU): exactly ONE JSON line on stdout even when a library (NCCL prints its version banner)
writes to file descriptor 1 during the run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_json_line_on_stdout():
    code = (
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "sys.stdout.flush()\n"
        "bench._REAL_STDOUT = os.dup(1)\n"
        "os.dup2(2, 1)\n"
        "os.write(1, b'NCCL version 0.0.0 (library chatter on fd 1)\\n')\n"
        "print('python-level chatter')\n"
        "bench._emit({'metric': 'flows/sec DDIM-50 @436x1024', 'value': 1.5})\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout
    assert json.loads(lines[0]) == {"metric": "flows/sec DDIM-50 @436x1024", "value": 1.5}
    assert "NCCL version" in r.stderr and "python-level chatter" in r.stderr


def test_reduction_rate_profile_is_parsed():
    """bench.py's `frac_of_l2_reduction_rate` divides by the rate measured by scripts/micro/red_rate.cu for the white-noise
    scatter pattern (profiles/r2_red_rate.txt): the committed profile must parse to a plausible number of reductions / s."""
    sys.path.insert(0, ROOT)
    import bench
    rate = bench.load_red_rate()
    assert rate is not None and 1e11 < rate < 1e12, rate
