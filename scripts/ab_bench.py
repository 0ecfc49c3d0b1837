"""Same-box A/B of an environment switch on the headline bench: python scripts/ab_bench.py VAR a b [reps]
Runs `bench.py --steps 3 --warmup 3 --no-gpu-baseline --skip-warp` alternately with VAR=a and VAR=b and prints flows/s."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
var, vals, reps = sys.argv[1], sys.argv[2:4], int(sys.argv[4]) if len(sys.argv) > 4 else 2
extra = os.environ.get("AB_ARGS", "--steps 3 --warmup 3 --no-gpu-baseline --skip-warp --skip-train").split()
for _ in range(reps):
    for v in vals:
        env = dict(os.environ, **{var: v})
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *extra], env=env, capture_output=True, text=True)
        line = [x for x in out.stdout.splitlines() if x.startswith("{")]
        if not line:
            print(var, v, "FAILED", out.stderr[-400:])
            continue
        d = json.loads(line[-1])
        print(f"{var}={v}: {d['value']:.3f} flows/s  e2e {d['e2e']['value']:.3f}  sm {d['clocks']['sm_mhz']} MHz  conv frac {d['roofline']['frac']:.3f}",
              flush=True)
