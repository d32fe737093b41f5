"""A/B a kernel variant on ONE box (box-to-box spread in the pool is 5-10 %, larger than most kernel changes):
    python scripts/ab_bench.py AAE_TC_MCAST=1 [--workload train] [--rounds 3]
runs bench.py alternately without / with the given environment settings and prints value, ms_per_step and the encoder stage times."""
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
env_set = dict(a.split("=", 1) for a in args if "=" in a and not a.startswith("--"))
rounds = int(args[args.index("--rounds") + 1]) if "--rounds" in args else 3
workload = args[args.index("--workload") + 1] if "--workload" in args else "infer"
cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--no-parity", "--workload", workload]
res = {"base": [], "variant": []}
for r in range(rounds):
    for name in ("base", "variant"):
        env = dict(os.environ)
        if name == "variant":
            env.update(env_set)
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(name, "FAILED:", out.stderr[-400:])
            continue
        d = json.loads(line[-1])
        res[name].append(d)
        print("%-8s round %d: value %.1f  ms/step %.3f  stages %s" % (name, r, d["value"], d["ms_per_step"],
                                                                      [round(x, 3) for x in d.get("roofline", {}).get("stage_ms", [])]))
for name, ds in res.items():
    if ds:
        print("%-8s median value %.1f  median ms/step %.3f" % (name, statistics.median(d["value"] for d in ds), statistics.median(d["ms_per_step"] for d in ds)))
