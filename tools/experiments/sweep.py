"""Run bench.py under several schedule settings (environment variables read by librbn_b200) and print one summary
line per setting.  Usage: python scratch/sweep.py "RBN_TC_OVERLAP=0" "RBN_TC_FWD_SMS0=40 RBN_TC_BWD_SMS0=74" ..."""
import json
import os
import subprocess
import sys

extra = os.environ.get("SWEEP_ARGS", "--steps 3 --warmup 2 --skip-e2e --skip-cpu").split()
for spec in sys.argv[1:]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=")
        env[k] = v
    p = subprocess.run([sys.executable, "bench.py"] + extra, env=env, capture_output=True, text=True)
    line = None
    for ln in p.stdout.splitlines():
        if ln.startswith("{"):
            line = json.loads(ln)
    if line is None:
        print(f"[{spec}] FAILED rc={p.returncode}\n{p.stdout[-1500:]}\n{p.stderr[-1500:]}", flush=True)
        continue
    r = line["roofline"]
    print(f"[{spec}] {line['value']:.1f} seq/s {line['ms_per_step']:.1f} ms frac={r['step_frac']:.3f} "
          f"clk={line['clocks']['sm_mhz']} logZ={line['logZ_mean']:.6f} {r['kernels_ms_per_step']}", flush=True)
