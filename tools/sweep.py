#!/usr/bin/env python3
"""Dev helper: run bench.py under several RF_W1_TUNE settings and print compact results."""
import json, os, subprocess, sys
tunes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"]
extra = sys.argv[2:]
for t in tunes:
    env = dict(os.environ, RF_W1_TUNE=t, RF_W1_PATH=("1" if t != "lb" else "0"))
    out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--e2e-steps", "1"] + extra,
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print("tune", t, "pairs/s %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"],
              "ok", d["config"]["results_match_oracle_sample"], "clk", d["clocks"]["sm_mhz"], flush=True)
    except Exception as e:
        print("tune", t, "FAILED", e, out.stdout[-500:], out.stderr[-1500:], flush=True)
