"""Probe the INT32 issue peak of this pool's B200 (bk_probe_int_peak: a register-resident LOP3/SHF mix on every SM) with the
clocks seen while it runs; writes gpurun_out/r02_int_peak.json (copied to profiles/, where bench.py reads roofline.peak)."""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
from blokus_self_play import probe_int_peak

lines = []
p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                      "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(l.strip()) for l in p.stdout], daemon=True).start()
vals = []
t0 = time.time()
while time.time() - t0 < 3.0:
    vals.append(probe_int_peak(0))
time.sleep(0.1)
p.terminate()
sm = sorted(float(l.split(",")[0]) for l in lines if l and l.split(",")[0].strip().replace(".", "").isdigit())
vals.sort()
out = {"lane_ops_per_s": vals[len(vals) // 2], "best": vals[-1], "worst": vals[0], "probes": len(vals),
       "nominal": 148 * 64 * 1.965e9, "sm_mhz_median_during_probe": sm[len(sm) // 2] if sm else None,
       "sm_mhz_max_seen": sm[-1] if sm else None, "clock_samples": len(sm), "last_sample": lines[-1] if lines else None,
       "how": "bk_probe_int_peak repeated for 3 s on one B200 (median of the probes); nvidia-smi sampled every 50 ms"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_int_peak.json"), "w"), indent=1)
print(json.dumps(out))
