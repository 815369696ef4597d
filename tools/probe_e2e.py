"""Scratch: where the e2e step's time goes (host clock around each C-ABI call)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
import torch
from blokus_self_play import GameBatch
n = 4096
b = GameBatch(n)
ids = torch.arange(n, dtype=torch.int32).pin_memory()
plies = torch.empty(n, dtype=torch.int32).pin_memory()
scores = torch.empty((n, 4), dtype=torch.int32).pin_memory()
hist = torch.empty((n, 360), dtype=torch.int16).pin_memory()
P = lambda t: C.c_void_p(t.data_ptr())
acc = [0.0] * 5
K = 50
for k in range(K + 5):
    t0 = time.perf_counter(); b.reset()
    t1 = time.perf_counter(); b.run_playout_raw(k, P(ids))
    t2 = time.perf_counter(); b.fetch_raw(P(plies), P(scores), None)
    t3 = time.perf_counter(); b.fetch_raw(None, None, P(hist))
    t4 = time.perf_counter(); s = int(plies.sum().item())
    t5 = time.perf_counter()
    if k >= 5:
        for i, d in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            acc[i] += d
print("per step us: reset %.1f  playout-enqueue %.1f  fetch plies+scores (incl. kernel wait) %.1f  fetch history %.1f  python sum %.1f" % tuple(1e6 * a / K for a in acc))
print("kernel ms", b.last_kernel_ms())
