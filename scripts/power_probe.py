"""SM clock / power under a sustained loop of one search shape (cached vs uncached norms).

    python scripts/power_probe.py Q [seconds]
"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pynvml
import torch

import imageretrievalresearch_b200 as irr

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 64
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
N, D = 1_000_000, 1536
g = torch.randn(N, D, device="cuda", dtype=torch.bfloat16)
q = torch.randn(Q, D, device="cuda", dtype=torch.bfloat16)
gal = irr.Gallery(g)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def sample(stop, out):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                    pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                    pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                    pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.02)


for mode, fn in (("cached", lambda: gal.search(q, 3)), ("uncached", lambda: irr.cosine_topk(q, g, 3)),
                 ("cached", lambda: gal.search(q, 3)), ("uncached", lambda: irr.cosine_topk(q, g, 3))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    stop, out = threading.Event(), []
    t = threading.Thread(target=sample, args=(stop, out))
    t.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    s.record()
    while time.time() - t0 < secs:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    stop.set()
    t.join()
    out = out[len(out) // 4:]
    sm = sorted(o[0] for o in out)[len(out) // 2]
    mem = sorted(o[1] for o in out)[len(out) // 2]
    pw = sorted(o[2] for o in out)[len(out) // 2]
    reasons = 0
    for o in out:
        reasons |= o[3]
    print(f"Q={Q} {mode}: {s.elapsed_time(e) / n:.4f} ms/search, SM {sm} MHz, MEM {mem} MHz, {pw:.0f} W, "
          f"throttle reasons 0x{reasons:x}", flush=True)
