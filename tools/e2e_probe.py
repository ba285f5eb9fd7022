"""Step-time probe: device-resident vs pinned-host input of the 10k-clip RX batch, three steps each, and raw H2D bandwidth."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from echoseal_b200 import detector, rx_gpu
from echoseal_b200.utils import BAND_PLAN
dev = torch.device("cuda", 0)
keys, _, clips = bench.make_clips_gpu(0, 10000, dev)
taps = [rx_gpu.matched_filter_taps(b, 48000) for b in BAND_PLAN]
def step(a):
    return detector.verify_batch(keys, a, list_size=8, mf_taps=taps, sub_batch=1000, host_threads=16)
host = torch.empty((10000, 144000), dtype=torch.float32, pin_memory=True); host.copy_(clips)
print("slice pinned:", host[1000:2000].is_pinned(), host[1000:2000].is_contiguous())
for name, a in (("dev", clips), ("host", host), ("dev", clips), ("host", host)):
    step(a); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); step(a); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(name, [round(t * 1e3, 1) for t in ts])
# raw H2D bandwidth from the pinned buffer
x = torch.empty((1000, 144000), dtype=torch.float32, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(10): x.copy_(host[k * 1000:(k + 1) * 1000], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D GB/s", 5.76 / dt)
