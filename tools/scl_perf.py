"""Quick SCL-8 throughput probe (CUDA events, L2-exceeding input set)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from echoseal_b200 import polar_gpu

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    g = torch.Generator(device="cuda").manual_seed(0)
    llr = torch.clamp(torch.randn((n, 1024), device="cuda", generator=g) * 2.0, -12, 12).contiguous()
    from echoseal_b200 import _native as N
    print("ctas/SM", N.lib().es_scl_ctas_per_sm(), "grid", N.lib().es_scl_grid_ctas(),
          "scratch MB", N.lib().es_scl_scratch_bytes() / 1e6)
    polar_gpu.list_decode(llr[:4096], list_size=L)
    torch.cuda.synchronize()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = polar_gpu.list_decode(llr, list_size=L)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"list_decode n={n} L={L}: {ms:.2f} ms  -> {n / ms * 1e3 / 1e6:.3f} M cw/s")
    # the detector's pairing: n/2 rows -> n codewords (+row, -row), first half of the tree shared
    half = llr[: n // 2]
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = polar_gpu.list_decode(half, list_size=L, neg_mode=1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"pair_decode n={n} L={L}: {ms:.2f} ms  -> {n / ms * 1e3 / 1e6:.3f} M cw/s")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); polar_gpu.hard_decide(llr); e1.record(); torch.cuda.synchronize()
    print(f"hard_decide n={n}: {e0.elapsed_time(e1):.3f} ms")

if __name__ == "__main__":
    main()
