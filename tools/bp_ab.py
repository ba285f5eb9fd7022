"""K1 A/B: TMA tensor-store form vs plain form (same chunk grid), CUDA-event times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from echoseal_b200 import rx_gpu
from echoseal_b200.utils import BAND_PLAN

def main():
    taps = [rx_gpu.matched_filter_taps(b, 48000) for b in BAND_PLAN]
    rx_gpu.set_filters(48000, taps)
    for B, n in [(2000, 144000), (1000, 144000)]:
        x = (torch.randn((B, n), device="cuda") * 0.1).contiguous()
        for plain in (0, 2, 1):
            from echoseal_b200 import _native as N
            import ctypes as C
            N.lib().es_rx_bandpass_force_plain(C.c_int(plain))
            y = rx_gpu.bandpass(x); del y
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); y = rx_gpu.bandpass(x); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1)); del y
            print(f"B={B} n={n} {('tma in+out', 'plain', 'tma out only')[plain]}: {min(ts):.3f} ms  ({B * n * 32 / min(ts) / 1e6:.0f} GB/s written)")
        rx_gpu.bandpass_force_plain(False)

if __name__ == "__main__":
    main()
