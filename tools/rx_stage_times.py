"""Wall-clock breakdown of one verify_batch sub-batch (host + device, synchronised per stage)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from echoseal_b200 import rx_gpu, polar_gpu, detector
from echoseal_b200.host_feeder import KeyBank
from echoseal_b200.utils import BAND_PLAN

def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    dev = torch.device("cuda", 0)
    keys, bank, clips = bench.make_clips_gpu(0, nb, dev)
    taps = [rx_gpu.matched_filter_taps(b, 48000) for b in BAND_PLAN]
    rx_gpu.set_filters(48000, taps)
    kidx = np.arange(nb, dtype=np.int32)
    def T(label, fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        print(f"{label:28s} {1e3 * (time.perf_counter() - t0):8.2f} ms"); return r
    from echoseal_b200 import _native as N
    for rep in range(2):
        print("--- rep", rep)
        N.KERNEL_TIMES = {}
        bank = T("KeyBank(keys)", lambda: KeyBank(keys))
        hdr_pn = T("hdr_pn->dev", lambda: torch.from_numpy(bank.hdr_pn(kidx)).to(dev))
        y = T("bandpass", lambda: rx_gpu.bandpass(clips))
        corr, aux = T("ncc (+K3 first pass)", lambda: rx_gpu.ncc(y, with_hist=True))
        pk, npk, st = T("peaks", lambda: rx_gpu.peaks(corr, aux))
        fr = T("frames", lambda: rx_gpu.frames(y, pk, npk, hdr_pn))
        del y
        corr_f = T("scan (K1+K2 fused)", lambda: rx_gpu.scan(clips))
        print("   fused vs staged corr max diff", float((corr_f - corr).abs().max()))
        del corr_f
        fr2 = T("frames_x", lambda: rx_gpu.frames_x(clips, pk, npk, hdr_pn))
        del fr2
        pk_h, npk_h, hdr_h = T("d2h peaks/hdr", lambda: (pk.cpu().numpy(), npk.cpu().numpy(), fr["hdr"].cpu().numpy()))
        enum = T("rx_enumerate", lambda: bank.rx_enumerate(kidx, clips.shape[1], pk_h, npk_h, hdr_h))
        I = enum["item_peak"].size
        ip = T("h2d items/pn", lambda: (torch.from_numpy(enum["item_peak"]).to(dev), torch.from_numpy(enum["pn"]).to(dev)))
        llr = T("llr", lambda: rx_gpu.llr(fr["mf_aligned"], ip[0], ip[1]))
        hd = T("scl_hard", lambda: polar_gpu.hard_decide(llr, neg_mode=1))
        out = T(f"scl_list ({4*I} cw)", lambda: polar_gpu.list_decode(llr, list_size=8, neg_mode=1))
        def collect():
            flags = torch.cat([hd[1][:, None], out["crc"]], dim=1)
            idx = torch.nonzero(flags, as_tuple=False)
            allpay = torch.cat([hd[0][:, None, :], out["payload"]], dim=1)
            hp = allpay[idx[:, 0], idx[:, 1]].cpu().numpy()
            ih = idx.cpu().numpy()
            return ih, hp
        ih, hp = T("collect hits", collect)
        ns = np.zeros((nb, 9), np.uint8)
        T(f"rx_validate ({len(ih)} hits)", lambda: bank.rx_validate(kidx, enum, ih[:, 0].astype(np.int64), ih[:, 1].astype(np.int32), hp, ns))
        torch.cuda.synchronize()
        print("kernel-only (CUDA events):", {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in N.KERNEL_TIMES.items()})
        N.KERNEL_TIMES = None

if __name__ == "__main__":
    main()
