"""Quick parity probe of the SCL list kernel against the device-arithmetic model (oracle with the kernel's phi):
bit-for-bit paths, CRC flags, metrics.  Developer tool for A/B runs of kernel variants (ES_B200_LIB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from echoseal_b200 import polar_gpu
from oracle import polar_oracle as po
from _inputs import awgn_llr_set, detector_like_llr_set

def cmp(out, model, tag):
    ok = (out["payload"].cpu().numpy() == np.packbits(model["path_info"], axis=2)).all() and \
         (out["crc"].cpu().numpy() == model["path_crc"]).all() and \
         (out["metric"].cpu().numpy() == model["path_metric"]).all() and \
         (out["npaths"].cpu().numpy() == model["npaths"]).all()
    print(("PASS " if ok else "FAIL ") + tag, flush=True)
    return ok

def main():
    good = True
    llr, _ = awgn_llr_set(192, seed=99)
    d = torch.from_numpy(llr).cuda()
    for L in (8, 4, 1):
        good &= cmp(polar_gpu.list_decode(d, list_size=L), po.scl_batch(llr, L=L, device_arith=True), f"awgn L={L}")
    tl = detector_like_llr_set(130, seed=5)
    dt = torch.from_numpy(tl).cuda()
    for L in (8, 3):
        good &= cmp(polar_gpu.list_decode(dt, list_size=L, neg_mode=1),
                    po.scl_batch(tl, L=L, device_arith=True, neg_mode=True), f"pair detector-like L={L}")
    # a batch large enough that every warp of the persistent grid decodes several groups (slot/ring state carried over)
    big = detector_like_llr_set(12000, seed=6)
    db = torch.from_numpy(big).cuda()
    out = polar_gpu.list_decode(db, list_size=8, neg_mode=1)
    sel = np.r_[0:64, 9400:9500, 11936:12000]
    model = po.scl_batch(big[sel], L=8, device_arith=True, neg_mode=True)
    sel2 = np.stack([2 * sel, 2 * sel + 1], 1).reshape(-1)
    sub = {k: v[torch.from_numpy(sel2).cuda()] for k, v in out.items()}
    good &= cmp(sub, model, "pair 12000 rows (sampled 228)")
    print("ALL PASS" if good else "SOME FAILED")
    sys.exit(0 if good else 1)

if __name__ == "__main__":
    main()
