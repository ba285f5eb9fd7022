"""Stress probe: the detector pairing on a large batch, kernel vs device-arithmetic model on EVERY codeword, and
run-to-run determinism.  Developer tool (ES_B200_LIB selects the variant)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from echoseal_b200 import polar_gpu
from oracle import polar_oracle as po
from _inputs import detector_like_llr_set

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    big = detector_like_llr_set(n, seed=6)
    db = torch.from_numpy(big).cuda()
    model = po.scl_batch(big, L=8, device_arith=True, neg_mode=True)
    mp = np.packbits(model["path_info"], axis=2)
    first = None
    for r in range(reps):
        out = polar_gpu.list_decode(db, list_size=8, neg_mode=1)
        pay = out["payload"].cpu().numpy(); met = out["metric"].cpu().numpy(); crc = out["crc"].cpu().numpy()
        bad_p = np.flatnonzero((pay != mp).any(axis=(1, 2)))
        bad_m = np.flatnonzero((met != model["path_metric"]).any(axis=1))
        print(f"rep {r}: payload mismatches {bad_p.size}, metric mismatches {bad_m.size} of {2 * n} codewords", flush=True)
        if bad_m.size:
            w = bad_m[:12]
            print("  first bad codewords:", w.tolist(), "groups", (w // 8).tolist())
            print("  max rel metric diff", float(np.max(np.abs(met[bad_m] - model["path_metric"][bad_m]) / np.maximum(model["path_metric"][bad_m], 1e-300))))
        if first is None:
            first = (pay, met)
        else:
            print("  identical to rep 0:", bool((pay == first[0]).all() and (met == first[1]).all()))

if __name__ == "__main__":
    main()
