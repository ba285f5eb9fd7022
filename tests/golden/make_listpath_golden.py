#!/usr/bin/env python3
"""Golden vectors for the LIST-PATH acceptance chain, produced by the REFERENCE (rtwm/detector.py:154-233 on top of
rtwm/fastpolar.py:254-359): LLR rows of a sealed payload over an AWGN channel at which the hard decision fails CRC
(rtwm/fastpolar.py:261-276) and the payload the reference accepts is a candidate of its list stage
(rtwm/fastpolar.py:335-349).  Unlike the identity-channel frames of make_positive_golden.py these rows are generic reals:
the prune decisions of the row that carries the codeword are far from exact metric ties (min relative gap stored), so the answer does not depend on the last
bit of anybody's libm.  The rows are handed to the reference's own _try_decode_frame through its _llr method.

    python tests/golden/make_listpath_golden.py
"""
import contextlib, io, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")
from oracle import tx_oracle as txo
from oracle import polar_oracle as po
from rtwm.detector import WatermarkDetector
from rtwm.polar_fast import encode as ref_encode
from rtwm.fastpolar import PolarCode

KEY = bytes([0x5A]) * 32
SIGMAS = (0.30, 0.33, 0.35)
WANT = 4            # cases kept
# which of the four ladder variants (llr0, -llr0, llr1, -llr1; rtwm/detector.py:168-190) carries the codeword
PLACEMENTS = (0, 1, 2, 3)


def ref_verdicts(rows, ctr):
    rx = WatermarkDetector(KEY, list_size=8)
    rx._llr = lambda frame, frame_ctr, pn_variant=0: rows[1 if pn_variant else 0].copy()
    frame = np.zeros(1215)
    with contextlib.redirect_stdout(io.StringIO()):
        ok = rx._try_decode_frame(frame, ctr)
        nonce = rx.session_nonce
        again = rx._try_decode_frame(frame, ctr)
        wrong = rx._try_decode_frame(frame, ctr + 1)
        rx.session_nonce = b"OTHERNON"
        mism = rx._try_decode_frame(frame, ctr)
    return [ok, again, wrong, mism], nonce


def main():
    pc = PolarCode(1024, 448, list_size=8, crc_size=8)
    k = txo.Keys(KEY)
    g = {}
    kept = 0
    seed = 0
    while kept < WANT and seed < 4000:
        seed += 1
        rng = np.random.default_rng(seed)
        ctr = int(rng.integers(1, 1 << 20))
        sigma = SIGMAS[seed % len(SIGMAS)]
        payload = txo.build_payload(k, ctr, b"NONCE123", bytes(11), bytes(rng.integers(0, 256, 12, dtype=np.uint8)))
        cw = ref_encode(payload).astype(np.float64)
        y = (2.0 * cw - 1.0) + sigma * rng.standard_normal(1024)
        row = np.clip(2.0 * y / sigma ** 2, -12.0, 12.0).astype(np.float32)      # LLR = log P1/P0
        noise_row = np.clip(rng.standard_normal(1024) * 2.0, -12.0, 12.0).astype(np.float32)
        place = PLACEMENTS[kept % 4]
        rows = np.stack([noise_row, noise_row])
        rows[place >> 1] = -row if (place & 1) else row
        # pre-filter with the oracle: every hard decision fails, the list holds the payload, no near-tie anywhere
        four = np.stack([rows[0], -rows[0], rows[1], -rows[1]])
        res = po.scl_batch(four, L=8)
        if res["hard_crc"].any():
            continue
        truth = np.unpackbits(np.frombuffer(payload, np.uint8))
        w = place
        hit = [a for a in range(int(res["npaths"][w])) if res["path_crc"][w, a] and (res["path_info"][w, a] == truth).all()]
        if not hit or float(res["stats"][w, 1]) < 1e-9:
            continue
        # the reference's own answer
        hard_ok = []
        for v in four:
            u = pc._polar_transform((v.astype(np.float64) > 0.0).astype(np.uint8)); u[pc.frozen] = 0
            d = u[pc._data_pos]
            hard_ok.append(bool(pc._crc_ok(d[:pc._info_len], d[pc._info_len:pc.K])))
        verdicts, nonce = ref_verdicts(rows, ctr)
        print(f"seed {seed} ctr {ctr} sigma {sigma} place {place} list rank {hit[0]} min rel gap {float(res["stats"][w, 1]):.2e} "
              f"hard {hard_ok} reference verdicts {verdicts}", flush=True)
        if any(hard_ok) or not verdicts[0]:
            continue
        pre = f"k{kept}/"
        g[pre + "ctr"] = np.array([ctr]); g[pre + "rows"] = rows; g[pre + "place"] = np.array([place])
        g[pre + "payload"] = np.frombuffer(payload, np.uint8)
        g[pre + "verdicts"] = np.array(verdicts); g[pre + "nonce"] = np.frombuffer(nonce or b"", np.uint8)
        g[pre + "min_rel_gap"] = np.array([float(res["stats"][w, 1])]); g[pre + "list_rank"] = np.array([hit[0]])
        kept += 1
    g["n"] = np.array([kept])
    np.savez_compressed(os.path.join(HERE, "listpath_golden.npz"), **g)
    print("kept", kept)


if __name__ == "__main__":
    main()
