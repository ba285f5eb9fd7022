#!/usr/bin/env python3
"""Positive-verdict golden vectors from the REFERENCE: identity matched filter seeded
into rx._mf_cache, ideal +-1 frame symbols (+ Gaussian noise), then the reference's own
_try_decode_frame / _llr / _decode_header / verify_raw_frame.  Inputs are rebuilt from seeds by the test.

    python tests/golden/make_positive_golden.py
"""
import contextlib, io, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")
from oracle import tx_oracle as txo
from rtwm.detector import WatermarkDetector
from rtwm.utils import BAND_PLAN, choose_band

KEY = bytes([0x5A]) * 32
CASES = [(0, 0.0, 1), (5, 0.05, 2), (1234, 0.12, 3), (70001, 0.10, 4),      # (ctr, sigma, noise seed): hard-decision hits
         # LIST-PATH positives: the hard decision fails CRC (rtwm/fastpolar.py:261-276) and the accepted payload is a
         # candidate of the list stage (rtwm/fastpolar.py:335-349).  Found by searching 2400 seeded frames with the
         # oracle (2 hits at sigma = 0.12; both sit on exact metric ties, i.e. they are tie-prone as well).
         (1959, 0.12, 1107), (356, 0.12, 1388)]


def frame_for(ctr, sigma, seed):
    k = txo.Keys(KEY)
    payload = txo.build_payload(k, ctr, b"NONCE123", bytes(11), bytes(range(12)))
    rng = np.random.default_rng(seed)
    return txo.frame_symbols(k, ctr, payload).astype(np.float64) + sigma * rng.standard_normal(1215)


def main():
    g = {}
    for ctr, sigma, seed in CASES:
        sym = frame_for(ctr, sigma, seed)
        rx = WatermarkDetector(KEY, list_size=8)
        for lo, hi in BAND_PLAN:
            rx._mf_cache[(lo, hi, 48000)] = np.array([1.0], np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            ok = rx._try_decode_frame(sym, ctr)
            nonce = rx.session_nonce
            again = rx._try_decode_frame(sym, ctr)
            wrong = rx._try_decode_frame(sym, ctr + 1)
            rx.session_nonce = b"OTHERNON"
            mism = rx._try_decode_frame(sym, ctr)
            l0 = rx._llr(sym, ctr, 0); l1 = rx._llr(sym, ctr, 1)
            hdr = rx._decode_header(sym, choose_band(KEY, ctr))
        # was the accepted payload the hard decision or a list candidate?  (same arithmetic as rtwm/fastpolar.py:261-276)
        from rtwm.fastpolar import PolarCode
        pc = PolarCode(1024, 448, list_size=8, crc_size=8)
        hard_ok = []
        for v in (l0, -l0, l1, -l1):
            u = pc._polar_transform((v.astype(np.float64) > 0.0).astype(np.uint8)); u[pc.frozen] = 0
            d = u[pc._data_pos]
            hard_ok.append(bool(pc._crc_ok(d[:pc._info_len], d[pc._info_len:pc.K])))
        pre = f"c{ctr}/"
        g[pre + "hard_crc"] = np.array(hard_ok)
        g[pre + "verdicts"] = np.array([ok, again, wrong, mism])
        g[pre + "nonce"] = np.frombuffer(nonce or b"", np.uint8)
        g[pre + "llr0"] = l0; g[pre + "llr1"] = l1
        g[pre + "hdr"] = np.array([float(hdr[0]), float(hdr[1]), float(hdr[2])])
        print(ctr, sigma, ok, again, wrong, mism, hdr, "hard CRC of (llr0,-llr0,llr1,-llr1):", hard_ok, flush=True)
    np.savez_compressed(os.path.join(HERE, "positive_golden.npz"), **g)


if __name__ == "__main__":
    main()
