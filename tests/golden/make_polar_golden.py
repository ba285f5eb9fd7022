#!/usr/bin/env python3
"""Generate golden vectors for the Polar(1024,448)+CRC-8 path by RUNNING the
reference (`/root/reference/rtwm/fastpolar.py`, `rtwm/polar_fast.py`) in this
container.  The reference cannot travel to the GPU box, so the vectors it
produces are committed under tests/golden/ together with this script.

    PYTHONPATH=/root/reference python tests/golden/make_polar_golden.py

Inputs are regenerated from seeds by `tests/_inputs.py` (shared with the tests);
only reference OUTPUTS are stored.
"""
from __future__ import annotations
import io, os, sys, contextlib, time
import numpy as np
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, "/root/reference")
from _inputs import awgn_llr_set, detector_like_llr_set  # noqa: E402

from rtwm.fastpolar import PolarCode          # reference
from rtwm import polar_fast as ref_polar_fast  # reference


class _TapAllPaths(PolarCode):
    """Reference decoder with the CRC test forced True: together with a
    recording always-False validator this makes `decode` hand us the hard
    decision and then EVERY final list path in ascending-metric order
    (rtwm/fastpolar.py:335-349)."""
    def _crc_ok(self, info, crc_bits):  # noqa: D401
        return True


def _one(args):
    llr, L = args
    pc = PolarCode(1024, 448, list_size=L, crc_size=8)
    bits, ok = pc.decode(llr)                       # no validator
    seen = []
    def rec(b):
        seen.append(b); return False
    bits_f, ok_f = pc.decode(llr, validator=rec)    # always-False validator
    tap = _TapAllPaths(1024, 448, list_size=L, crc_size=8)
    allp = []
    def rec2(b):
        allp.append(np.frombuffer(b, dtype=np.uint8).copy()); return False
    tap.decode(llr, validator=rec2)
    # allp[0] = hard decision, allp[1:] = final paths sorted by metric
    paths = np.zeros((1 + L, 55), np.uint8)
    for i, p in enumerate(allp[: 1 + L]):
        paths[i] = p
    return (np.packbits(bits), ok, np.packbits(bits_f), ok_f, len(seen), paths, len(allp))


def run_set(name, llrs, L, pool):
    t0 = time.time()
    res = pool.map(_one, [(l, L) for l in llrs], chunksize=1)
    out = dict(
        bits=np.stack([r[0] for r in res]), ok=np.array([r[1] for r in res]),
        bits_falseval=np.stack([r[2] for r in res]), ok_falseval=np.array([r[3] for r in res]),
        n_crc_cands=np.array([r[4] for r in res], np.int32),
        paths=np.stack([r[5] for r in res]), n_paths=np.array([r[6] for r in res], np.int32),
    )
    print(f"{name}: {len(llrs)} codewords L={L} in {time.time()-t0:.0f}s  ok={out['ok'].mean():.3f}", flush=True)
    return out


def main():
    n_awgn = int(os.environ.get("N_AWGN", 256))
    n_tie = int(os.environ.get("N_TIE", 48))
    gold = {}
    with Pool(int(os.environ.get("NPROC", 8))) as pool:
        llr, info = awgn_llr_set(n_awgn, seed=7)
        for k, v in run_set("awgn_L8", llr, 8, pool).items():
            gold[f"awgn8_{k}"] = v
        gold["awgn8_info"] = np.packbits(info, axis=1)
        # smaller list sizes on a subset
        for L in (1, 2, 4):
            for k, v in run_set(f"awgn_L{L}", llr[:32], L, pool).items():
                gold[f"awgn{L}_{k}"] = v
        tl = detector_like_llr_set(n_tie, seed=11)
        for k, v in run_set("tie_L8", tl, 8, pool).items():
            gold[f"tie8_{k}"] = v
    # encoder vectors (rtwm/polar_fast.py:26-53)
    rng = np.random.default_rng(3)
    pay = rng.integers(0, 256, (32, 55), dtype=np.uint8)
    pay[0] = 0; pay[1] = 255
    with contextlib.redirect_stdout(io.StringIO()):
        cw = np.stack([ref_polar_fast.encode(p.tobytes()) for p in pay])
    gold["enc_payload"] = pay
    gold["enc_codeword"] = np.packbits(cw, axis=1)
    pc = PolarCode(1024, 448)
    gold["frozen"] = np.packbits(pc.frozen.astype(np.uint8))
    np.savez_compressed(os.path.join(HERE, "polar_golden.npz"), **gold)
    print("wrote polar_golden.npz")


if __name__ == "__main__":
    main()
