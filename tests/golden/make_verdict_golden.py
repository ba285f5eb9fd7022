#!/usr/bin/env python3
"""Run the UNMODIFIED reference verify() (list_size=8, real fastpolar decoder) on a few golden clips and
store verdict + timing.  ~1 s per SCL decode, so only clips with few attempts are used.

    python tests/golden/make_verdict_golden.py
"""
import contextlib, io, os, sys, time, json
import numpy as np
from multiprocessing import Pool
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")
NAMES = ["noise_44", "bench_19", "bench_17", "short_1s", "plain_noise", "chirp_aa"]

def run(name):
    from _inputs import make_clip
    from rtwm.detector import WatermarkDetector
    audio, key = make_clip(name)
    rx = WatermarkDetector(key, list_size=8)
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        v = rx.verify(audio, 48000)
    return name, bool(v), time.time() - t0

if __name__ == "__main__":
    with Pool(len(NAMES)) as p:
        res = p.map(run, NAMES)
    out = {n: {"verdict": v, "seconds": round(s, 1)} for n, v, s in res}
    json.dump(out, open(os.path.join(HERE, "rx_verdicts.json"), "w"), indent=1)
    print(out)
