#!/usr/bin/env python3
"""Generate golden vectors for the RX scan path and the TX path by RUNNING the reference
(`/root/reference/rtwm`) in this container.  Inputs are rebuilt from seeds by tests/_inputs.py;
this script asserts that the TX oracle used for that reproduces the reference embedder exactly,
then stores reference OUTPUTS (thresholds, peaks, header tuples, attempted counters, LLRs, ...).

    python tests/golden/make_rx_golden.py
"""
from __future__ import annotations
import contextlib, io, os, re, sys, hashlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")
import secrets
from _inputs import CLIP_SPECS, make_clip, FS
from oracle import tx_oracle as txo

import rtwm.detector as D
import rtwm.embedder as E
from rtwm.utils import BAND_PLAN, choose_band

N_LLR_FULL = 6        # attempts per band whose LLR vectors are stored in full


class RefEmbedderFactory:
    """reference WatermarkEmbedder with secrets.token_bytes replaced by the seeded stream."""
    def __call__(self, key, seed):
        rnd = txo.seeded_rand(seed)
        secrets.token_bytes = rnd          # rtwm.embedder and rtwm.crypto both call secrets.token_bytes
        with contextlib.redirect_stdout(io.StringIO()):
            tx = E.WatermarkEmbedder(key)
        class W:
            def __init__(s): s.tx = tx
            @property
            def frame_ctr(s): return s.tx.frame_ctr
            @frame_ctr.setter
            def frame_ctr(s, v): s.tx.frame_ctr = v
            def process(s, x):
                with contextlib.redirect_stdout(io.StringIO()):
                    return s.tx.process(x)
        return W()


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], np.uint8).copy()


def tap_clip(name, gold):
    audio_ref, key = make_clip(name, RefEmbedderFactory())
    audio_orc, _ = make_clip(name)
    assert np.array_equal(audio_ref, audio_orc), f"TX oracle != reference embedder on {name}"
    gold[f"{name}/audio_sha"] = sha(audio_ref)
    gold[f"{name}/audio_head"] = audio_ref[:64].copy()
    rx = D.WatermarkDetector(key, list_size=8)
    attempts = []
    llrs = {}
    hdrs = []
    orig_llr = rx._llr
    orig_hdr = rx._decode_header
    cur = {"band": None}

    def llr_tap(frame, frame_id, pn_variant=0):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            out = orig_llr(frame, frame_id, pn_variant)
        m = re.search(r"best_s=(-?\d+)", buf.getvalue())
        llrs.setdefault(cur["band"], []).append((frame_id, pn_variant, int(m.group(1)) if m else 0, out.copy()))
        return out

    def hdr_tap(frame, band):
        r = orig_hdr(frame, band)
        hdrs.append((BAND_PLAN.index(band), bool(r[0]), int(r[1]), float(r[2])))
        return r

    rx._llr = llr_tap
    rx._decode_header = hdr_tap
    D.polar_dec = lambda llr, **kw: None       # stub the decoder: enumerate every attempt, fast
    for bi, band in enumerate(BAND_PLAN):
        cur["band"] = bi
        hdrs.clear()
        with contextlib.redirect_stdout(io.StringIO()):
            # stage taps: re-execute the reference's own scan lines through its public helper
            ok = rx._scan_band_multi_frame(audio_ref, band)
        assert ok is False
        # stage values recomputed with the very same library calls the reference uses
        from scipy.signal import lfilter, correlate
        from rtwm.utils import butter_bandpass
        b, a = butter_bandpass(*band, 48000, order=4)
        y = lfilter(b, a, audio_ref.astype(np.float32, copy=False))
        tpl = lfilter(b, a, lfilter(b, a, rx._pre_sy)); tpl = tpl / float(np.sqrt(np.sum(tpl * tpl)) + 1e-12)
        pre = f"{name}/b{bi}/"
        if y.size >= 63:
            e_y = np.sqrt(np.convolve(y * y, np.ones(63, dtype=np.float32), mode="valid")) + 1e-12
            corr = correlate(y, tpl, mode="valid") / e_y
            med = float(np.median(corr)); mad = float(np.median(np.abs(corr - med))) + 1e-12
            thr = min(med + 4.5 * 1.4826 * mad, 0.95)
            peaks = []
            for i in np.flatnonzero(corr >= thr):
                lo = max(0, i - 607); hi = min(corr.size, i + 608)
                if corr[i] >= corr[lo:hi].max():
                    peaks.append(int(i))
            fb = 0
            if not peaks:
                peaks = [int(v) for v in np.argsort(corr)[-min(5, corr.size):][::-1]]; fb = 1
            gold[pre + "stats"] = np.array([med, mad, thr, float(len(peaks)), float(fb)])
            gold[pre + "peaks"] = np.array(peaks[:64], np.int64)
            gold[pre + "corr_at_peaks"] = corr[np.array(peaks[:64], np.int64)]
            gold[pre + "corr_sub"] = corr[::97].copy()
            gold[pre + "y_sub"] = y[::97].copy()
            gold[pre + "corr_minmax"] = np.array([corr.min(), corr.max()])
        gold[pre + "tpl"] = tpl
        gold[pre + "hdr"] = np.array([(h[1], h[2], h[3]) for h in hdrs], np.float64).reshape(-1, 3)
        L = llrs.get(bi, [])
        # every attempt is (ctr, variant 0) followed by (ctr, variant 1)
        gold[pre + "att_ctr"] = np.array([t[0] for t in L if t[1] == 0], np.int64)
        gold[pre + "att_best_s"] = np.array([[t[2] for t in L if t[1] == 0], [t[2] for t in L if t[1] == 1]], np.int32)
        gold[pre + "llr_sum"] = np.array([[float(np.sum(t[3], dtype=np.float64)) for t in L if t[1] == v] for v in (0, 1)])
        gold[pre + "llr_abs"] = np.array([[float(np.sum(np.abs(t[3]), dtype=np.float64)) for t in L if t[1] == v] for v in (0, 1)])
        full = [t[3] for t in L[: 2 * N_LLR_FULL]]
        gold[pre + "llr_full"] = np.stack(full) if full else np.zeros((0, 1024), np.float32)
        gold[pre + "mf_taps"] = rx._matched_filter_taps(band)
    gold[f"{name}/hop0"] = np.array([BAND_PLAN.index(choose_band(key, 0))])
    print(name, "done", {bi: len(llrs.get(bi, [])) // 2 for bi in range(4)}, flush=True)


def tx_vectors(gold):
    """TX known answers: frames for a few counters/keys with frozen randomness + process() blocks."""
    for key_b, seed in ((0xAA, 52), (0x5C, 9)):
        key = bytes([key_b]) * 32
        fac = RefEmbedderFactory()
        w = fac(key, seed)
        frames = []
        for ctr in (0, 1, 2, 255, 1024, 70000):
            w.tx.frame_ctr = ctr
            with contextlib.redirect_stdout(io.StringIO()):
                frames.append(w.tx._make_frame_chips())
        gold[f"tx/{key_b:02x}/frames"] = np.stack(frames)
        # block-wise process (1024-sample blocks as the live app, rtwm/audioio.py:18)
        w = fac(key, seed + 1)
        rng = np.random.default_rng(seed)
        x = (0.1 * rng.standard_normal(8 * 1024)).astype(np.float32)
        x[2048:3072] *= 12.0        # a loud block: exercises the headroom limiter
        x[4096:5120] = 0.0          # a silent block: exercises the absolute floor
        out = np.concatenate([w.process(x[i:i + 1024]) for i in range(0, x.size, 1024)])
        gold[f"tx/{key_b:02x}/proc_out"] = out.astype(np.float32)


def crypto_vectors(gold):
    from rtwm.crypto import SecureChannel
    for key_b in (0xAA, 0x01):
        key = bytes([key_b]) * 32
        sc = SecureChannel(key)
        gold[f"crypto/{key_b:02x}/pn"] = np.stack([np.packbits(sc.pn_bits(c, 1215)) for c in (0, 1, 255, 1024, 2 ** 31 + 5)])
        gold[f"crypto/{key_b:02x}/hop"] = np.array([BAND_PLAN.index(choose_band(key, c)) for c in range(512)], np.uint8)
        secrets.token_bytes = lambda n: bytes(range(n))
        gold[f"crypto/{key_b:02x}/seal"] = np.frombuffer(sc.seal(b"ESAL" + bytes(23)), np.uint8).copy()


def main():
    gold = {}
    crypto_vectors(gold)
    tx_vectors(gold)
    for name in CLIP_SPECS:
        tap_clip(name, gold)
    np.savez_compressed(os.path.join(HERE, "rx_golden.npz"), **gold)
    print("wrote rx_golden.npz", os.path.getsize(os.path.join(HERE, "rx_golden.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
