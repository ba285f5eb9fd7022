"""CPU ORACLE for the RX scan stages — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of rtwm/detector.py's per-band scan: band-pass, normalised preamble
correlation, adaptive threshold, NMS peak pick, header decode, matched-filter taps, despread
-> LLR, candidate-counter enumeration and the 4-variant decode ladder.  Each function cites the
reference lines it follows.  The IIR band-pass (scipy.signal.lfilter 1.18.1 in the reference, a
third-party dependency) is restated in C (oracle/dsp_oracle.c, direct-form II transposed).

Parity status: PINNED against outputs of the reference itself run in the build container
(tests/golden/rx_golden.npz, made by tests/golden/make_rx_golden.py)."""
from __future__ import annotations
import ctypes as C
import hmac
import struct
import numpy as np

from . import polar_oracle as _po

PRE_L, HDR_BITS, HDR_REPEAT, HDR_L, N_POLAR = 63, 16, 8, 128, 1024
FRAME_LEN = PRE_L + HDR_L + N_POLAR            # 1215  (rtwm/detector.py:13-19)
TIGHT_DELTA, WIDE_DELTA = 3, 200               # rtwm/detector.py:20-21
MAX_TRIES, PEAK_LIMIT = 400, 25                # rtwm/detector.py:107-108
BAND_PLAN = [(4000, 6000), (8000, 10000), (16000, 18000), (18000, 22000)]   # rtwm/utils.py:19-24


def mseq_63() -> np.ndarray:
    """rtwm/utils.py:135-145"""
    state, seq = 0b111111, np.zeros(63, np.uint8)
    for i in range(63):
        newbit = ((state >> 5) ^ (state >> 4)) & 1
        seq[i] = state & 1
        state = ((state << 1) & 0b111111) | newbit
    return seq


def butter_bandpass(lo, hi, fs=48000, order=4):
    """rtwm/utils.py:52-55 — scipy.signal.butter is the reference's own filter designer."""
    from scipy.signal import butter
    nyq = 0.5 * fs
    return butter(order, [lo / nyq, hi / nyq], "band")


def lfilter(b, a, x, zi=None):
    """Direct-form II transposed IIR in float64 = scipy.signal.lfilter(b, a, x[, zi]) semantics
    (rtwm/detector.py:60, rtwm/embedder.py:141-143)."""
    b = np.ascontiguousarray(b, np.float64); a = np.ascontiguousarray(a, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty_like(x)
    n = max(b.size, a.size)
    z = np.zeros(n, np.float64)
    if zi is not None:
        z[: n - 1] = zi
    lib = _po.lib()
    lib.es_oracle_lfilter(b.ctypes.data_as(C.c_void_p), C.c_int(b.size), a.ctypes.data_as(C.c_void_p),
                          C.c_int(a.size), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                          C.c_long(x.size), z.ctypes.data_as(C.c_void_p))
    if zi is not None:
        return y, z[: n - 1].copy()
    return y


def choose_band_index(key: bytes, ctr: int) -> int:
    """rtwm/utils.py:27-36"""
    return hmac.new(key, struct.pack(">I", ctr), "sha256").digest()[0] % 4


def preamble_template(band, fs=48000):
    """rtwm/detector.py:67-69"""
    b, a = butter_bandpass(*band, fs)
    pre = 2.0 * mseq_63().astype(np.float32) - 1.0
    tpl = lfilter(b, a, lfilter(b, a, pre))
    return tpl / float(np.sqrt(np.sum(tpl * tpl)) + 1e-12)


def scan_band(signal: np.ndarray, band, fs=48000):
    """rtwm/detector.py:59-99: returns dict(y, corr, med, mad, thr, peaks, used_fallback) or None if the
    clip is shorter than the template (:72-73)."""
    b, a = butter_bandpass(*band, fs)
    y = lfilter(b, a, signal.astype(np.float32, copy=False))
    tpl = preamble_template(band, fs)
    L = tpl.size
    if y.size < L:
        return None
    y2 = y * y
    e_y = np.sqrt(np.convolve(y2, np.ones(L, dtype=np.float32), mode="valid")) + 1e-12
    corr = np.correlate(y, tpl, mode="valid") / e_y          # scipy.signal.correlate 'valid', real input
    med = float(np.median(corr))
    mad = float(np.median(np.abs(corr - med))) + 1e-12
    thr = min(med + 4.5 * 1.4826 * mad, 0.95)
    md = FRAME_LEN // 2
    peaks = []
    cand = np.flatnonzero(corr >= thr)
    for i in cand:
        lo = max(0, i - md); hi = min(corr.size, i + md + 1)
        if corr[i] >= corr[lo:hi].max():
            peaks.append(int(i))
    fallback = False
    if not peaks:
        k = min(5, corr.size)
        peaks = [int(v) for v in np.argsort(corr)[-k:][::-1]]
        fallback = True
    return dict(y=y, corr=corr, med=med, mad=mad, thr=thr, peaks=peaks, used_fallback=fallback)


def matched_filter_taps(band, fs=48000) -> np.ndarray:
    """rtwm/detector.py:260-294"""
    b, a = butter_bandpass(*band, fs)
    M = max(256, max(len(a), len(b)) * 64)
    imp = np.zeros(M, np.float32); imp[0] = 1.0
    g_tx = lfilter(b, a, imp).astype(np.float32)
    g_eff = np.convolve(g_tx, g_tx).astype(np.float32)
    e = g_eff * g_eff
    c = np.cumsum(e)
    total = float(c[-1]) + 1e-20
    idx = int(np.searchsorted(c, 0.999 * total))
    g_eff = g_eff[: idx + 1] if idx + 1 < g_eff.size else g_eff
    h = g_eff[::-1].copy()
    h /= (np.sqrt(float(np.sum(h * h))) + 1e-12)
    return h


def decode_header(frame: np.ndarray, h: np.ndarray, hdr_pn_sy: np.ndarray):
    """rtwm/detector.py:452-515 -> (ok, val, score, best_s, margin)."""
    seg = frame[PRE_L:PRE_L + HDR_L].astype(np.float32, copy=False)
    if seg.size < HDR_L:
        return False, 0, 0.0, 0, 0.0
    prefix_len = min(len(h) - 1, PRE_L)
    seg_full = np.concatenate((frame[PRE_L - prefix_len:PRE_L].astype(np.float32), seg)) if prefix_len > 0 else seg
    mf = np.convolve(seg_full, h, mode="full").astype(np.float32, copy=False)
    offset = (len(h) - 1) + prefix_len
    MAX_SHIFT = min(seg.size // 2 + prefix_len, 4 * len(h))
    mem = len(h) - 1
    if MAX_SHIFT < mem:
        MAX_SHIFT = mem
    start = max(0, offset - MAX_SHIFT)
    stop = min(mf.size, offset + seg.size + MAX_SHIFT)
    mf_win = mf[start:stop]
    base = offset - start
    guard = int(max(8, min(32, len(h) // 8)))
    best_s, best_score = 0, -1.0
    for s in range(-MAX_SHIFT, MAX_SHIFT + 1):
        i0, i1 = base + s, base + s + seg.size
        if i0 < 0 or i1 > mf_win.size:
            continue
        a_ = mf_win[i0:i1]
        score = abs(float(np.sum(a_[guard:] * hdr_pn_sy[guard:])))
        if score > best_score:
            best_score, best_s = score, s
    i0 = base + best_s
    d = mf_win[i0:i0 + seg.size] * hdr_pn_sy
    sums = d.reshape(HDR_BITS, HDR_REPEAT).sum(axis=1)
    bits = (sums < 0.0).astype(np.uint8)
    margin = float(np.mean(np.abs(sums)) / (np.sqrt(np.mean(d * d)) + 1e-12))
    val = 0
    for bb in bits:
        val = (val << 1) | int(bb)
    score = float(np.mean(np.abs(sums)) / (np.std(d) + 1e-12))
    ok = bool((np.count_nonzero(sums > 0) >= 10) and (margin > 0.5))
    return ok, val, score, best_s, margin


def llr(frame: np.ndarray, h: np.ndarray, pn_payload_bits: np.ndarray):
    """rtwm/detector.py:296-416 given the matched-filter taps and the payload PN bits of the chosen
    variant -> (llr float32[1024], best_s)."""
    N = N_POLAR
    pn_sy = 2.0 * pn_payload_bits.astype(np.float32) - 1.0
    mem = len(h) - 1
    payload_start = PRE_L + HDR_L
    if payload_start >= frame.size:
        return np.zeros(N, np.float32), 0
    rx_payload = frame[payload_start:].astype(np.float32, copy=False)
    prefix_len = min(mem, payload_start)
    rx_full = np.concatenate([frame[payload_start - prefix_len:payload_start].astype(np.float32), rx_payload]) \
        if prefix_len > 0 else rx_payload
    mf = np.convolve(rx_full, h, mode="full").astype(np.float32, copy=False)
    offset = prefix_len + mem
    n = min(pn_sy.size, rx_payload.size)
    pn_sy = pn_sy[:n]
    raw_shift = min(n // 2, 4 * len(h), HDR_L)
    MAX_SHIFT = max(mem, raw_shift)
    start = max(0, offset - MAX_SHIFT)
    stop = min(mf.size, offset + n + MAX_SHIFT)
    mf_win = mf[start:stop]
    base = offset - start
    guard = int(min(n // 4, max(len(h) // 2, 24)))
    if guard >= n:
        guard = max(0, n // 4)
    best_s, best_score = 0, -1.0
    for s in range(-MAX_SHIFT, MAX_SHIFT + 1):
        i0 = base + s; i1 = i0 + n
        if i0 < 0 or i1 > mf_win.size:
            continue
        d = mf_win[i0:i1] * pn_sy
        score = float(np.mean(np.abs(d[guard:])))
        if score > best_score:
            best_score, best_s = score, s
    i0 = base + best_s
    despread = mf_win[i0:i0 + n] * pn_sy
    tail = despread[guard:] if despread.size > guard + 8 else despread
    mu = float(np.mean(tail))
    llr_raw = despread - mu
    mad = float(np.median(np.abs(tail - float(np.median(tail))))) + 1e-12
    sigma = max(1.4826 * mad, float(np.std(tail)) + 1e-12, 0.1)
    scale = float(np.clip(2.0 / (sigma * sigma), 0.5, 30.0))
    out = np.clip(llr_raw * scale, -12.0, 12.0).astype(np.float32, copy=False)
    if out.size != N:
        o = np.zeros(N, np.float32); m = min(out.size, N); o[:m] = out[:m]; out = o
    return out, best_s


def candidate_counters(start: int, hdr_ok: bool, ctr_lo16: int, band_idx: int, band_of_ctr) -> list[int]:
    """rtwm/detector.py:117-142.  band_of_ctr(ctr) -> band index."""
    ctr_est = int(round(start / FRAME_LEN))
    cands = []
    if hdr_ok:
        for ctr in range(max(0, ctr_est - WIDE_DELTA), ctr_est + WIDE_DELTA + 1):
            if (ctr & 0xFFFF) == ctr_lo16 and band_of_ctr(ctr) == band_idx:
                cands.append(ctr)
    else:
        for ctr in range(max(0, ctr_est - TIGHT_DELTA), ctr_est + TIGHT_DELTA + 1):
            if band_of_ctr(ctr) == band_idx:
                cands.append(ctr)
        if not cands:
            for ctr in range(max(0, ctr_est - WIDE_DELTA), ctr_est + WIDE_DELTA + 1):
                if band_of_ctr(ctr) == band_idx:
                    cands.append(ctr)
    return cands


def verify(audio: np.ndarray, key32: bytes, list_size: int = 8, fs: int = 48000, session_nonce=None,
           return_details: bool = False):
    """Full restatement of WatermarkDetector.verify (rtwm/detector.py:44-233) for audio already at `fs`:
    hop-0 band first, then the others; per band the scan, <= 25 peaks, header, candidate counters, the
    400-try budget and the 4-variant SCL ladder with the AEAD validator / magic / counter / nonce latch.
    The SCL decodes of one band are batched through the C oracle (same results, validator applied in
    the reference's order afterwards)."""
    from . import tx_oracle as txo
    k = txo.Keys(key32)
    hdr_pn = 2.0 * k.pn_bits(0, 128).astype(np.float32) - 1.0
    hop = {}

    def band_of(c):
        v = hop.get(c)
        if v is None:
            v = hop[c] = k.band_index(c)
        return v

    hop0 = band_of(0)
    details = dict(attempts={}, n_scl=0)
    for bi in [hop0] + [b for b in range(4) if b != hop0]:
        band = BAND_PLAN[bi]
        r = scan_band(audio, band, fs)
        if r is None:
            continue
        h = matched_filter_taps(band, fs)
        att = []
        tried, stop = 0, False
        for start in r["peaks"][:PEAK_LIMIT]:
            if start + FRAME_LEN > r["y"].size:
                continue
            frame = r["y"][start:start + FRAME_LEN]
            ok, val, _, _, _ = decode_header(frame, h, hdr_pn)
            for ctr in candidate_counters(start, ok, val, bi, band_of):
                att.append((start, ctr)); tried += 1
                if tried >= MAX_TRIES:
                    stop = True
                    break
            if stop:
                break
        details["attempts"][bi] = att
        if not att:
            continue
        llrs = np.empty((4 * len(att), 1024), np.float32)
        for a_i, (start, ctr) in enumerate(att):
            frame = r["y"][start:start + FRAME_LEN]
            pn_full = k.pn_bits(ctr, FRAME_LEN)
            l0, _ = llr(frame, h, pn_full[PRE_L + HDR_L:])
            l1, _ = llr(frame, h, pn_full[:N_POLAR])
            llrs[4 * a_i + 0] = l0; llrs[4 * a_i + 1] = -l0
            llrs[4 * a_i + 2] = l1; llrs[4 * a_i + 3] = -l1
        res = _po.scl_batch(llrs, L=list_size)
        details["n_scl"] += llrs.shape[0]
        for a_i, (start, ctr) in enumerate(att):
            def validator(payload, ctr=ctr):
                try:
                    pt = k.open(payload)
                except Exception:
                    return False
                return pt.startswith(b"ESAL") and int.from_bytes(pt[4:8], "big") == ctr
            blob = None
            for v in range(4):
                bits, okv = _po.select(res, 4 * a_i + v, validator)
                if okv:
                    blob = np.packbits(bits).tobytes()
                    break
            if blob is None:
                continue
            pt = k.open(blob)
            nonce = pt[8:16]
            if session_nonce is None or nonce == session_nonce:
                details["nonce"] = nonce
                return (True, details) if return_details else True
    return (False, details) if return_details else False
