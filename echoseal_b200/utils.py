"""Host-side constants and helpers of the path — mirrors rtwm/utils.py (same names, same
argument meaning) with batched variants added for the GPU feeder.

Stays on the host by design (BASELINE.json north_star): band plan, Butterworth design (a host
constant uploaded to the kernels), the HMAC hop schedule and the AES-CTR PN generator."""
from __future__ import annotations
import hashlib
import hmac
import math
import struct
from functools import lru_cache
from typing import Tuple

import numpy as np
from scipy.signal import butter
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

# rtwm/utils.py:19-24
BAND_PLAN: list[Tuple[int, int]] = [
    (4_000, 6_000),
    (8_000, 10_000),
    (16_000, 18_000),
    (18_000, 22_000),
]


def choose_band_index(key: bytes, frame_ctr: int) -> int:
    return hmac.digest(key, struct.pack(">I", frame_ctr), "sha256")[0] % len(BAND_PLAN)


def choose_band(key: bytes, frame_ctr: int) -> tuple[int, int]:
    """Keyed hop selection, HMAC-SHA256(key, ctr_be32)[0] mod 4 (rtwm/utils.py:27-36)."""
    return BAND_PLAN[choose_band_index(key, frame_ctr)]


def hop_table(key: bytes, lo: int, hi: int) -> np.ndarray:
    """Band index for every counter in [lo, hi) — the batched form the detector's candidate
    enumeration needs (rtwm/detector.py:122-142 calls choose_band once per counter)."""
    out = np.empty(max(0, hi - lo), np.uint8)
    pack = struct.Struct(">I").pack
    dig = hmac.digest
    for i, c in enumerate(range(lo, hi)):
        out[i] = dig(key, pack(c), "sha256")[0] & 3
    return out


def db_to_lin(db: float) -> float:
    """rtwm/utils.py:40-42"""
    return 10.0 ** (db / 20.0)


def lin_to_db(lin: float) -> float:
    return 20.0 * np.log10(lin + 1e-12)


@lru_cache(maxsize=None)
def _butter_cached(lo: float, hi: float, fs: int, order: int):
    nyq = 0.5 * fs
    b, a = butter(order, [lo / nyq, hi / nyq], "band")
    b.setflags(write=False); a.setflags(write=False)
    return b, a


def butter_bandpass(lo: float, hi: float, fs: int, *, order: int = 4):
    """(b, a) float64 of the order-`order` Butterworth band-pass (rtwm/utils.py:52-55).
    A host constant: designed once with scipy (the reference's own dependency) and uploaded."""
    return _butter_cached(float(lo), float(hi), int(fs), int(order))


def resample_ratio(fs_target: int, fs_orig: int) -> tuple[int, int]:
    g = math.gcd(fs_orig, fs_target)
    return fs_target // g, fs_orig // g


def mseq_63() -> np.ndarray:
    """6-stage MLS, taps [6,5], seed 0b111111 (rtwm/utils.py:135-145)."""
    state = 0b111111
    seq = np.zeros(63, dtype=np.uint8)
    for i in range(63):
        newbit = ((state >> 5) ^ (state >> 4)) & 1
        seq[i] = state & 1
        state = ((state << 1) & 0b111111) | newbit
    return seq


class StreamPRNG:
    """AES-128-ECB counter stream: block(ctr, j) = AES_k((ctr << 64) | j), k = BLAKE2s-128(master,
    person="EchoSeal") (rtwm/utils.py:83-124)."""

    def __init__(self, master_key: bytes):
        sub_key = hashlib.blake2s(master_key, digest_size=16, person=b"EchoSeal").digest()
        self._cipher = Cipher(algorithms.AES(sub_key), modes.ECB())

    def bytes(self, frame_ctr: int, n: int = 64) -> bytes:
        nblk = (n + 15) // 16
        buf = b"".join(((frame_ctr << 64) + j).to_bytes(16, "big") for j in range(nblk))
        enc = self._cipher.encryptor()
        return (enc.update(buf) + enc.finalize())[:n]

    def bytes_batch(self, ctrs: np.ndarray, n: int) -> np.ndarray:
        """uint8[len(ctrs), n] — one ECB call over every (ctr, block) pair."""
        ctrs = np.asarray(ctrs, dtype=np.uint64).reshape(-1)
        nblk = (n + 15) // 16
        blocks = np.zeros((ctrs.size, nblk, 2), dtype=">u8")
        blocks[:, :, 0] = ctrs[:, None]
        blocks[:, :, 1] = np.arange(nblk, dtype=np.uint64)[None, :]
        enc = self._cipher.encryptor()
        raw = enc.update(blocks.tobytes()) + enc.finalize()
        return np.frombuffer(raw, dtype=np.uint8).reshape(ctrs.size, nblk * 16)[:, :n]


def pn_bits(prng: StreamPRNG, frame_ctr: int, n_bits: int) -> np.ndarray:
    """rtwm/utils.py:127-132"""
    data = prng.bytes(frame_ctr, (n_bits + 7) // 8)
    return np.unpackbits(np.frombuffer(data, dtype="u1"))[:n_bits]
