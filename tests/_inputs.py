"""Seeded synthetic inputs shared by the golden generators, the parity tests and
bench.py.  Pure numpy; no reference / oracle import here."""
from __future__ import annotations
import numpy as np

N, K, INFO = 1024, 448, 440
SIGMAS = (0.15, 0.3, 0.4, 0.5)


def awgn_llr_set(n: int, seed: int = 7, sigmas=SIGMAS, encode=None):
    """Config 4: random 440-bit payloads, polar-encoded, BPSK (+1 ⇔ bit 1) + AWGN,
    llr = 2·rx/σ² as float32, σ cycling over `sigmas`.  Returns (llr[n,1024] f32, info[n,440] u8)."""
    if encode is None:
        from _polar_host import encode_bits as encode
    rng = np.random.default_rng(seed)
    llr = np.empty((n, N), np.float32)
    info = np.empty((n, INFO), np.uint8)
    for i in range(n):
        sig = sigmas[i % len(sigmas)]
        info[i] = rng.integers(0, 2, INFO, dtype=np.uint8)
        cw = encode(info[i])
        rx = (2.0 * cw.astype(np.float64) - 1.0) + sig * rng.standard_normal(N)
        llr[i] = (2.0 * rx / (sig * sig)).astype(np.float32)
    return llr, info


def detector_like_llr_set(n: int, seed: int = 11):
    """Tie-prone set (SURVEY.md §7 hard part 1): first half AWGN σ=0.7 around random
    codewords, second half `clip(N(0,2), ±12)` float32 (what `_llr` emits on garbage)."""
    from _polar_host import encode_bits
    rng = np.random.default_rng(seed)
    out = np.empty((n, N), np.float32)
    for i in range(n):
        if i < n // 2:
            cw = encode_bits(rng.integers(0, 2, INFO, dtype=np.uint8))
            rx = (2.0 * cw.astype(np.float64) - 1.0) + 0.7 * rng.standard_normal(N)
            out[i] = (2.0 * rx / 0.49).astype(np.float32)
        else:
            out[i] = np.clip(rng.normal(0.0, 2.0, N), -12.0, 12.0).astype(np.float32)
    return out
