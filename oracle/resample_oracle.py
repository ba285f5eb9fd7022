"""CPU ORACLE for the resampler — TEST INFRASTRUCTURE ONLY.

numpy restatement of scipy.signal.resample_poly (scipy 1.18.1, third-party dependency of the reference;
call site rtwm/utils.py:58-66): Kaiser(5.0)-windowed sinc of 2*10*max(up,down)+1 taps (cast to the input
dtype, scaled by `up`, zero-pre-padded by down - half_len % down), zero-stuffed up-sampling, FIR,
down-sampling, and removal of (half_len + n_pre_pad) // down leading outputs.
Parity: pinned against scipy itself in tests/test_oracle_rx.py::test_resample_oracle_matches_scipy."""
from __future__ import annotations
import math
import numpy as np
from scipy.signal import firwin


def resample_poly(x: np.ndarray, up: int, down: int) -> np.ndarray:
    g = math.gcd(up, down)
    up //= g; down //= g
    if up == down == 1:
        return x.copy()
    n_in = x.size
    n_out = n_in * up
    n_out = n_out // down + bool(n_out % down)
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0))
    if x.dtype == np.float32:
        h = h.astype(np.float32)
    h = h * up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    hp = np.concatenate([np.zeros(n_pre_pad, h.dtype), h]).astype(np.float64)
    J = -(-hp.size // up)
    m = np.arange(n_out, dtype=np.int64)
    t = (m + n_pre_remove) * down
    i_hi = t // up
    phase = t - i_hi * up
    out = np.zeros(n_out, np.float64)
    xd = x.astype(np.float64)
    for j in range(J):
        idx = phase + j * up
        i = i_hi - j
        ok = (idx < hp.size) & (i >= 0) & (i < n_in)
        out[ok] += hp[idx[ok]] * xd[i[ok]]
    return out.astype(x.dtype)
