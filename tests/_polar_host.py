"""Tiny pure-numpy polar ENCODER used only to build seeded test inputs
(CRC-8 poly 0x07 + F^{(x)10} butterflies, natural order; rtwm/fastpolar.py:237-252,
362-389).  Independent of the CUDA path and of oracle/ so inputs do not depend on
either side of a parity comparison."""
from __future__ import annotations
import numpy as np
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from echoseal_b200.polar_tables import data_positions

_POS = data_positions(1024, 448)


def crc8_bits(bits: np.ndarray) -> np.ndarray:
    reg = 0
    for b in bits:
        reg ^= (int(b) & 1) << 7
        reg = ((reg << 1) ^ 0x07) & 0xFF if reg & 0x80 else (reg << 1) & 0xFF
    return np.unpackbits(np.array([reg], np.uint8))


def polar_transform(u: np.ndarray) -> np.ndarray:
    x = u.astype(np.uint8).copy()
    n = x.size
    h = 1
    while h < n:
        v = x.reshape(-1, 2, h)
        v[:, 0, :] ^= v[:, 1, :]
        h *= 2
    return x


def encode_bits(info440: np.ndarray) -> np.ndarray:
    u = np.zeros(1024, np.uint8)
    u[_POS] = np.concatenate([info440.astype(np.uint8), crc8_bits(info440)])
    return polar_transform(u)
