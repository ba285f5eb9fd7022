"""B200-native drop-in for rtwm/polar_fast.py and rtwm/fastpolar.PolarCode: same names, arguments,
return types and error behaviour (rtwm/polar_fast.py:26-87, rtwm/fastpolar.py:193-359); the arithmetic
runs in the sm_100a kernels (es_polar_encode / es_scl_hard / es_scl_list).  `decode_batch` is additive.
list_size 9..32 runs on the wide-list kernel; above 32 it is served with SCL-32 and a one-time warning
(polar_gpu.effective_list_size)."""
from __future__ import annotations
from typing import Callable, Optional, Tuple
import numpy as np
import torch

from . import polar_gpu

N_DEFAULT = 1024
K_DEFAULT = 448


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("echoseal_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _check_code(N: int, K: int, list_size: int, crc_size: int):
    # rtwm/fastpolar.py:210-217 (+ the table-length check of :10-13)
    if N <= 0 or (N & (N - 1)) != 0:
        raise ValueError("N must be a power of 2 and > 0")
    if not (0 < K <= N):
        raise ValueError("0 < K <= N must hold")
    if list_size < 1:
        raise ValueError("list_size must be >= 1")
    if not (0 < crc_size < K):
        raise ValueError("0 < crc_size < K must hold")
    if N != 1024:
        raise ValueError(f"Q_Nmax must have {N} entries (has 1024)")
    if crc_size != 8 or K % 8:
        raise ValueError("the B200 path implements CRC-8 and byte-aligned K only")
    return polar_gpu.effective_list_size(list_size)


def select(hard_payload, hard_crc, path_payload, path_crc, npaths, validator=None):
    """The reference's selection rule (rtwm/fastpolar.py:261-276, 332-359) applied to one codeword's
    kernel outputs.  Returns (payload bytes, ok)."""
    def _val(p):
        if validator is None:
            return True
        try:
            return bool(validator(bytes(p)))
        except Exception:
            return False
    if hard_crc and _val(hard_payload):
        return bytes(hard_payload), True
    best_crc = None
    for a in range(int(npaths)):
        if path_crc[a]:
            if _val(path_payload[a]):
                return bytes(path_payload[a]), True
            if best_crc is None:
                best_crc = bytes(path_payload[a])
    if best_crc is not None:
        return best_crc, False
    # lowest-metric non-CRC path (paths are already in ascending-metric order)
    for a in range(int(npaths)):
        if not path_crc[a]:
            return bytes(path_payload[a]), False
    return bytes(hard_payload), False


def decode_batch(llr, *, K: int = K_DEFAULT, list_size: int = 8, skip_list_on_hard_crc: bool = True):
    """llr float32[n,1024] (numpy or CUDA tensor) -> list of (payload bytes, ok) with validator=None
    semantics (hard-decision fast path first, rtwm/fastpolar.py:269-276)."""
    list_size = _check_code(1024, K, list_size, 8)
    dev = _dev()
    t = llr if isinstance(llr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(llr, np.float32))
    t = t.to(device=dev, dtype=torch.float32).contiguous()
    pay_h, crc_h = polar_gpu.hard_decide(t, K=K)
    index = None
    if skip_list_on_hard_crc:
        index = torch.nonzero(crc_h == 0, as_tuple=False).flatten().to(torch.int32)
    out = polar_gpu.list_decode(t, list_size=list_size, index=index, K=K)
    ph, ch = pay_h.cpu().numpy(), crc_h.cpu().numpy()
    pp, pc, npth = out["payload"].cpu().numpy(), out["crc"].cpu().numpy(), out["npaths"].cpu().numpy()
    return [select(ph[w], ch[w], pp[w], pc[w], npth[w]) for w in range(t.shape[0])]


def encode(payload: bytes, *, N: int = N_DEFAULT, K: int = K_DEFAULT, list_size: int = 8, crc_size: int = 8,
           debug: bool = False) -> np.ndarray:
    """rtwm/polar_fast.py:26-53 -> uint8[1024] code bits"""
    list_size = _check_code(N, K, list_size, crc_size)
    info_bytes = (K - crc_size) // 8
    if len(payload) != info_bytes:
        raise ValueError(f"payload must be {info_bytes} bytes (got {len(payload)})")
    p = torch.from_numpy(np.frombuffer(payload, np.uint8).copy()[None]).to(_dev())
    bits, _ = polar_gpu.encode(p, K=K)
    return bits[0].cpu().numpy()


def decode(llr: np.ndarray, *, N: int = N_DEFAULT, K: int = K_DEFAULT, list_size: int = 8, crc_size: int = 8,
           return_ok: bool = False, debug: bool = False,
           validator: Optional[Callable[[bytes], bool]] = None) -> Optional[bytes] | Tuple[bytes, bool]:
    """rtwm/polar_fast.py:55-87"""
    list_size = _check_code(N, K, list_size, crc_size)
    llr = np.asarray(llr)
    if llr.ndim != 1 or llr.size != N:
        raise ValueError(f"LLR length {llr.size} != N {N}")
    t = torch.from_numpy(np.ascontiguousarray(llr, np.float32)[None]).to(_dev())
    pay_h, crc_h = polar_gpu.hard_decide(t, K=K)
    ph, ch = pay_h.cpu().numpy()[0], int(crc_h.cpu().numpy()[0])
    if ch and validator is None:
        payload, ok = bytes(ph), True
    else:
        out = polar_gpu.list_decode(t, list_size=list_size, K=K)
        payload, ok = select(ph, ch, out["payload"].cpu().numpy()[0], out["crc"].cpu().numpy()[0],
                             int(out["npaths"].cpu().numpy()[0]), validator)
    if return_ok:
        return payload, ok
    return None if not ok else payload


class PolarCode:
    """rtwm/fastpolar.py:193-359: encode(info_bits) -> uint8[N]; decode(llr, validator) -> (uint8[K-8], ok)."""

    def __init__(self, N: int, K: int, list_size: int = 8, crc_size: int = 8, debug: bool = False):
        list_size = _check_code(N, K, list_size, crc_size)
        self.N, self.K, self.list_size, self.crc_size, self.debug = N, K, list_size, crc_size, debug
        from .polar_tables import frozen_mask, data_positions
        self.frozen = frozen_mask(N, K)
        self._data_pos = data_positions(N, K)
        self._info_len = K - crc_size

    def encode(self, info_bits: np.ndarray) -> np.ndarray:
        info_bits = np.asarray(info_bits)
        if info_bits.ndim != 1:
            raise ValueError("info_bits must be a 1D array")
        if info_bits.size != self._info_len:
            raise ValueError(f"info_bits must have length {self._info_len}")
        return encode(np.packbits(info_bits.astype(np.uint8)).tobytes(), N=self.N, K=self.K,
                      list_size=self.list_size, crc_size=self.crc_size)

    def decode(self, llr: np.ndarray, validator=None) -> Tuple[np.ndarray, bool]:
        llr = np.asarray(llr)
        if llr.ndim != 1 or llr.size != self.N:
            raise ValueError(f"llr must be 1D length {self.N}")
        payload, ok = decode(llr, N=self.N, K=self.K, list_size=self.list_size, crc_size=self.crc_size,
                             return_ok=True, validator=validator)
        return np.unpackbits(np.frombuffer(payload, np.uint8)), ok
