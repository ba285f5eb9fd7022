"""Tensor-level wrappers over the TX kernels of the C-ABI (es_tx_set_filters / es_tx_frames /
es_tx_mix).  Used by echoseal_b200.embedder; no CPU fallback."""
from __future__ import annotations
import ctypes as C
import numpy as np
import torch

from . import _native as N
from . import polar_gpu
from .utils import BAND_PLAN, butter_bandpass, mseq_63

FRAME_LEN = 1215
_sig = None


def set_filters(fs: int, preamble_bits: np.ndarray | None = None):
    global _sig
    pre = np.ascontiguousarray(mseq_63() if preamble_bits is None else preamble_bits, dtype=np.uint8)
    if pre.size != 63:
        raise ValueError("the B200 path supports the 63-chip MLS preamble only")
    sig = (fs, torch.cuda.current_device(), pre.tobytes())
    if sig == _sig:
        return
    bb = np.zeros((4, 9)); aa = np.zeros((4, 9))
    for i, band in enumerate(BAND_PLAN):
        b, a = butter_bandpass(*band, fs, order=4)
        bb[i] = b / a[0]; aa[i] = a / a[0]
    p = lambda a_: a_.ctypes.data_as(C.c_void_p)
    N.check(N.lib().es_tx_set_filters(p(bb), p(aa), p(pre)), "es_tx_set_filters")
    _sig = sig


def frames(payload: torch.Tensor, pn: torch.Tensor, hdr_pn: torch.Tensor, band: torch.Tensor,
           ctr_lo16: torch.Tensor, K: int = 448) -> torch.Tensor:
    """K7: payload u8[F,55] (sealed), pn u8[F,152] (1215 PN bits MSB-first), hdr_pn u8[F,16],
    band i32[F], ctr_lo16 i32[F] -> chips f32[F,1215]."""
    N.require_cuda(payload, pn, hdr_pn, band, ctr_lo16)
    F = payload.shape[0]
    if payload.dtype != torch.uint8 or payload.shape[1] != (K - 8) // 8:
        raise ValueError(f"payload must be uint8 [F,{(K - 8) // 8}]")
    if pn.shape != (F, 152) or hdr_pn.shape != (F, 16) or pn.dtype != torch.uint8 or hdr_pn.dtype != torch.uint8:
        raise ValueError("pn must be uint8 [F,152] and hdr_pn uint8 [F,16]")
    if band.dtype != torch.int32 or ctr_lo16.dtype != torch.int32:
        raise ValueError("band / ctr_lo16 must be int32")
    polar_gpu.set_code(1024, K)
    chips = torch.empty((F, FRAME_LEN), dtype=torch.float32, device=payload.device)
    with N.timed("tx_frames"):
        N.check(N.lib().es_tx_frames(N.ptr(payload), N.ptr(pn), N.ptr(hdr_pn), N.ptr(band), N.ptr(ctr_lo16),
                                     C.c_int(F), N.ptr(chips), N.stream_ptr()), "es_tx_frames")
    return chips


def mix(x: torch.Tensor, chips: torch.Tensor, alpha: float, floor_scale: float):
    """K8: x, chips f32[S,B] -> (out f32[S,B], scale f32[S]) with the reference's per-block level rule."""
    N.require_cuda(x, chips)
    if x.shape != chips.shape or x.dtype != torch.float32 or chips.dtype != torch.float32 or x.dim() != 2:
        raise ValueError("x and chips must be float32 [S,B]")
    S, B = x.shape
    out = torch.empty_like(x)
    scale = torch.empty((S,), dtype=torch.float32, device=x.device)
    with N.timed("tx_mix"):
        N.check(N.lib().es_tx_mix(N.ptr(x), N.ptr(chips), C.c_int(S), C.c_int(B), C.c_double(alpha),
                                  C.c_double(floor_scale), N.ptr(out), N.ptr(scale), N.stream_ptr()), "es_tx_mix")
    return out, scale
