"""Tensor-level wrappers over the RX scan kernels of the C-ABI (include/echoseal_b200.h:
es_rx_set_filters / es_rx_bandpass / es_rx_ncc / es_rx_peaks / es_rx_frames / es_rx_llr).
Used by echoseal_b200.detector; no CPU fallback."""
from __future__ import annotations
import ctypes as C
import numpy as np
import torch
from scipy.signal import lfilter as _host_lfilter   # host constants only (taps/templates), never signal data

from . import _native as N
from .utils import BAND_PLAN, butter_bandpass, mseq_63

PRE_L, HDR_L, NPAY, FRAME_LEN, PEAK_LIMIT, MAXH = 63, 128, 1024, 1215, 25, 192


def preamble_template(band, fs: int) -> np.ndarray:
    """float64[63] doubly band-passed, unit-norm MLS preamble (rtwm/detector.py:67-69). Host constant."""
    b, a = butter_bandpass(*band, fs, order=4)
    pre = 2.0 * mseq_63().astype(np.float32) - 1.0
    tpl = _host_lfilter(b, a, _host_lfilter(b, a, pre))
    return tpl / float(np.sqrt(np.sum(tpl * tpl)) + 1e-12)


def matched_filter_taps(band, fs: int) -> np.ndarray:
    """float32 taps of the time-reversed TX*RX cascade truncated at 99.9 % energy
    (rtwm/detector.py:260-294). Host constant, input of the K4 kernel."""
    b, a = butter_bandpass(*band, fs, order=4)
    M = max(256, max(len(a), len(b)) * 64)
    imp = np.zeros(M, dtype=np.float32)
    imp[0] = 1.0
    g_tx = _host_lfilter(b, a, imp).astype(np.float32)
    g_eff = np.convolve(g_tx, g_tx).astype(np.float32)
    e = g_eff * g_eff
    c = np.cumsum(e)
    total = float(c[-1]) + 1e-20
    idx = int(np.searchsorted(c, 0.999 * total))
    g_eff = g_eff[:idx + 1] if idx + 1 < g_eff.size else g_eff
    h = g_eff[::-1].copy()
    h /= (np.sqrt(float(np.sum(h * h))) + 1e-12)
    return h


_filter_sig = None


def set_filters(fs: int, mf_taps: list[np.ndarray]):
    """Upload the per-band constants (filter coefficients, preamble templates, matched-filter taps)."""
    global _filter_sig
    sig = (fs, torch.cuda.current_device(), tuple(t.tobytes() for t in mf_taps))
    if sig == _filter_sig:
        return
    bb = np.zeros((4, 9)); aa = np.zeros((4, 9)); tpl = np.zeros((4, PRE_L))
    mf = np.zeros((4, MAXH), np.float32); ml = np.zeros(4, np.int32)
    for i, band in enumerate(BAND_PLAN):
        b, a = butter_bandpass(*band, fs, order=4)
        bb[i] = b / a[0]; aa[i] = a / a[0]
        tpl[i] = preamble_template(band, fs)
        h = np.asarray(mf_taps[i], np.float32)
        if h.size > MAXH:
            raise ValueError(f"matched filter of band {band} has {h.size} taps > {MAXH}")
        mf[i, :h.size] = h; ml[i] = h.size
    p = lambda a_: a_.ctypes.data_as(C.c_void_p)
    N.check(N.lib().es_rx_set_filters(p(bb), p(aa), p(tpl), p(mf), p(ml)), "es_rx_set_filters")
    _filter_sig = sig


def bandpass_force_plain(on: bool):
    """tests: run K1 without the TMA store path (same chunk grid, so the two forms are comparable bit for bit)"""
    N.lib().es_rx_bandpass_force_plain(C.c_int(1 if on else 0))


def bandpass(x: torch.Tensor) -> torch.Tensor:
    """x float32[B,n] -> y float64[B,4,n] (K1)."""
    N.require_cuda(x)
    B, n = x.shape
    y = torch.empty((B, 4, n), dtype=torch.float64, device=x.device)
    with N.timed("bandpass"):
        N.check(N.lib().es_rx_bandpass(N.ptr(x), C.c_int(B), C.c_int(n), C.c_longlong(x.stride(0)), N.ptr(y),
                                       N.stream_ptr()), "es_rx_bandpass")
    return y


K3_LONG_MIN = 1 << 21      # correlation length from which the multi-CTA form of K3 is used
K3_TWO_PASS_MIN = 8192     # shorter rows take K3's general form, which does not use the producer's histogram
NCC_SPEC_CAP = 6144


def ncc(y: torch.Tensor, with_hist: bool = False):
    """y float64[B,4,n] -> corr float64[B,4,n-62] (K2).  with_hist: also returns K3's first pass, formed while the
    correlation values are at hand (es_rx_ncc_hist): (corr, aux) with aux = (hist u32[B*4,2048], central-bin values
    f64[B*4,6144], their count u32[B*4]) for peaks(corr, aux) — or aux = None when the row length does not use it."""
    N.require_cuda(y)
    B, _, n = y.shape
    nc = max(0, n - (PRE_L - 1))
    corr = torch.empty((B, 4, nc), dtype=torch.float64, device=y.device)
    aux = None
    if with_hist and K3_TWO_PASS_MIN <= nc < K3_LONG_MIN:
        aux = (torch.empty((B * 4, 2048), dtype=torch.int32, device=y.device),
               torch.empty((B * 4, NCC_SPEC_CAP), dtype=torch.float64, device=y.device),
               torch.empty((B * 4,), dtype=torch.int32, device=y.device))
    if nc > 0:
        with N.timed("ncc"):
            if aux is None:
                N.check(N.lib().es_rx_ncc(N.ptr(y), C.c_int(B), C.c_int(n), N.ptr(corr), N.stream_ptr()), "es_rx_ncc")
            else:
                N.check(N.lib().es_rx_ncc_hist(N.ptr(y), C.c_int(B), C.c_int(n), N.ptr(corr), N.ptr(aux[0]), N.ptr(aux[1]),
                                               N.ptr(aux[2]), N.stream_ptr()), "es_rx_ncc_hist")
    return (corr, aux) if with_hist else corr



def scan(x: torch.Tensor) -> torch.Tensor:
    """Band-pass + normalised correlation in one pass (K1+K2 fused, es_rx_scan): float32 [B, n] -> corr float64
    [B, 4, n-62]; the filtered signal is never written (rtwm/detector.py:59-60,75-79)."""
    B, n = x.shape
    nc = n - (PRE_L - 1)
    if x.stride(1) != 1:
        x = x.contiguous()
    N.require_cuda(x if x.is_contiguous() else x[:1, :1])      # rows may be a strided view of a wider buffer (x_stride)
    corr = torch.empty((B, 4, max(nc, 0)), dtype=torch.float64, device=x.device)
    if nc > 0:
        with N.timed("scan"):
            N.check(N.lib().es_rx_scan(N.ptr(x), C.c_int(B), C.c_int(n), C.c_longlong(x.stride(0)), N.ptr(corr),
                                       N.stream_ptr()), "es_rx_scan")
    return corr


def frames_x(x: torch.Tensor, pk: torch.Tensor, npk: torch.Tensor, hdr_pn: torch.Tensor):
    """Per-peak front end (K4) from the audio itself: the candidate frames are band-passed on their own
    (es_rx_frames_x), then decoded exactly as frames() does.  Same outputs as frames()."""
    B, n = x.shape
    if x.stride(1) != 1:
        x = x.contiguous()
    N.require_cuda(x if x.is_contiguous() else x[:1, :1], pk, npk, hdr_pn)
    dev = x.device
    out = dict(
        mf_aligned=torch.empty((B, 4, PEAK_LIMIT, NPAY), dtype=torch.float32, device=dev),
        llr_best_s=torch.zeros((B, 4, PEAK_LIMIT), dtype=torch.int32, device=dev),
        hdr=torch.zeros((B, 4, PEAK_LIMIT, 4), dtype=torch.float32, device=dev),
        hdr_best_s=torch.zeros((B, 4, PEAK_LIMIT), dtype=torch.int32, device=dev),
    )
    yfr = torch.empty((B, 4, PEAK_LIMIT, FRAME_LEN), dtype=torch.float64, device=dev)
    with N.timed("frames"):
        N.check(N.lib().es_rx_frames_x(N.ptr(x), C.c_int(B), C.c_int(n), C.c_longlong(x.stride(0)), N.ptr(pk), N.ptr(npk),
                                       N.ptr(hdr_pn), N.ptr(yfr), N.ptr(out["mf_aligned"]), N.ptr(out["llr_best_s"]),
                                       N.ptr(out["hdr"]), N.ptr(out["hdr_best_s"]), N.stream_ptr()), "es_rx_frames_x")
    return out


def peaks_force_general(on: bool):
    """Test hook: route every row of K3 through the general multi-pass form (same results as the two-pass form)."""
    N.lib().es_rx_peaks_force_general(C.c_int(1 if on else 0))


def peaks(corr: torch.Tensor, aux=None):
    """corr float64[B,4,nc] -> (peaks i32[B,4,25] (-1 padded), npeaks i32[B,4], stats f64[B,4,4] =
    med, mad, thr, used_fallback) (K3).  Long recordings (nc >= 2^21) take the multi-CTA form.  aux: the second result of
    ncc(y, with_hist=True) for this corr — K3 then reads corr once instead of twice."""
    N.require_cuda(corr)
    B, _, nc = corr.shape
    pk = torch.full((B, 4, PEAK_LIMIT), -1, dtype=torch.int32, device=corr.device)
    npk = torch.zeros((B, 4), dtype=torch.int32, device=corr.device)
    st = torch.zeros((B, 4, 4), dtype=torch.float64, device=corr.device)
    if nc <= 0:
        return pk, npk, st
    if nc >= K3_LONG_MIN:
        lib = N.lib()
        lib.es_rx_peaks_long_scratch_bytes.restype = C.c_size_t
        nbytes = int(lib.es_rx_peaks_long_scratch_bytes(C.c_int(B), C.c_int(nc)))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=corr.device)
        ovf = torch.zeros(1, dtype=torch.int32, device=corr.device)
        with N.timed("peaks_long", 17):
            N.check(lib.es_rx_peaks_long(N.ptr(corr), C.c_int(B), C.c_int(nc), N.ptr(scratch), C.c_size_t(nbytes),
                                         N.ptr(pk), N.ptr(npk), N.ptr(st), N.ptr(ovf), N.stream_ptr()), "es_rx_peaks_long")
        if int(ovf.item()) == 0:
            return pk, npk, st
    with N.timed("peaks"):
        if aux is not None and nc < K3_LONG_MIN:
            N.check(N.lib().es_rx_peaks_hist(N.ptr(corr), C.c_int(B), C.c_int(nc), N.ptr(aux[0]), N.ptr(aux[1]), N.ptr(aux[2]),
                                             N.ptr(pk), N.ptr(npk), N.ptr(st), N.stream_ptr()), "es_rx_peaks_hist")
        else:
            N.check(N.lib().es_rx_peaks(N.ptr(corr), C.c_int(B), C.c_int(nc), N.ptr(pk), N.ptr(npk), N.ptr(st),
                                        N.stream_ptr()), "es_rx_peaks")
    return pk, npk, st


def frames(y: torch.Tensor, pk: torch.Tensor, npk: torch.Tensor, hdr_pn: torch.Tensor):
    """Per-peak front end (K4).  hdr_pn uint8[B,16] = packed 128 header-PN bits of each clip's key.
    Returns dict(mf_aligned f32[B,4,25,1024], llr_best_s i32[B,4,25], hdr f32[B,4,25,4] = ok(-1 = no
    frame), val, score, margin, hdr_best_s i32[B,4,25])."""
    N.require_cuda(y, pk, npk, hdr_pn)
    B, _, n = y.shape
    dev = y.device
    out = dict(
        mf_aligned=torch.empty((B, 4, PEAK_LIMIT, NPAY), dtype=torch.float32, device=dev),
        llr_best_s=torch.zeros((B, 4, PEAK_LIMIT), dtype=torch.int32, device=dev),
        hdr=torch.zeros((B, 4, PEAK_LIMIT, 4), dtype=torch.float32, device=dev),
        hdr_best_s=torch.zeros((B, 4, PEAK_LIMIT), dtype=torch.int32, device=dev),
    )
    with N.timed("frames"):
        N.check(N.lib().es_rx_frames(N.ptr(y), C.c_int(B), C.c_int(n), N.ptr(pk), N.ptr(npk), N.ptr(hdr_pn),
                                     N.ptr(out["mf_aligned"]), N.ptr(out["llr_best_s"]), N.ptr(out["hdr"]),
                                     N.ptr(out["hdr_best_s"]), N.stream_ptr()), "es_rx_frames")
    return out


def llr(mf_aligned: torch.Tensor, item_peak: torch.Tensor, pn_packed: torch.Tensor) -> torch.Tensor:
    """K5: item_peak i32[I] = flat (clip*4+band)*25+slot index, pn_packed uint8[I,152] (1215 PN bits,
    MSB-first) -> llr f32[2*I,1024], row 2i = PN variant 0, row 2i+1 = variant 1 (rtwm/detector.py:306-314)."""
    N.require_cuda(mf_aligned, item_peak, pn_packed)
    I = int(item_peak.numel())
    out = torch.empty((2 * I, NPAY), dtype=torch.float32, device=mf_aligned.device)
    if I:
        if pn_packed.shape != (I, 152) or pn_packed.dtype != torch.uint8:
            raise ValueError("pn_packed must be uint8 [items,152]")
        with N.timed("llr"):
            N.check(N.lib().es_rx_llr(N.ptr(mf_aligned), N.ptr(item_peak), N.ptr(pn_packed), C.c_int(I), N.ptr(out),
                                      N.stream_ptr()), "es_rx_llr")
    return out


_resample_cache = {}


def resample_design(up: int, down: int, f64: bool):
    """Host constants of scipy.signal.resample_poly(x, up, down) (rtwm/utils.py:58-66): the zero-pre-padded
    Kaiser(5.0) low-pass scaled by `up`, in polyphase order [up][J], and the number of leading outputs to drop."""
    from scipy.signal import firwin
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0))
    if not f64:
        h = h.astype(np.float32)            # scipy matches the dtype of x
    h = h * up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    hp = np.concatenate([np.zeros(n_pre_pad, h.dtype), h]).astype(np.float64)
    J = -(-hp.size // up)
    poly = np.zeros((up, J), np.float64)
    for ph in range(up):
        seg = hp[ph::up]
        poly[ph, :seg.size] = seg
    return poly, J, n_pre_remove


def resample(x: torch.Tensor, fs_in: int, fs_out: int) -> torch.Tensor:
    """K9: x float32/float64 [B, n_in] on the GPU -> float32 [B, n_out] at fs_out."""
    import math
    N.require_cuda(x)
    if x.dtype not in (torch.float32, torch.float64) or x.dim() != 2:
        raise ValueError("x must be float32/float64 [B, n]")
    g = math.gcd(fs_in, fs_out)
    up, down = fs_out // g, fs_in // g
    if up == down == 1:
        return x.to(torch.float32)
    f64 = x.dtype == torch.float64
    key = (up, down, f64, x.device.index)
    if key not in _resample_cache:
        poly, J, nrem = resample_design(up, down, f64)
        _resample_cache[key] = (torch.from_numpy(poly).to(x.device), J, nrem)
    taps, J, nrem = _resample_cache[key]
    B, n_in = x.shape
    n_out = n_in * up
    n_out = n_out // down + (1 if n_out % down else 0)
    y = torch.empty((B, n_out), dtype=torch.float32, device=x.device)
    with N.timed("resample"):
        N.check(N.lib().es_rx_resample(N.ptr(x), C.c_int(1 if f64 else 0), C.c_int(B), C.c_longlong(n_in),
                                       C.c_longlong(x.stride(0)), C.c_int(up), C.c_int(down), N.ptr(taps), C.c_int(J),
                                       C.c_longlong(nrem), C.c_longlong(n_out), N.ptr(y), N.stream_ptr()), "es_rx_resample")
    return y
