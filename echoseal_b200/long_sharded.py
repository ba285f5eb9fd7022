"""ONE long recording split in TIME over several GPUs (BASELINE config 3 at N > 1).

Rank r owns the correlation indices [a_r, b_r) of every band and holds the samples
[a_r - 607 - 768, b_r + 1215) of the recording: 768 samples of band-pass warm-up, the +-607 NMS window, the
62-sample template tail and the 1215-sample frame of a peak that starts in its range.  K1 / K2 are local.
The adaptive threshold of the reference is a GLOBAL order statistic (median and MAD of the whole band,
rtwm/detector.py:83-86), so the two-level histogram selection of es_rx_peaks_long runs phase by phase with three
exchanges per statistic: sum of the 2048-bin histogram, sum of the 2 x 2048 sub-bin histogram, union of the few
values in the selected sub-bin(s).  NMS is local given the halo; the first 25 peaks per band in global time
order are the first 25 of the merged per-rank lists; their 1215-sample frames are summed into one small array
(disjoint owners, zeros elsewhere) and every rank runs the tiny decode stage on it (at most 4 x 400 candidates), so
the verdict needs no further exchange.

The algorithm is written once as a per-rank generator that yields its exchange requests; `run_distributed` serves
them with torch.distributed (NCCL), `run_simulated` runs W virtual ranks in lock-step inside one process (one GPU)
-- that is how the GPU tests check W = 2, 3 against the single-GPU path without needing several GPUs.

Filtering a segment from zero state 768 samples early differs from filtering the whole recording by < 4e-13 of the
peak (the same bound as between the chunks of K1 on one GPU), so statistics agree to ~1e-12 and the sync offsets
are equal unless a correlation value sits within that distance of the threshold or of a neighbour.

When to use it: NOT for speed at BASELINE sizes.  A whole hour of audio is one 0.09-0.12 s pass on one B200
(bench.py --workload long); on 2 GPUs the time split took 0.49 s (profiles/r01_config3_long_x105_2gpu_sharded.json),
dominated by the host-side phase sequencing and by every rank resampling the recording.  It exists for recordings
that do not fit one GPU's memory (the four fp64 band signals cost 64 B per 48 kHz sample: ~47 min of audio per
10 GB) and as the one place of the path with a real exchange step.  Recordings shorter than one frame return False
without touching the ranks."""
from __future__ import annotations
import ctypes as C
import numpy as np
import torch

from . import _native as N
from . import rx_gpu
from .sharding import shard_range

PRE_L, FRAME_LEN, PEAK_LIMIT, NMS_HALF, WARM = 63, 1215, 25, 607, 768


def _layout(n_range: int):
    out = (C.c_longlong * 14)()
    N.check(N.lib().es_rx_peaks_long_layout(C.c_int(1), C.c_int(n_range), out), "es_rx_peaks_long_layout")
    return [int(v) for v in out]


def _rank_program(det, signal: np.ndarray, rank: int, world: int, dev, out: dict):
    """Generator: yields ("sum", tensor) -> tensor summed over ranks, ("cat", tensor) -> list of every rank's tensor."""
    n = int(signal.size)
    nc = n - (PRE_L - 1)
    a, b = shard_range(nc, rank, world)
    s0 = max(0, a - NMS_HALF - WARM)
    s1 = min(n, b + FRAME_LEN)
    x = torch.from_numpy(np.ascontiguousarray(signal[s0:s1], dtype=np.float32)).to(dev).reshape(1, -1)
    rx_gpu.set_filters(det.fs_target, det._taps())
    y = rx_gpu.bandpass(x)                       # [1, 4, nl]
    corr = rx_gpu.ncc(y)                         # [1, 4, nl - 62]; local index j <-> global s0 + j
    nloc = int(corr.shape[2])
    lo, hi = a - s0, b - s0
    lib = N.lib()
    lib.es_rx_peaks_long_scratch_bytes.restype = C.c_size_t
    nbytes = int(lib.es_rx_peaks_long_scratch_bytes(C.c_int(1), C.c_int(max(1, hi - lo))))
    scratch = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    L = _layout(max(1, hi - lo))
    off_meta, off_h1, off_h2, off_buf, _, meta_sz, o_cnt, _, o_ovf, _, o_med, o_mad, o_thr, cap = L
    hist1 = scratch[off_h1:off_h1 + 4 * 2048 * 4].view(torch.int32)
    hist2 = scratch[off_h2:off_h2 + 4 * 4096 * 4].view(torch.int32)
    buf = scratch[off_buf:off_buf + 4 * cap * 8].view(torch.float64).view(4, cap)
    meta = scratch[off_meta:off_meta + 4 * meta_sz].view(4, meta_sz)
    pk = torch.full((1, 4, PEAK_LIMIT), -1, dtype=torch.int32, device=dev)
    npk = torch.zeros((1, 4), dtype=torch.int32, device=dev)
    st = torch.zeros((1, 4, 4), dtype=torch.float64, device=dev)
    ovf = torch.zeros(1, dtype=torch.int32, device=dev)

    def phase(p):
        if hi > lo:
            with N.timed("peaks_long_phase", 3):
                N.check(lib.es_rx_peaks_long_phase(C.c_int(p), N.ptr(corr), C.c_int(1), C.c_int(nloc), C.c_int(lo), C.c_int(hi),
                                                   C.c_int(nc), N.ptr(scratch), C.c_size_t(nbytes), N.ptr(pk), N.ptr(npk),
                                                   N.ptr(st), N.ptr(ovf), N.stream_ptr()), "es_rx_peaks_long_phase")

    def cnt_view():
        return meta[:, o_cnt:o_cnt + 4].contiguous().view(torch.int32).view(4)

    def merge_gathered():
        """union of every rank's gathered values -> this rank's buffer, total count -> this rank's record"""
        cnt = cnt_view().clamp(max=cap)
        parts = yield ("cat", torch.cat([cnt.to(torch.float64).view(4, 1), buf], dim=1))
        tot = torch.zeros(4, dtype=torch.int64)
        for bi in range(4):
            vals = [p[bi, 1:1 + int(p[bi, 0].item())] for p in parts]
            v = torch.cat(vals) if vals else buf.new_zeros(0)
            m = min(int(v.numel()), cap)
            buf[bi, :m] = v[:m]
            tot[bi] = v.numel()
        meta[:, o_cnt:o_cnt + 4] = tot.to(torch.int32).to(dev).view(4, 1).view(torch.uint8).view(4, 4)
        return bool((tot > cap).any())

    over = False
    phase(0)
    hist1.copy_((yield ("sum", hist1.clone())))
    phase(1)
    hist2.copy_((yield ("sum", hist2.clone())))
    phase(2)
    over |= yield from merge_gathered()
    phase(3)
    hist1.copy_((yield ("sum", hist1.clone())))
    phase(4)
    hist2.copy_((yield ("sum", hist2.clone())))
    phase(5)
    over |= yield from merge_gathered()
    phase(6)
    # ---- merge the per-rank peak lists (global indices) / top-k candidates
    pk_l = pk[0].clone()
    valid = pk_l >= 0
    vals = torch.where(valid, corr[0].gather(1, pk_l.clamp(min=0).long()), torch.full_like(pk_l, -2, dtype=torch.float64))
    rec = torch.cat([torch.where(valid, pk_l + s0, pk_l).to(torch.float64), vals, npk[0].to(torch.float64).view(4, 1),
                     st[0, :, 3:4], ovf.to(torch.float64).expand(4).view(4, 1)], dim=1)          # [4, 25 + 25 + 3]
    if hi <= lo:
        rec[:, 2 * PEAK_LIMIT] = 0; rec[:, 2 * PEAK_LIMIT + 1] = 1
    recs = [r.cpu().numpy() for r in (yield ("cat", rec))]
    over = over or any(r[0, 2 * PEAK_LIMIT + 2] != 0 for r in recs)
    peaks = np.full((4, PEAK_LIMIT), -1, np.int64)
    npeaks = np.zeros(4, np.int32)
    fallback = np.zeros(4, np.int32)
    for bi in range(4):
        normal = [int(i) for r in recs if r[bi, 2 * PEAK_LIMIT + 1] == 0
                  for i in r[bi, :min(PEAK_LIMIT, int(r[bi, 2 * PEAK_LIMIT]))]]
        if normal:
            normal.sort()
            sel = normal[:PEAK_LIMIT]
        else:      # nobody exceeded the threshold: top min(5, nc) by (value desc, index desc) (rtwm/detector.py:98-99)
            fallback[bi] = 1
            cands = [(float(r[bi, PEAK_LIMIT + q]), int(r[bi, q])) for r in recs for q in range(PEAK_LIMIT) if r[bi, q] >= 0]
            cands.sort(key=lambda t: (-t[0], -t[1]))
            sel = [i for _, i in cands[:min(5, nc)]]
        peaks[bi, :len(sel)] = sel
        npeaks[bi] = len(sel)
    med = meta[:, o_med:o_med + 8].contiguous().view(torch.float64).view(4).cpu().numpy()
    mad = meta[:, o_mad:o_mad + 8].contiguous().view(torch.float64).view(4).cpu().numpy()
    thr = meta[:, o_thr:o_thr + 8].contiguous().view(torch.float64).view(4).cpu().numpy()
    # ---- frames of the selected peaks: owners contribute, everybody gets the compact array
    yc = torch.zeros((1, 4, PEAK_LIMIT * FRAME_LEN), dtype=torch.float64, device=dev)
    pkc = torch.full((1, 4, PEAK_LIMIT), -1, dtype=torch.int32, device=dev)
    for bi in range(4):
        for q in range(int(npeaks[bi])):
            g = int(peaks[bi, q])
            if g + FRAME_LEN > n:                 # rtwm/detector.py:112-113: no room for a frame -> skipped
                continue
            pkc[0, bi, q] = q * FRAME_LEN
            if a <= g < b:
                yc[0, bi, q * FRAME_LEN:(q + 1) * FRAME_LEN] = y[0, bi, g - s0:g - s0 + FRAME_LEN]
    yc = (yield ("sum", yc))
    out.update(dict(peaks=peaks, npeaks=npeaks, med=med, mad=mad, thr=thr, fallback=fallback, overflow=bool(over),
                    frames=yc, frame_index=pkc, n=n))


def _decode(det, res: dict, dev) -> bool:
    """K4-K6 + host validation on the compact frame array (identical on every rank)."""
    from .detector import _decode_phase
    hdr_pn = torch.from_numpy(np.packbits(det._hdr_pn_bits)[None]).to(dev)
    npk = torch.from_numpy(res["npeaks"].astype(np.int32)[None]).to(dev)
    fr = rx_gpu.frames(res["frames"], res["frame_index"], npk, hdr_pn)
    pk_true = res["peaks"].astype(np.int32)[None]
    enum = det._bank.rx_enumerate(np.zeros(1, np.int32), res["n"], pk_true, res["npeaks"].astype(np.int32)[None],
                                  fr["hdr"].cpu().numpy())
    ns = np.zeros((1, 9), np.uint8)
    if det.session_nonce:
        ns[0, 0] = 1; ns[0, 1:] = np.frombuffer(det.session_nonce, np.uint8)
    v, _ = _decode_phase(det._bank, np.zeros(1, np.int32), enum, fr["mf_aligned"], det._list_size, ns, dev)
    if ns[0, 0]:
        det.session_nonce = ns[0, 1:].tobytes()
    res["attempts"] = [len(enum["item_ctr"])]
    res["n_scl"] = 4 * int(enum["band_count"].sum())
    return bool(v[0])


def run_simulated(det, signal48: np.ndarray, world: int, dev=None):
    """W virtual ranks in lock-step on one GPU.  Returns (verdict, per-rank result dicts)."""
    dev = dev if dev is not None else torch.device("cuda", torch.cuda.current_device())
    signal48 = np.asarray(signal48, np.float32).reshape(-1)
    if signal48.size < 63 + 1215:
        # shorter than one frame after the template: the reference returns False without scanning (rtwm/detector.py:72-73,
        # :112-113); the rank programs need at least one correlation index per rank
        return False, [dict(verdict=False) for _ in range(world)]
    outs = [dict() for _ in range(world)]
    gens = [_rank_program(det, signal48, r, world, dev, outs[r]) for r in range(world)]
    reqs = [next(g) for g in gens]
    alive = True
    while alive:
        kind = reqs[0][0]
        assert all(r[0] == kind for r in reqs)
        if kind == "sum":
            total = torch.stack([r[1] for r in reqs]).sum(0)
            reply = [total.clone() for _ in gens]
        else:
            reply = [[r[1].clone() for r in reqs] for _ in gens]
        nxt = []
        for g, rp in zip(gens, reply):
            try:
                nxt.append(g.send(rp))
            except StopIteration:
                alive = False
        reqs = nxt
    if outs[0]["overflow"]:
        # degenerate data (e.g. long digital silence: > 32768 equal values in the median's sub-bin): the single-GPU
        # path has the radix-select fallback for that
        return bool(det.verify(signal48, det.fs_target)), outs
    return _decode(det, outs[0], dev), outs


def verify_recording_time_sharded(det, audio, fs_in: int, device=None) -> bool:
    """`det.verify(audio, fs_in)` with the recording split in time over the ranks of the default process group
    (every rank passes the same audio and gets the same verdict).  See the module docstring."""
    import torch.distributed as dist
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    signal = det._resample(np.asarray(audio), fs_in)
    if isinstance(signal, torch.Tensor):
        signal = signal.detach().cpu().numpy()
    signal = np.asarray(signal, np.float32).reshape(-1)
    if signal.size < 63 + 1215:
        return False                                  # every rank sees the same audio and returns the same answer
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return run_simulated(det, signal, 1, dev)[0]
    world, rank = dist.get_world_size(), dist.get_rank()
    out = {}
    g = _rank_program(det, signal, rank, world, dev, out)
    req = next(g)
    while True:
        if req[0] == "sum":
            t = req[1].contiguous()
            dist.all_reduce(t)
            reply = t
        else:
            t = req[1].contiguous()
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            reply = parts
        try:
            req = g.send(reply)
        except StopIteration:
            break
    det.last_sharded = out
    if out["overflow"]:       # same on every rank (the flag is part of the exchanged records): all fall back together
        return bool(det.verify(signal, det.fs_target))
    return _decode(det, out, dev)
