"""Multi-GPU sharding of the path: clips / codewords / streams are independent, so the
units are block-partitioned across ranks with NO collective on the data path; the only exchange is the
final verdict gather.  Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).

One LONG recording (config 3) shards by BAND: the reference scans the four bands independently
(rtwm/detector.py:44-53: one `_scan_band_multi_frame` per band, each with its own median / MAD threshold, NMS,
25-peak limit and 400-try budget), so band i of the reference's order goes to rank i mod world and again the
only exchange is the gather of the (verdict, latched nonce) records.  Splitting the recording in TIME across
GPUs (halos + all-reduced histograms for the global order statistics) lives in long_sharded.py."""
from __future__ import annotations
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition: rank r owns [lo, hi); sizes differ by at most one."""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_verdicts(local: np.ndarray, n_units: int, device=None) -> np.ndarray:
    """All ranks get the full bool[n_units] verdict vector (uint8 all_gather of padded shards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local, bool)
    world, rank = dist.get_world_size(), dist.get_rank()
    width = (n_units + world - 1) // world
    dev = device if device is not None else torch.device("cpu")
    mine = torch.zeros(width, dtype=torch.uint8, device=dev)
    lo, hi = shard_range(n_units, rank, world)
    mine[: hi - lo] = torch.from_numpy(np.asarray(local, np.uint8)).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    full = np.zeros(n_units, bool)
    for r in range(world):
        a, b = shard_range(n_units, r, world)
        full[a:b] = out[r][: b - a].cpu().numpy().astype(bool)
    return full


def verify_sharded(keys, audio, verify_fn, device=None) -> np.ndarray:
    """Every rank verifies its own block of clips with `verify_fn(keys_block, audio_block) -> bool[]`
    and the verdicts are gathered; no other communication."""
    n = len(keys)
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n, rank, world)
    local = verify_fn(keys[lo:hi], audio[lo:hi]) if hi > lo else np.zeros(0, bool)
    return gather_verdicts(local, n, device)


def band_order(band_key: bytes):
    """The reference's scan order: the hop-0 band first, then the rest in BAND_PLAN order (rtwm/detector.py:46-52)."""
    from .utils import BAND_PLAN, choose_band
    hop0 = tuple(choose_band(band_key, 0))
    return [hop0] + [tuple(b) for b in BAND_PLAN if tuple(b) != hop0]


def verify_recording_sharded(det, audio, fs_in: int, device=None) -> bool:
    """`det.verify(audio, fs_in)` for ONE recording with its four band scans spread over the ranks.
    `det` is a WatermarkDetector (same key on every rank); every rank passes the same audio.  Returns the same
    verdict on every rank and latches `det.session_nonce` exactly as the sequential scan would: from the first
    band, in the reference's order, that accepted a frame."""
    rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    order = band_order(det._band_key)
    signal = det._resample(np.asarray(audio), fs_in)
    if isinstance(signal, torch.Tensor):
        signal = signal.detach().cpu().numpy()
    signal = np.asarray(signal, dtype=np.float32).reshape(-1)
    rec = np.zeros((len(order), 10), np.uint8)           # per band: verdict, has_nonce, nonce[8]
    entry_nonce = det.session_nonce
    for i, band in enumerate(order):
        if i % world != rank:
            continue
        det.session_nonce = entry_nonce                  # every band scan starts from the caller's latch state
        ok = bool(det._scan_band_multi_frame(signal, band))
        rec[i, 0] = 1 if ok else 0
        if ok and det.session_nonce:
            rec[i, 1] = 1
            rec[i, 2:] = np.frombuffer(det.session_nonce, np.uint8)
    det.session_nonce = entry_nonce
    if world > 1:
        dev = device if device is not None else torch.device("cpu")
        mine = torch.from_numpy(rec).to(dev)
        out = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(out, mine)
        rec = np.zeros_like(rec)
        for r in range(world):
            part = out[r].cpu().numpy()
            rec[r::world] = part[r::world]               # rank r filled rows r, r + world, ...
    for i in range(len(order)):                          # first accepting band in the reference's order wins
        if rec[i, 0]:
            if rec[i, 1]:
                det.session_nonce = rec[i, 2:].tobytes()
            return True
    return False
