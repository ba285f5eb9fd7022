"""Multi-GPU sharding of the path: clips / codewords / streams are independent, so the
units are block-partitioned across ranks with NO collective on the data path; the only exchange is the
final verdict gather.  Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
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
