"""N>1 host logic on the CPU: world_size-2 (and 3) gloo process groups exercise the block partition and
the final verdict gather — the only communication of the path."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    from echoseal_b200.sharding import shard_range, verify_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys = [bytes([i % 251]) * 32 for i in range(n)]
    audio = np.arange(n * 4, dtype=np.float32).reshape(n, 4)
    seen = []
    def fake_verify(kb, ab):
        # stands in for the GPU path: verdict = f(key, audio) so misrouted shards are detected
        seen.append((len(kb), float(ab[:, 0].sum()) if len(kb) else 0.0))
        return np.array([(k[0] + int(a[0])) % 3 == 0 for k, a in zip(kb, ab)], bool)
    full = verify_sharded(keys, audio, fake_verify)
    lo, hi = shard_range(n, rank, world)
    q.put((rank, lo, hi, seen[0][0] if seen else 0, full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 11), (2, 8), (3, 7), (2, 1)])
def test_verify_sharded_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world * 131 + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [((i % 251) + 4 * i) % 3 == 0 for i in range(n)]
    covered = []
    for rank, lo, hi, nseen, full in sorted(res):
        assert full == want                      # every rank ends with the full verdict vector
        assert nseen == hi - lo                  # and verified only its own block
        covered += list(range(lo, hi))
    assert covered == list(range(n))             # blocks tile the batch exactly once


def test_shard_range_properties():
    from echoseal_b200.sharding import shard_range
    for n in (0, 1, 7, 10000, 10001):
        for w in (1, 2, 3, 8):
            ends = [shard_range(n, r, w) for r in range(w)]
            assert ends[0][0] == 0 and ends[-1][1] == n
            assert all(ends[i][1] == ends[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in ends]
            assert max(sizes) - min(sizes) <= 1
