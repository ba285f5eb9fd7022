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


class _FakeDetector:
    """Duck-typed WatermarkDetector for the band-sharding logic: band scans are table look-ups."""
    def __init__(self, band_key, accept, nonces):
        self._band_key, self._accept, self._nonces = band_key, accept, nonces
        self.session_nonce = None
        self.scanned = []
    def _resample(self, audio, fs_in):
        return np.asarray(audio, np.float32)
    def _scan_band_multi_frame(self, signal, band):
        self.scanned.append(tuple(band))
        assert self.session_nonce is None                 # each band scan starts from the entry latch state
        if tuple(band) in self._accept:
            self.session_nonce = self._nonces[tuple(band)]
            return True
        return False


def _band_worker(rank, world, port, accept_idx, q):
    sys.path.insert(0, ROOT)
    from echoseal_b200.sharding import band_order, verify_recording_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    key = bytes(range(32))
    order = band_order(key)
    nonces = {b: bytes([65 + i]) * 8 for i, b in enumerate(order)}
    det = _FakeDetector(key, {order[i] for i in accept_idx}, nonces)
    v = verify_recording_sharded(det, np.zeros(100, np.float32), 48000)
    q.put((rank, v, det.session_nonce, [order.index(b) for b in det.scanned]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,accept_idx", [(2, ()), (2, (3,)), (2, (1, 2)), (3, (2,)), (4, (0, 3))])
def test_long_recording_band_sharding_gloo(world, accept_idx):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + world * 17 + sum(accept_idx) * 7 + len(accept_idx)) % 2000
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, accept_idx, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scanned = []
    for rank, v, nonce, idx in sorted(res):
        assert v == (len(accept_idx) > 0)                 # same verdict on every rank
        want = bytes([65 + min(accept_idx)]) * 8 if accept_idx else None
        assert nonce == want                              # latch = first accepting band in the reference order
        assert idx == [i for i in range(4) if i % world == rank]
        scanned += idx
    assert sorted(scanned) == [0, 1, 2, 3]                # every band scanned exactly once


def test_band_order_is_reference_order():
    sys.path.insert(0, ROOT)
    from echoseal_b200.sharding import band_order
    from echoseal_b200.utils import BAND_PLAN, choose_band
    for k in (bytes(32), bytes(range(32)), b"\xaa" * 32):
        o = band_order(k)
        assert o[0] == tuple(choose_band(k, 0)) and sorted(o) == sorted(tuple(b) for b in BAND_PLAN)
        assert o[1:] == [tuple(b) for b in BAND_PLAN if tuple(b) != o[0]]
