"""GPU parity of the TX path (K7 frame synthesis, K8 mix) through the C-ABI / drop-in embedder
against reference-generated golden vectors and the TX oracle."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "rx_golden.npz"))


def _patched_embedder(key, seed):
    """WatermarkEmbedder with secrets.token_bytes replaced by the seeded stream the golden generator used."""
    import secrets
    from oracle import tx_oracle as txo
    from echoseal_b200 import embedder
    secrets.token_bytes = txo.seeded_rand(seed)
    return embedder.WatermarkEmbedder(key)


def test_frames_match_reference():
    for key_b, seed in ((0xAA, 52), (0x5C, 9)):
        tx = _patched_embedder(bytes([key_b]) * 32, seed)
        frames = []
        for ctr in (0, 1, 2, 255, 1024, 70000):
            tx.frame_ctr = ctr
            before = tx.frame_ctr
            frames.append(tx._make_frame_chips())
            assert tx.frame_ctr == before            # rtwm tests/test_embedder.py:82-91
        got = np.stack(frames)
        ref = G[f"tx/{key_b:02x}/frames"]
        assert got.shape == ref.shape == (6, 1215) and got.dtype == np.float32
        # TX parity bar: 1e-4 relative to the frame peak; in practice float32 round-off of an fp64 filter
        assert np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max()


def test_process_blocks_match_reference():
    for key_b, seed in ((0xAA, 52), (0x5C, 9)):
        tx = _patched_embedder(bytes([key_b]) * 32, seed + 1)
        rng = np.random.default_rng(seed)
        x = (0.1 * rng.standard_normal(8 * 1024)).astype(np.float32)
        x[2048:3072] *= 12.0
        x[4096:5120] = 0.0
        out = np.concatenate([tx.process(x[i:i + 1024]) for i in range(0, x.size, 1024)])
        ref = G[f"tx/{key_b:02x}/proc_out"]
        assert out.dtype == np.float32 and out.shape == ref.shape
        assert np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max()
        assert np.abs(out - ref).max() <= 1e-6
        assert tx.frame_ctr == 7                      # 8192 samples -> 7 frames generated


def test_whole_clip_matches_oracle_embedder():
    from _inputs import make_clip, CLIP_SPECS
    import secrets
    from oracle import tx_oracle as txo
    from echoseal_b200 import embedder
    def fac(key, seed):
        secrets.token_bytes = txo.seeded_rand(seed)
        return embedder.WatermarkEmbedder(key)
    for name in ("chirp_aa", "bench_17"):
        got, _ = make_clip(name, fac)
        ref, _ = make_clip(name)
        assert np.abs(got - ref).max() <= 1e-6


def test_bad_key_raises():
    from echoseal_b200 import embedder
    with pytest.raises(ValueError):
        embedder.WatermarkEmbedder(b"123")


def test_embedder_bank_matches_single_stream_semantics():
    """Lock-step multi-stream TX == S independent reference-style embedders (frames bit-for-bit given the
    same randomness; mix within float32 round-off), and the detector side can open what it seals."""
    import torch
    from echoseal_b200 import embedder
    from oracle import tx_oracle as txo
    keys = [bytes([i + 1]) * 32 for i in range(5)]
    stream = np.random.default_rng(0).integers(0, 256, 1 << 16, dtype=np.uint8).tobytes()
    pos = [0]
    def rand(n):
        b = stream[pos[0]:pos[0] + n]; pos[0] += n
        return b
    bank = embedder.EmbedderBank(keys, rand=rand)
    sn = bank.session_nonce.copy()
    bank.frame_ctr[:] = [0, 7, 1000, 65535, 2 ** 32 - 1]
    ctr0 = bank.frame_ctr.copy()
    rng = np.random.default_rng(1)
    x = (0.1 * rng.standard_normal((5, 3000))).astype(np.float32)
    # the bank draws 23 random bytes per frame, frame-major over (stream, frame): replay them for the oracle
    start = pos[0]
    out = bank.process(x)
    nf = 3                                      # ceil(3000 / 1215)
    rnd = np.frombuffer(stream[start:start + 23 * 5 * nf], np.uint8).reshape(5, nf, 23)
    for s in range(5):
        k = txo.Keys(keys[s])
        chips = np.concatenate([
            txo.frame_chips(k, (int(ctr0[s]) + f) % 2 ** 32,
                            txo.build_payload(k, (int(ctr0[s]) + f) % 2 ** 32, sn[s].tobytes(),
                                              rnd[s, f, :11].tobytes(), rnd[s, f, 11:].tobytes()))
            for f in range(nf)])
        ref = txo.mix(x[s], chips[:3000])
        assert np.abs(out[s] - ref).max() <= 1e-6
    assert list(bank.frame_ctr) == [3, 10, 1003, 65538, 2]
    # second block continues from the FIFO remainder
    out2 = bank.process(x[:, :500])
    assert out2.shape == (5, 500) and list(bank.frame_ctr) == [3, 10, 1003, 65538, 2]


def test_live_service_equals_bank_and_reports_latency():
    """TxService (pinned double-buffered staging, one CUDA stream, host crypto of the next frames prefetched in the
    background) gives block for block what the synchronous EmbedderBank gives with the same randomness."""
    from echoseal_b200 import embedder
    keys = [bytes([i + 3]) * 32 for i in range(7)]
    stream = np.random.default_rng(5).integers(0, 256, 1 << 18, dtype=np.uint8).tobytes()

    def make_rand():
        pos = [0]
        def rand(n):
            b = stream[pos[0]:pos[0] + n]; pos[0] += n
            return b
        return rand
    bank = embedder.EmbedderBank(keys, rand=make_rand())
    svc = embedder.TxService(keys, block=1024, rand=make_rand())
    rng = np.random.default_rng(2)
    prev = None
    for blk in range(9):                                   # crosses several frame boundaries (1215-sample frames)
        x = (0.1 * rng.standard_normal((7, 1024))).astype(np.float32)
        a = bank.process(x)
        b = svc.process_block(x)
        assert a.shape == b.shape == (7, 1024)
        assert (a == b).all(), blk
        if prev is not None:
            assert (prev[0] == prev[1]).all()              # the previous output buffer is still intact
        prev = (a, b.copy() if blk % 2 else b)
    assert (bank.frame_ctr == svc.bank.frame_ctr).all()
    assert len(svc.latency_ms) == 9 and all(t > 0 for t in svc.latency_ms)
    with pytest.raises(ValueError):
        svc.process_block(np.zeros((7, 512), np.float32))
