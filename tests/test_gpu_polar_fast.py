"""The polar_fast / PolarCode drop-ins, mirroring the reference's own polar tests
(reference tests/test_polar.py:40-107, tests/test_roundtrip.py:12-31,65-87,303-316) on the CUDA path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from _inputs import awgn_llr_set


@pytest.fixture(scope="module")
def pf():
    from echoseal_b200 import polar_fast
    return polar_fast


def test_noise_free_roundtrip(pf):
    # reference tests/test_polar.py:40-60: +-10 LLR, every payload recovered
    rng = np.random.default_rng(0)
    for _ in range(8):
        payload = rng.integers(0, 256, 55, dtype=np.uint8).tobytes()
        cw = pf.encode(payload)
        assert cw.dtype == np.uint8 and cw.shape == (1024,) and set(np.unique(cw)) <= {0, 1}
        llr = (2.0 * cw.astype(np.float32) - 1.0) * 10.0
        assert pf.decode(llr) == payload
        out, ok = pf.decode(llr, return_ok=True)
        assert ok and out == payload


def test_awgn_sigma_015(pf):
    # reference tests/test_polar.py:63-107: AWGN sigma = 0.15, default_rng(1234) / default_rng(4321)
    for seed in (1234, 4321):
        rng = np.random.default_rng(seed)
        payload = rng.integers(0, 256, 55, dtype=np.uint8).tobytes()
        cw = pf.encode(payload)
        rx = (2.0 * cw.astype(np.float64) - 1.0) + 0.15 * rng.standard_normal(1024)
        llr = (2.0 * rx / 0.15 ** 2).astype(np.float32)
        assert pf.decode(llr) == payload


def test_polarcode_class_and_validator(pf):
    pc = pf.PolarCode(1024, 448, list_size=8, crc_size=8)
    assert pc.frozen.sum() == 576 and not pc.frozen[0]          # least-reliable info set (quirk 1)
    rng = np.random.default_rng(2)
    info = rng.integers(0, 2, 440, dtype=np.uint8)
    cw = pc.encode(info)
    llr = (2.0 * cw.astype(np.float32) - 1.0) * 8.0
    bits, ok = pc.decode(llr)
    assert ok and (bits == info).all()
    # validator that rejects everything -> (bits, False), the CRC-passing candidate is still returned
    bits2, ok2 = pc.decode(llr, validator=lambda b: False)
    assert not ok2 and (bits2 == info).all()
    # validator that raises is swallowed like in the reference (rtwm/fastpolar.py:271-275)
    def boom(b):
        raise RuntimeError("x")
    bits3, ok3 = pc.decode(llr, validator=boom)
    assert not ok3
    # validator sees 55-byte payloads
    seen = []
    pc.decode(llr, validator=lambda b: seen.append(len(b)) or True)
    assert seen == [55]


def test_decode_batch_matches_oracle(pf):
    from oracle import polar_oracle as po
    llr, info = awgn_llr_set(256, seed=31)
    res = pf.decode_batch(llr)
    ref = po.scl_batch(llr, L=8)
    for w, (payload, ok) in enumerate(res):
        bits, rok = po.select(ref, w, None)
        assert ok == rok and payload == np.packbits(bits).tobytes()


def test_error_behaviour(pf):
    with pytest.raises(ValueError):
        pf.encode(b"short")                                  # rtwm/polar_fast.py:42-43
    with pytest.raises(ValueError):
        pf.decode(np.zeros(1000, np.float32))                # rtwm/polar_fast.py:73-74
    with pytest.raises(ValueError):
        pf.decode(np.zeros((2, 512), np.float32))
    with pytest.raises(ValueError):
        pf.PolarCode(1000, 448)                              # N not a power of two (rtwm/fastpolar.py:210-211)
    with pytest.raises(ValueError):
        pf.PolarCode(1024, 2000)                             # K > N
    with pytest.raises(ValueError):
        pf.PolarCode(1024, 448, list_size=0)
    pc32 = pf.PolarCode(1024, 448, list_size=32)             # the reference's own quick test uses 32: the wide-list kernel
    llr = (2.0 * pf.encode(bytes(range(55))).astype(np.float32) - 1.0) * 4.0
    bits, ok = pc32.decode(llr)
    assert ok and np.packbits(bits).tobytes() == bytes(range(55))
    with pytest.warns(RuntimeWarning):                       # the reference default (256) is accepted and served with SCL-32
        pc256 = pf.PolarCode(1024, 448, list_size=256)
    bits, ok = pc256.decode(llr)
    assert ok and np.packbits(bits).tobytes() == bytes(range(55))
    with pytest.raises(ValueError):
        pf.PolarCode(1024, 448).encode(np.zeros(100, np.uint8))


def test_other_K(pf):
    """K is a parameter of the reference API (polar_fast.encode(K=...)); K % 8 == 0 on this path."""
    from oracle import polar_oracle as po
    rng = np.random.default_rng(5)
    for K in (64, 256, 512):
        payload = rng.integers(0, 256, (K - 8) // 8, dtype=np.uint8).tobytes()
        cw = pf.encode(payload, K=K)
        assert (cw == po.encode(np.unpackbits(np.frombuffer(payload, np.uint8)), K=K)).all()
        llr = (2.0 * cw.astype(np.float32) - 1.0) * 6.0 + rng.normal(0, 1.0, 1024).astype(np.float32)
        out, ok = pf.decode(llr, K=K, return_ok=True)
        bits, rok = po.decode(llr, L=8, K=K)
        assert ok == rok and out == np.packbits(bits).tobytes()
    from echoseal_b200 import polar_gpu
    polar_gpu.set_code(1024, 448)
