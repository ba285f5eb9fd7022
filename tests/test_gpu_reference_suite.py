"""The reference's own (passing) API tests, re-run against the B200 drop-ins with only the import
changed — tests/test_embedder.py:82-133,188-190, tests/test_embedder_functionality.py:49-89,
tests/test_embedder_detector_alignment.py:22-72, tests/test_false_positive.py:8-19,
tests/test_edge_cases.py:14-20,64-71, tests/test_crypto.py in the reference tree."""
import types
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from scipy.signal import lfilter

KEY = b"\xAA" * 32


@pytest.fixture(scope="module")
def api():
    from echoseal_b200.embedder import WatermarkEmbedder, TxParams
    from echoseal_b200.detector import WatermarkDetector, FRAME_LEN, PRE_L, HDR_L
    from echoseal_b200.utils import butter_bandpass, choose_band
    from echoseal_b200.polar_fast import encode as polar_encode
    from echoseal_b200.crypto import SecureChannel
    return types.SimpleNamespace(**locals())


def test_frame_counter_ownership(api):
    tx = api.WatermarkEmbedder(KEY, api.TxParams())
    start = tx.frame_ctr
    _ = tx._make_frame_chips()
    assert tx.frame_ctr == start
    _ = tx.process(np.zeros(1000, dtype=np.float32))
    assert tx.frame_ctr == start + 1


def test_payload_sealed_length_and_ctr_roundtrip(api):
    tx = api.WatermarkEmbedder(KEY, api.TxParams())
    for ctr in [0, 1, 7, 255]:
        tx.frame_ctr = ctr
        blob = tx._build_payload()
        assert isinstance(blob, bytes) and len(blob) == 55
        plain = api.SecureChannel(KEY).open(blob)
        assert plain.startswith(b"ESAL")
        assert int.from_bytes(plain[4:8], "big") == ctr


def test_no_clipping_headroom(api):
    tx = api.WatermarkEmbedder(KEY, api.TxParams())
    out = tx.process(np.full(4096, 0.97, dtype=np.float32))
    assert float(np.max(np.abs(out))) <= 0.98001


def test_choose_band_is_deterministic(api):
    for ctr in [0, 1, 5, 17, 255]:
        assert api.choose_band(KEY, ctr) == api.choose_band(KEY, ctr)


def test_process_function_injects_chips(api):
    fsig = np.zeros(2000, dtype=np.float32)
    tx = api.WatermarkEmbedder(KEY, api.TxParams())
    out1 = tx.process(fsig)
    assert out1.shape == fsig.shape and out1.dtype == np.float32
    assert not np.allclose(out1, fsig)
    out2 = tx.process(fsig)
    assert out2.shape == fsig.shape and not np.allclose(out2, fsig)
    assert not np.allclose(out1, out2)              # buffer rollover: different chips


def test_payload_uniqueness_and_decrypt(api):
    tx = api.WatermarkEmbedder(KEY, api.TxParams())
    seen = set()
    for ctr in range(4):
        tx.frame_ctr = ctr
        blob = tx._build_payload()
        plain = api.SecureChannel(KEY).open(blob)
        assert plain.startswith(b"ESAL") and int.from_bytes(plain[4:8], "big") == ctr
        assert blob not in seen
        seen.add(blob)


def test_embedder_detector_sequences_align(api):
    key = bytes.fromhex("00" * 32)
    embedder, detector = api.WatermarkEmbedder(key), api.WatermarkDetector(key)
    np.testing.assert_allclose(embedder._preamble_sy, detector._pre_sy)
    np.testing.assert_allclose(embedder._hdr_pn_sy, detector._hdr_pn_sy)
    for ctr in (0, 1, 255, 1024):
        pn_tx = embedder.sec.pn_bits(ctr, api.PRE_L + api.HDR_L + embedder.p.N)
        pn_rx = detector.sec.pn_bits(ctr, api.FRAME_LEN)
        np.testing.assert_array_equal(pn_tx, pn_rx)


def test_embedder_frame_filtering_matches_spec(api):
    key = bytes.fromhex("00" * 32)
    embedder = api.WatermarkEmbedder(key)
    embedder._build_payload = types.MethodType(lambda self: bytes(range(55)), embedder)
    frame_ctr = 5
    embedder.frame_ctr = frame_ctr
    frame = embedder._make_frame_chips()
    assert frame.size == api.FRAME_LEN
    data_bits = api.polar_encode(bytes(range(55)), N=embedder.p.N, K=embedder.p.K)
    data_symbols = 2.0 * data_bits.astype(np.float32) - 1.0
    ctr_lo16 = np.uint16(frame_ctr & 0xFFFF)
    hdr_bits = np.unpackbits(np.array([ctr_lo16 >> 8, ctr_lo16 & 0xFF], dtype=np.uint8))
    hdr_sy = (2.0 * np.repeat(hdr_bits, 8).astype(np.float32) - 1.0) * embedder._hdr_pn_sy
    pn_full = embedder.sec.pn_bits(frame_ctr, api.PRE_L + api.HDR_L + embedder.p.N)
    pn_symbols = 2.0 * pn_full[api.PRE_L + api.HDR_L:].astype(np.float32) - 1.0
    symbols = np.concatenate((embedder._preamble_sy, hdr_sy, data_symbols * pn_symbols))
    band = api.choose_band(embedder._band_key, frame_ctr)
    b, a = api.butter_bandpass(*band, embedder.p.fs, order=4)
    zi0 = np.zeros(max(len(a), len(b)) - 1, dtype=np.float32)
    y_pre, zi1 = lfilter(b, a, embedder._preamble_sy, zi=zi0)
    y_rest, _ = lfilter(b, a, symbols[63:], zi=zi1)
    np.testing.assert_allclose(frame, np.concatenate((y_pre, y_rest)), rtol=1e-5, atol=1e-5)


def test_noise_is_not_authentic(api):
    # reference tests/test_false_positive.py:8-19
    rng = np.random.default_rng(0)
    noise = (0.05 * rng.standard_normal(3 * 48000)).astype(np.float32)
    assert api.WatermarkDetector(KEY).verify(noise, 48000) is False


def test_empty_and_wrong_key(api):
    # reference tests/test_edge_cases.py:14-20,64-71 ; tests/test_crypto.py:24-29
    assert api.WatermarkDetector(KEY).verify(np.zeros(0, np.float32), 48000) is False
    tx = api.WatermarkEmbedder(KEY)
    rng = np.random.default_rng(1)
    wm = tx.process((0.05 * rng.standard_normal(48000)).astype(np.float32))
    assert api.WatermarkDetector(b"\xBB" * 32).verify(wm, 48000) is False
    sc = api.SecureChannel(KEY)
    blob = bytearray(sc.seal(b"x" * 27))
    blob[20] ^= 1
    with pytest.raises(Exception):
        sc.open(bytes(blob))
    assert sc.open(sc.seal(b"y" * 27)) == b"y" * 27


def test_pn_sign_convention(api):
    # reference tests/test_detector.py:58-75 (payload offset corrected to 63+128 for the 1215-chip frame)
    tx = api.WatermarkEmbedder(KEY)
    tx.frame_ctr = 0
    frame = tx._make_frame_chips()
    pn = 2 * api.SecureChannel(KEY).pn_bits(0, frame.size)[191:].astype(np.float64) - 1
    assert abs(float(np.mean(frame[191:] * pn))) < 0.2


def test_quick_roundtrip_list_size_32(api):
    """The reference's own tests/test_roundtrip_quick.py builds WatermarkDetector(key, list_size=32): accepted here and
    decoded with 32 paths (wide-list kernel), same verdict, attempt lists and number of SCL decodes as the oracle run with
    the same list size (rtwm/detector.py:27,44-152)."""
    import warnings
    from oracle import detector_oracle as do
    key = bytes(range(32))
    tx = api.WatermarkEmbedder(key)
    rng = np.random.default_rng(3)
    t = np.arange(24000) / 48000.0
    speech = (0.1 * np.sin(2 * np.pi * 220 * t) + 0.01 * rng.standard_normal(t.size)).astype(np.float32)
    wm = np.concatenate([tx.process(speech[i:i + 1024]) for i in range(0, speech.size, 1024)])
    with warnings.catch_warnings():
        warnings.simplefilter("error")                      # no clamp warning at 32
        rx = api.WatermarkDetector(key, list_size=32)
    got = rx.verify(wm, 48000)
    want, det = do.verify(wm.astype(np.float32), key, list_size=32, return_details=True)
    assert got == want
    r = rx.last_result
    assert r.n_scl == det["n_scl"]
    for bi in range(4):
        assert r.attempts[bi] == [(int(a), int(b)) for a, b in det["attempts"][bi]]
