"""Native host feeder (echoseal_b200/csrc/host_feeder.cpp) against the Python host crypto
(echoseal_b200/crypto.py, utils.py — themselves pinned to the reference by tests/test_host_logic.py)
and against the candidate enumeration of the oracle.  CPU-only: no kernel is launched."""
import os
import numpy as np
import pytest

from echoseal_b200 import host_feeder
from echoseal_b200.crypto import SecureChannel
from echoseal_b200.utils import choose_band_index

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "rx_golden.npz"))
KEYS = [bytes([0xAA]) * 32, bytes([0x01]) * 32, bytes(range(32))]


def test_key_derivation_pn_hop_match_reference_vectors():
    bank = host_feeder.KeyBank(KEYS)
    for i, kb in enumerate((0xAA, 0x01)):
        pn = bank.pn(i, [0, 1, 255, 1024, 2 ** 31 + 5])
        pn[:, -1] &= 0xFE            # bit 1215 is past the frame (np.packbits pads it with 0)
        assert (pn == G[f"crypto/{kb:02x}/pn"]).all()
        assert (bank.hop(i, 0, 512) == G[f"crypto/{kb:02x}/hop"]).all()
    sc = SecureChannel(KEYS[2])
    assert (bank.hdr_pn()[2] == np.packbits(sc.pn_bits(0, 128))).all()
    assert [int(v) for v in bank.hop(2, 100, 140)] == [choose_band_index(KEYS[2], c) for c in range(100, 140)]


def test_bad_key_length():
    with pytest.raises(ValueError):
        host_feeder.KeyBank([b"abc"])


def test_tx_prepare_matches_python_crypto():
    bank = host_feeder.KeyBank(KEYS)
    rng = np.random.default_rng(0)
    F = 200
    kidx = rng.integers(0, 3, F).astype(np.int32)
    ctr = rng.integers(0, 2 ** 32, F, dtype=np.uint64).astype(np.uint32)
    sn = rng.integers(0, 256, (F, 8), dtype=np.uint8)
    rnd = rng.integers(0, 256, (F, 23), dtype=np.uint8)
    out = bank.tx_prepare(kidx, ctr, sn, rnd)
    scs = [SecureChannel(k) for k in KEYS]
    for i in range(F):
        sc = scs[kidx[i]]
        meta = b"ESAL" + int(ctr[i]).to_bytes(4, "big") + sn[i].tobytes() + rnd[i, :11].tobytes()
        nonce = rnd[i, 11:].tobytes()
        assert out["payload"][i].tobytes() == nonce + sc._aead.encrypt(nonce, meta, b"")
        assert sc.open(out["payload"][i].tobytes()) == meta
        pni = out["pn"][i].copy(); pni[-1] &= 0xFE
        assert (pni == np.packbits(sc.pn_bits(int(ctr[i]), 1215))).all()
        assert out["band"][i] == choose_band_index(KEYS[kidx[i]], int(ctr[i]))
        assert out["ctr_lo16"][i] == int(ctr[i]) & 0xFFFF


def test_rx_enumerate_matches_oracle_rule_and_budget():
    from oracle import detector_oracle as do
    bank = host_feeder.KeyBank(KEYS)
    rng = np.random.default_rng(1)
    nb, n = 6, 144000
    kidx = np.array([0, 1, 2, 0, 1, 2], np.int32)
    peaks = np.full((nb, 4, 25), -1, np.int32); npk = np.zeros((nb, 4), np.int32)
    hdr = np.zeros((nb, 4, 25, 4), np.float32)
    for ci in range(nb):
        for bi in range(4):
            m = int(rng.integers(0, 26))
            npk[ci, bi] = m
            peaks[ci, bi, :m] = np.sort(rng.integers(0, n - 62, m))
            hdr[ci, bi, :m, 0] = rng.integers(0, 2, m)
            hdr[ci, bi, :m, 1] = rng.integers(0, 400, m)       # small values so header-gated hits happen
    e = bank.rx_enumerate(kidx, n, peaks, npk, hdr)
    pos = 0
    for ci in range(nb):
        key = KEYS[kidx[ci]]
        sc = SecureChannel(key)
        assert e["item_offset"][ci] == pos
        for bi in range(4):
            tried, stop, want = 0, False, []
            for slot in range(npk[ci, bi]):
                start = int(peaks[ci, bi, slot])
                if start + 1215 > n:
                    continue
                for c in do.candidate_counters(start, hdr[ci, bi, slot, 0] > 0.5, int(hdr[ci, bi, slot, 1]), bi,
                                               lambda c_: choose_band_index(key, c_)):
                    want.append(((ci * 4 + bi) * 25 + slot, c)); tried += 1
                    if tried >= 400:
                        stop = True; break
                if stop:
                    break
            assert e["band_count"][ci, bi] == len(want)
            got = list(zip(e["item_peak"][pos:pos + len(want)].tolist(), e["item_ctr"][pos:pos + len(want)].tolist()))
            assert got == want
            assert (e["item_clip"][pos:pos + len(want)] == ci).all()
            for j, (_, c) in enumerate(want[:3]):
                pnj = e["pn"][pos + j].copy(); pnj[-1] &= 0xFE
                assert (pnj == np.packbits(sc.pn_bits(c, 1215))).all()
            pos += len(want)
    assert e["item_offset"][nb] == pos == e["item_peak"].size
    assert e["band_count"].max() <= 400


def test_rx_validate_order_and_nonce_latch():
    bank = host_feeder.KeyBank(KEYS)
    sc = SecureChannel(KEYS[0])
    hop0 = choose_band_index(KEYS[0], 0)
    other = (hop0 + 1) % 4
    # one clip, two bands with 2 attempts each
    bc = np.zeros((1, 4), np.int32); bc[0, hop0] = 2; bc[0, other] = 2
    off = np.array([0, 4], np.int64)
    # items are stored in BAND_PLAN order
    bands_sorted = sorted([hop0, other])
    ctrs = {hop0: [7, 9], other: [11, 13]}
    item_ctr = np.array(ctrs[bands_sorted[0]] + ctrs[bands_sorted[1]], np.uint32)
    enum = dict(band_count=bc, item_offset=off, item_ctr=item_ctr)
    def blob(ctr, nonce8):
        return np.frombuffer(sc.seal(b"ESAL" + ctr.to_bytes(4, "big") + nonce8 + bytes(11)), np.uint8)
    def item_index(band, a):
        return (0 if band == bands_sorted[0] else 2) + a
    n1, n2 = b"AAAAAAAA", b"BBBBBBBB"
    # valid candidates: other band attempt 0 (nonce n2), hop0 band attempt 1 variant 2 slot 3 (nonce n1);
    # plus garbage and a wrong-counter blob earlier in the order
    hits = [
        (4 * item_index(hop0, 0) + 0, 0, np.zeros(55, np.uint8)),                 # garbage
        (4 * item_index(hop0, 0) + 1, 2, blob(9, n1)),                             # valid tag, wrong counter for ctr=7
        (4 * item_index(hop0, 1) + 2, 3, blob(9, n1)),                             # the winner
        (4 * item_index(other, 0) + 0, 0, blob(11, n2)),
    ]
    hits.sort(key=lambda t: (t[0], t[1]))
    cw = np.array([h[0] for h in hits], np.int64); sl = np.array([h[1] for h in hits], np.int32)
    pl = np.stack([h[2] for h in hits])
    ns = np.zeros((1, 9), np.uint8)
    v, pt = bank.rx_validate(np.array([0], np.int32), enum, cw, sl, pl, ns)
    assert v[0] == 1 and pt[0, :8].tobytes() == b"ESAL" + (9).to_bytes(4, "big")
    assert ns[0, 0] == 1 and ns[0, 1:].tobytes() == n1
    # latched to another nonce: hop0-band candidate is rejected, the other band's (n2) accepted
    ns = np.zeros((1, 9), np.uint8); ns[0, 0] = 1; ns[0, 1:] = np.frombuffer(n2, np.uint8)
    v, pt = bank.rx_validate(np.array([0], np.int32), enum, cw, sl, pl, ns)
    assert v[0] == 1 and pt[0, 4:8].tobytes() == (11).to_bytes(4, "big")
    # latched to a third nonce: nothing verifies
    ns = np.zeros((1, 9), np.uint8); ns[0, 0] = 1; ns[0, 1:] = 7
    v, _ = bank.rx_validate(np.array([0], np.int32), enum, cw, sl, pl, ns)
    assert v[0] == 0
