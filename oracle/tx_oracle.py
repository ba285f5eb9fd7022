"""CPU ORACLE for the TX path — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of rtwm/embedder.py (frame synthesis + level-controlled mix) and of the host
crypto it consumes (rtwm/crypto.py, rtwm/utils.py: HKDF, ChaCha20-Poly1305, AES-ECB PN stream,
HMAC hop).  Also used by the tests to BUILD watermarked input clips from seeds, so that no large
audio fixtures have to be committed.

Parity status: PINNED — tests/golden/make_rx_golden.py asserts, at generation time, that this
restatement reproduces the reference embedder's output sample-for-sample on every golden clip
(the checksums stored in tests/golden/rx_golden.npz are re-checked by tests/test_oracle_rx.py),
and tests/test_oracle_tx.py re-checks it against stored reference frames."""
from __future__ import annotations
import hashlib
import hmac
import struct
import numpy as np
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
from cryptography.hazmat.primitives.ciphers.aead import ChaCha20Poly1305
from cryptography.hazmat.primitives.kdf.hkdf import HKDF
from cryptography.hazmat.primitives import hashes

from . import detector_oracle as do
from . import polar_oracle as po

EPS = 1e-12
MIX_HEADROOM = 0.98


class Keys:
    """rtwm/crypto.py:14-30 + rtwm/utils.py:94 (sub-key) + rtwm/embedder.py:33 (band key = raw key)."""

    def __init__(self, key32: bytes):
        if len(key32) != 32:
            raise ValueError("master_key must be 32 bytes (256 bit)")
        okm = HKDF(algorithm=hashes.SHA256(), length=64, salt=None, info=b"EchoSeal:KDF:v1").derive(key32)
        self.aead = ChaCha20Poly1305(okm[:32])
        sub = hashlib.blake2s(okm[32:], digest_size=16, person=b"EchoSeal").digest()
        self._ecb = Cipher(algorithms.AES(sub), modes.ECB())
        self.band_key = key32

    def pn_bits(self, ctr: int, n: int) -> np.ndarray:
        """rtwm/utils.py:115-132"""
        nblk = ((n + 7) // 8 + 15) // 16
        buf = b"".join(((ctr << 64) + j).to_bytes(16, "big") for j in range(nblk))
        enc = self._ecb.encryptor()
        raw = enc.update(buf) + enc.finalize()
        return np.unpackbits(np.frombuffer(raw, np.uint8))[:n]

    def band_index(self, ctr: int) -> int:
        return hmac.new(self.band_key, struct.pack(">I", ctr), "sha256").digest()[0] % 4

    def seal(self, nonce12: bytes, plaintext: bytes) -> bytes:
        """rtwm/crypto.py:33-37 with the nonce supplied by the caller (the reference draws it from
        secrets.token_bytes)."""
        return nonce12 + self.aead.encrypt(nonce12, plaintext, b"")

    def open(self, blob: bytes) -> bytes:
        if len(blob) < 28:
            raise ValueError("ciphertext too short")
        return self.aead.decrypt(blob[:12], blob[12:], b"")


def build_payload(keys: Keys, ctr: int, session_nonce8: bytes, pad11: bytes, nonce12: bytes) -> bytes:
    """rtwm/embedder.py:153-168 with the three random fields passed in."""
    meta = b"ESAL" + ctr.to_bytes(4, "big") + session_nonce8 + pad11
    assert len(meta) == 27
    return keys.seal(nonce12, meta)


def frame_symbols(keys: Keys, ctr: int, payload55: bytes) -> np.ndarray:
    """+-1 chips before filtering: preamble | header | spread payload (rtwm/embedder.py:96-127)."""
    data_bits = po.encode(np.unpackbits(np.frombuffer(payload55, np.uint8)))
    pre = 2.0 * do.mseq_63().astype(np.float32) - 1.0
    data_sy = 2.0 * data_bits.astype(np.float32) - 1.0
    lo16 = ctr & 0xFFFF
    hdr_bits = np.unpackbits(np.array([lo16 >> 8, lo16 & 0xFF], np.uint8))
    hdr_bpsk = 2.0 * np.repeat(hdr_bits, 8).astype(np.float32) - 1.0
    hdr_pn = 2.0 * keys.pn_bits(0, 128).astype(np.float32) - 1.0
    pn_full = keys.pn_bits(ctr, 1215)
    pn_sy = 2.0 * pn_full[191:].astype(np.float32) - 1.0
    return np.concatenate((pre, hdr_bpsk * hdr_pn, data_sy * pn_sy)).astype(np.float32)


def frame_chips(keys: Keys, ctr: int, payload55: bytes, fs: int = 48000) -> np.ndarray:
    """rtwm/embedder.py:78-151: zero-state band-pass of the frame symbols in the hop band
    (preamble then the rest with the carried state == one zero-state pass), peak-normalise only if
    the peak exceeds 3.0, float32 out."""
    sym = frame_symbols(keys, ctr, payload55)
    band = do.BAND_PLAN[keys.band_index(ctr)]
    b, a = do.butter_bandpass(*band, fs)
    zi0 = np.zeros(max(len(a), len(b)) - 1)
    y_pre, zi1 = do.lfilter(b, a, sym[:63], zi=zi0)
    y_rest, _ = do.lfilter(b, a, sym[63:], zi=zi1)
    chips = np.concatenate((y_pre, y_rest))
    peak = float(np.max(np.abs(chips))) + EPS
    if peak > 3.0:
        chips = chips * (1.0 / peak)
    return chips.astype(np.float32)


def mix(x: np.ndarray, chips: np.ndarray, target_rel_db=-10.0, floor_rel_dbfs=-35.0) -> np.ndarray:
    """rtwm/embedder.py:50-75 for one process() call: x + chips*scale, scale tied to the block RMS with
    an absolute floor and a clip-headroom limiter."""
    x = x.astype(np.float32, copy=False)
    in_rms = float(np.sqrt(np.mean(x * x)) + EPS)
    chips = chips.astype(np.float32, copy=False)
    scale = max(10.0 ** (target_rel_db / 20.0) * in_rms, 10.0 ** (floor_rel_dbfs / 20.0))
    headroom = MIX_HEADROOM - float(np.max(np.abs(x)))
    if headroom < 0.0:
        headroom = 0.0
    peak = float(np.max(np.abs(chips))) + EPS
    scale = min(scale, headroom / peak) if peak > 0.0 else 0.0
    return x + chips * scale


class Embedder:
    """Stateful restatement of WatermarkEmbedder (rtwm/embedder.py:29-75) with injectable randomness:
    `rand(n)` returns n bytes and replaces secrets.token_bytes in call order
    (session nonce 8 at construction; per frame: pad 11 then AEAD nonce 12)."""

    def __init__(self, key32: bytes, rand, fs: int = 48000):
        self.k = Keys(key32)
        self.fs = fs
        self.rand = rand
        self.frame_ctr = 0
        self.buf = np.empty(0, np.float32)
        self.session_nonce = rand(8)

    def make_frame(self) -> np.ndarray:
        pad = self.rand(11)
        nonce = self.rand(12)
        payload = build_payload(self.k, self.frame_ctr, self.session_nonce, pad, nonce)
        return frame_chips(self.k, self.frame_ctr, payload, self.fs)

    def process(self, samples: np.ndarray) -> np.ndarray:
        need = samples.size
        while self.buf.size < need:
            self.buf = np.concatenate((self.buf, self.make_frame()))
            self.frame_ctr = (self.frame_ctr + 1) % (2 ** 32)
        chips = self.buf[:need]
        self.buf = self.buf[need:]
        return mix(samples, chips)


def seeded_rand(seed: int):
    rng = np.random.default_rng(seed)
    return lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()
