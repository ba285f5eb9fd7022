"""Host-side keying / AEAD — mirrors rtwm/crypto.py (SecureChannel).  Stays on the host by
design: its outputs (PN bits) are kernel inputs, its inputs (candidate payloads) are kernel outputs."""
from __future__ import annotations
import secrets
import numpy as np
from cryptography.hazmat.primitives.ciphers.aead import ChaCha20Poly1305
from cryptography.hazmat.primitives.kdf.hkdf import HKDF
from cryptography.hazmat.primitives import hashes

from .utils import StreamPRNG, pn_bits as _pn_bits


class SecureChannel:
    """HKDF-SHA256(master, info="EchoSeal:KDF:v1") -> aead_key | prng_key (rtwm/crypto.py:14-30)."""

    def __init__(self, master_key: bytes) -> None:
        if len(master_key) != 32:
            raise ValueError("master_key must be 32 bytes (256 bit)")
        okm = HKDF(algorithm=hashes.SHA256(), length=64, salt=None, info=b"EchoSeal:KDF:v1").derive(master_key)
        self._aead = ChaCha20Poly1305(okm[:32])
        self._prng = StreamPRNG(okm[32:])

    def seal(self, plaintext: bytes) -> bytes:
        """nonce(12) || ct || tag(16) (rtwm/crypto.py:33-37)."""
        nonce = secrets.token_bytes(12)
        return nonce + self._aead.encrypt(nonce, plaintext, b"")

    def open(self, blob: bytes) -> bytes:
        """raises on failure (rtwm/crypto.py:39-43)."""
        if len(blob) < 12 + 16:
            raise ValueError("ciphertext too short")
        return self._aead.decrypt(blob[:12], blob[12:], b"")

    def pn_bits(self, frame_ctr: int, n_bits: int) -> np.ndarray:
        return _pn_bits(self._prng, frame_ctr, n_bits)

    def pn_bytes_batch(self, ctrs, n_bits: int) -> np.ndarray:
        """Packed PN (MSB-first per byte, as np.unpackbits reads them) for many counters at once."""
        return self._prng.bytes_batch(ctrs, (n_bits + 7) // 8)
