"""ctypes front-end of the native host feeder (echoseal_b200/csrc/host_feeder.cpp): batched,
multi-threaded key derivation, HMAC hop tables, AES-ECB PN bits, candidate-counter enumeration with
the 400-try budget, ChaCha20-Poly1305 validation of CRC-passing candidates, and TX frame inputs.
Host-only: these are the inputs / outputs of the GPU kernels (BASELINE.json north_star)."""
from __future__ import annotations
import ctypes as C
import numpy as np

from . import _native as N

_configured = False


def _lib():
    global _configured
    L = N.lib()
    if not _configured:
        L.es_host_keys_new.restype = C.c_void_p
        L.es_host_keys_free.argtypes = [C.c_void_p]
        L.es_host_rx_enumerate.restype = C.c_int64
        _configured = True
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class KeyBank:
    """Derived material for a set of 32-byte master keys (rtwm/crypto.py:14-30, rtwm/utils.py:94)."""

    def __init__(self, keys: list[bytes], nthreads: int = 0):
        for k in keys:
            if len(k) != 32:
                raise ValueError("master_key must be 32 bytes (256 bit)")
        self.n = len(keys)
        buf = np.frombuffer(b"".join(keys), np.uint8) if keys else np.zeros(0, np.uint8)
        self._h = C.c_void_p(_lib().es_host_keys_new(_p(np.ascontiguousarray(buf)), C.c_int(self.n), C.c_int(nthreads)))
        self.threads = int(_lib().es_host_threads(self._h))

    def __del__(self):
        try:
            if self._h:
                _lib().es_host_keys_free(self._h)
                self._h = None
        except Exception:
            pass

    def hdr_pn(self, key_idx: np.ndarray | None = None) -> np.ndarray:
        n = self.n if key_idx is None else int(key_idx.size)
        out = np.empty((n, 16), np.uint8)
        ki = None if key_idx is None else np.ascontiguousarray(key_idx, np.int32)
        if _lib().es_host_hdr_pn(self._h, _p(ki), C.c_int(n), _p(out)) != 0:
            raise IndexError("key index out of range")
        return out

    def hop(self, key: int, lo: int, hi: int) -> np.ndarray:
        out = np.empty(hi - lo, np.uint8)
        if _lib().es_host_hop(self._h, C.c_int(key), C.c_uint32(lo), C.c_uint32(hi), _p(out)) != 0:
            raise IndexError("bad hop request")
        return out

    def pn(self, key: int, ctrs) -> np.ndarray:
        ctrs = np.ascontiguousarray(ctrs, np.uint64)
        out = np.empty((ctrs.size, 152), np.uint8)
        if _lib().es_host_pn(self._h, C.c_int(key), _p(ctrs), C.c_int(ctrs.size), _p(out)) != 0:
            raise IndexError("bad key index")
        return out

    # ---------------------------------------------------------------- RX
    def rx_enumerate(self, key_idx: np.ndarray, n_samples: int, peaks: np.ndarray, npeaks: np.ndarray, hdr: np.ndarray):
        """-> dict(band_count i32[nb,4], item_offset i64[nb+1], item_peak i32[I], item_ctr u32[I],
        item_clip i32[I], pn u8[I,152]) in clip-major / band-major / attempt order."""
        nb = int(key_idx.size)
        key_idx = np.ascontiguousarray(key_idx, np.int32)
        peaks = np.ascontiguousarray(peaks, np.int32); npeaks = np.ascontiguousarray(npeaks, np.int32)
        hdr = np.ascontiguousarray(hdr, np.float32)
        bc = np.zeros((nb, 4), np.int32)
        off = np.zeros(nb + 1, np.int64)
        L = _lib()
        I = int(L.es_host_rx_enumerate(self._h, _p(key_idx), C.c_int(nb), C.c_int(n_samples), _p(peaks), _p(npeaks),
                                       _p(hdr), _p(bc), _p(off), None, None, None, None))
        if I < 0:
            raise RuntimeError("es_host_rx_enumerate failed")
        ip = np.empty(I, np.int32); ic = np.empty(I, np.uint32); icl = np.empty(I, np.int32)
        pn = np.empty((I, 152), np.uint8)
        if I:
            I2 = int(L.es_host_rx_enumerate(self._h, _p(key_idx), C.c_int(nb), C.c_int(n_samples), _p(peaks), _p(npeaks),
                                            _p(hdr), _p(bc), _p(off), _p(ip), _p(ic), _p(icl), _p(pn)))
            if I2 != I:
                raise RuntimeError("es_host_rx_enumerate: pass mismatch")
        return dict(band_count=bc, item_offset=off, item_peak=ip, item_ctr=ic, item_clip=icl, pn=pn)

    def rx_validate(self, key_idx: np.ndarray, enum: dict, hit_cw: np.ndarray, hit_slot: np.ndarray,
                    hit_payload: np.ndarray, nonce_state: np.ndarray):
        """-> (verdict u8[nb], plaintext u8[nb,27]); nonce_state u8[nb,9] is updated in place."""
        nb = int(key_idx.size)
        key_idx = np.ascontiguousarray(key_idx, np.int32)
        hit_cw = np.ascontiguousarray(hit_cw, np.int64); hit_slot = np.ascontiguousarray(hit_slot, np.int32)
        hit_payload = np.ascontiguousarray(hit_payload, np.uint8)
        verdict = np.zeros(nb, np.uint8); pt = np.zeros((nb, 27), np.uint8)
        assert nonce_state.dtype == np.uint8 and nonce_state.shape == (nb, 9) and nonce_state.flags.c_contiguous
        rc = _lib().es_host_rx_validate(self._h, _p(key_idx), C.c_int(nb), _p(enum["band_count"]), _p(enum["item_offset"]),
                                        _p(enum["item_ctr"]), _p(hit_cw), _p(hit_slot), _p(hit_payload),
                                        C.c_int64(hit_cw.size), _p(nonce_state), _p(verdict), _p(pt))
        if rc != 0:
            raise RuntimeError("es_host_rx_validate failed")
        return verdict, pt

    # ---------------------------------------------------------------- TX
    def tx_prepare(self, key_idx, ctr, session_nonce, rnd):
        """key_idx i32[F], ctr u32[F], session_nonce u8[F,8], rnd u8[F,23] (pad 11 + AEAD nonce 12) ->
        dict(payload u8[F,55], pn u8[F,152], hdr_pn u8[F,16], band i32[F], ctr_lo16 i32[F])."""
        key_idx = np.ascontiguousarray(key_idx, np.int32); ctr = np.ascontiguousarray(ctr, np.uint32)
        F = int(ctr.size)
        session_nonce = np.ascontiguousarray(session_nonce, np.uint8).reshape(F, 8)
        rnd = np.ascontiguousarray(rnd, np.uint8).reshape(F, 23)
        out = dict(payload=np.empty((F, 55), np.uint8), pn=np.empty((F, 152), np.uint8),
                   hdr_pn=np.empty((F, 16), np.uint8), band=np.empty(F, np.int32), ctr_lo16=np.empty(F, np.int32))
        rc = _lib().es_host_tx_prepare(self._h, _p(key_idx), _p(ctr), _p(session_nonce), _p(rnd), C.c_int(F),
                                       _p(out["payload"]), _p(out["pn"]), _p(out["hdr_pn"]), _p(out["band"]),
                                       _p(out["ctr_lo16"]))
        if rc != 0:
            raise RuntimeError("es_host_tx_prepare failed")
        return out
