"""ctypes front-end of oracle/polar_oracle.c (CPU ORACLE — test infrastructure only).

`decode()` applies the reference's selection rules (rtwm/fastpolar.py:261-276, 332-359)
on top of the sorted path list the C code returns, so it can be called exactly like
`rtwm.fastpolar.PolarCode.decode(llr, validator)`.

Parity: pinned by tests/test_oracle_polar.py against tests/golden/polar_golden.npz
(outputs of the reference itself)."""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_FAST = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "_build", "liboracle.so")
    so2 = os.path.join(_HERE, "_build", "liboracle_fastphi.so")
    srcs = [os.path.join(_HERE, f) for f in ("polar_oracle.c", "dsp_oracle.c")]
    srcs += [os.path.join(_HERE, "..", "echoseal_b200", "csrc", f) for f in ("phi_impl.h", "phi_tables.h")]
    if force or not os.path.exists(so) or not os.path.exists(so2) or \
            any(os.path.getmtime(s) > min(os.path.getmtime(so), os.path.getmtime(so2)) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.es_oracle_scl_decode_batch.restype = C.c_int
        _LIB.es_oracle_crc8.restype = C.c_uint8
    return _LIB


def lib_fastphi():
    """Device-arithmetic model: the oracle decoder built with the CUDA kernel's phi() routine."""
    global _LIB_FAST
    if _LIB_FAST is None:
        build()
        _LIB_FAST = C.CDLL(os.path.join(_HERE, "_build", "liboracle_fastphi.so"))
        _LIB_FAST.es_oracle_scl_decode_batch.restype = C.c_int
    return _LIB_FAST


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def frozen_default(K: int = 448) -> np.ndarray:
    import sys
    sys.path.insert(0, os.path.join(_HERE, ".."))
    from echoseal_b200.polar_tables import frozen_mask
    return frozen_mask(1024, K).astype(np.uint8)


def scl_batch(llr: np.ndarray, L: int = 8, K: int = 448, frozen=None, skip_on_hard_crc: bool = False,
              threads: int | None = None, device_arith: bool = False, neg_mode: bool = False):
    """llr float32[rows,1024] -> dict(hard_info, hard_crc, path_info, path_metric, path_crc, npaths, stats).
    neg_mode: the detector's pairing (rtwm/detector.py:405-413): codeword 2r = +row r, codeword 2r+1 = -row r."""
    llr = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, 1024)
    ncw = llr.shape[0] * (2 if neg_mode else 1)
    fr = np.ascontiguousarray(frozen_default(K) if frozen is None else frozen, dtype=np.uint8)
    ninfo = K - 8
    out = dict(
        hard_info=np.zeros((ncw, ninfo), np.uint8), hard_crc=np.zeros(ncw, np.int32),
        path_info=np.zeros((ncw, L, ninfo), np.uint8), path_metric=np.full((ncw, L), np.inf),
        path_crc=np.zeros((ncw, L), np.int32), npaths=np.zeros(ncw, np.int32),
        stats=np.zeros((ncw, 4), np.float64),
    )
    rc = (lib_fastphi() if device_arith else lib()).es_oracle_scl_decode_batch(_p(llr), C.c_int(ncw), _p(fr), C.c_int(K), C.c_int(L),
                                          C.c_int((1 if skip_on_hard_crc else 0) | (2 if neg_mode else 0)), C.c_int(threads or (os.cpu_count() or 1)),
                                          _p(out["hard_info"]), _p(out["hard_crc"]), _p(out["path_info"]),
                                          _p(out["path_metric"]), _p(out["path_crc"]), _p(out["npaths"]),
                                          _p(out["stats"]))
    if rc:
        raise RuntimeError(f"oracle scl rc={rc}")
    return out


def select(res: dict, w: int, validator=None):
    """Reference selection (rtwm/fastpolar.py:269-276, 332-359) for codeword w of a scl_batch result.
    Returns (info_bits uint8[440], ok)."""
    def _val(bits):
        if validator is None:
            return True
        try:
            return bool(validator(np.packbits(bits).tobytes()))
        except Exception:
            return False
    if res["hard_crc"][w] and _val(res["hard_info"][w]):
        return res["hard_info"][w].copy(), True
    best_crc = None
    best_any = res["hard_info"][w].copy()
    best_any_m = np.inf
    for a in range(int(res["npaths"][w])):
        bits, m = res["path_info"][w, a], res["path_metric"][w, a]
        if res["path_crc"][w, a]:
            if _val(bits):
                return bits.copy(), True
            if best_crc is None:
                best_crc = bits.copy()
        elif m < best_any_m:
            best_any_m, best_any = m, bits.copy()
    if best_crc is not None:
        return best_crc, False
    return best_any, False


def decode(llr: np.ndarray, validator=None, L: int = 8, K: int = 448):
    """Drop-in for PolarCode(1024,K,L,8).decode(llr, validator) -> (info_bits, ok)."""
    llr = np.asarray(llr)
    if llr.ndim != 1 or llr.size != 1024:
        raise ValueError("llr must be 1D length 1024")
    res = scl_batch(llr.astype(np.float32)[None], L=L, K=K)
    return select(res, 0, validator)


def encode(info_bits: np.ndarray, K: int = 448) -> np.ndarray:
    fr = frozen_default(K)
    info = np.ascontiguousarray(info_bits, dtype=np.uint8)
    cw = np.zeros(1024, np.uint8)
    lib().es_oracle_polar_encode(_p(fr), C.c_int(K), _p(info), _p(cw))
    return cw
