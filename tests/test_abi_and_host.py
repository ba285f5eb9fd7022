"""CPU-side checks: the C-ABI library loads and exports every symbol include/echoseal_b200.h declares
(no compute calls without a GPU); host-side logic (crypto, hop schedule, constants) matches the
reference-generated golden vectors; the product never imports the oracle."""
import ctypes
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "rx_golden.npz"))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from echoseal_b200 import _native
    return _native.lib()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "echoseal_b200.h")).read()
    names = sorted(set(re.findall(r"\b(es_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_compute_entry_points_fail_loudly_without_setup(lib):
    # no GPU work is launched: the call is rejected before any kernel because no code was uploaded
    import echoseal_b200._native as N
    rc = lib.es_scl_list(None, None, 1, 0, 99, None, ctypes.c_size_t(0), None, None, None, None, None)
    assert rc != 0 and len(lib.es_last_error()) > 0
    with pytest.raises(N.NativeError):
        N.check(rc, "es_scl_list")


def test_product_has_no_oracle_or_cpu_fallback():
    pkg = os.path.join(ROOT, "echoseal_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_host_crypto_matches_reference_vectors():
    from echoseal_b200.crypto import SecureChannel
    from echoseal_b200.utils import choose_band, choose_band_index, BAND_PLAN, hop_table
    for key_b in (0xAA, 0x01):
        key = bytes([key_b]) * 32
        sc = SecureChannel(key)
        pn = np.stack([np.packbits(sc.pn_bits(c, 1215)) for c in (0, 1, 255, 1024, 2 ** 31 + 5)])
        assert (pn == G[f"crypto/{key_b:02x}/pn"]).all()
        batch = sc.pn_bytes_batch([0, 1, 255, 1024, 2 ** 31 + 5], 1215).copy()
        batch[:, -1] &= 0xFE
        assert (batch == G[f"crypto/{key_b:02x}/pn"]).all()
        assert [choose_band_index(key, c) for c in range(512)] == list(G[f"crypto/{key_b:02x}/hop"])
        assert (hop_table(key, 0, 512) == G[f"crypto/{key_b:02x}/hop"]).all()
        assert choose_band(key, 3) == BAND_PLAN[G[f"crypto/{key_b:02x}/hop"][3]]
        blob = G[f"crypto/{key_b:02x}/seal"].tobytes()
        assert sc.open(blob) == b"ESAL" + bytes(23)
        assert len(sc.seal(b"x" * 27)) == 55
    with pytest.raises(ValueError):
        SecureChannel(b"short")


def test_constants_and_tables():
    from echoseal_b200.utils import mseq_63, db_to_lin, butter_bandpass, BAND_PLAN
    from echoseal_b200.polar_tables import frozen_mask, data_positions
    bits = "".join(str(int(b)) for b in mseq_63())
    assert bits == "100000100001100010100111101000111001001011011101100110101011111"
    assert abs(db_to_lin(-20.0) - 0.1) < 1e-15
    for band in BAND_PLAN:
        b, a = butter_bandpass(*band, 48000)
        assert b.size == a.size == 9 and (b[1::2] == 0).all()
    fr = frozen_mask()
    PG = np.load(os.path.join(ROOT, "tests", "golden", "polar_golden.npz"))
    assert (np.packbits(fr.astype(np.uint8)) == PG["frozen"]).all()
    assert data_positions().size == 448
    with pytest.raises(ValueError):
        frozen_mask(512, 100)
