"""phi(d) = log1p(exp(-d)) of the SCL kernel (echoseal_b200/csrc/phi_impl.h) built for the host: every
operation is an IEEE add / mul / fma plus table reads, so this is bit-identical to the device code.
Checked against mpmath (exact) and glibc (what the reference's np.logaddexp uses)."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "tools", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libphi.so")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tools", "phi_model.c"), "-lm"])
    return C.CDLL(so)


def _run(lib, d):
    d = np.ascontiguousarray(d, np.float64)
    a = np.empty_like(d); b = np.empty_like(d)
    lib.phi_fast_array(d.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), C.c_int(d.size))
    lib.phi_libm_array(d.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_int(d.size))
    return a, b


def test_ulp_error_vs_mpmath(lib):
    import mpmath as mp
    mp.mp.prec = 120
    rng = np.random.default_rng(0)
    d = np.concatenate([rng.uniform(0, 40, 6000), rng.uniform(0, 1, 1500), 10 ** rng.uniform(-18, 0, 1000),
                        rng.uniform(30, 60, 500), rng.uniform(600, 800, 500)])
    fast, libm = _run(lib, d)
    ef, el = [], []
    for x, o, l in zip(d, fast, libm):
        ex = mp.log1p(mp.exp(-mp.mpf(float(x))))
        u = float(np.spacing(abs(float(ex)))) if ex != 0 else 5e-324
        ef.append(abs(float((mp.mpf(float(o)) - ex) / u))); el.append(abs(float((mp.mpf(float(l)) - ex) / u)))
    ef, el = np.array(ef), np.array(el)
    print(f"phi_fast: max {ef.max():.3f} ulp, mean {ef.mean():.3f}; glibc: max {el.max():.3f}, mean {el.mean():.3f}; "
          f"bit-equal to glibc on {np.mean(fast == libm):.4f}")
    assert ef.max() < 1.5 and ef.mean() < 0.35
    assert np.mean(fast == libm) > 0.95


def test_edge_cases_equal_glibc(lib):
    d = np.array([0.0, 1e-300, 1e-17, 5e-17, 1.1e-16, 2.2e-16, 3e-16, 700, 708.4, 744, 745, 745.13, 745.2, 746,
                  1000, 1399, 1400, 1401, 1e6, 1e300, np.inf])
    fast, libm = _run(lib, d)
    assert (fast == libm).all()
    assert fast[0] == 0.693147180559945309417232121458176568
