"""phi(d) = log1p(exp(-d)) of the SCL kernel (echoseal_b200/csrc/phi_impl.h) built for the host: every
operation is an IEEE add / mul / fma plus table reads, so this is bit-identical to the device code.
Checked against mpmath (exact) and glibc (what the reference's np.logaddexp uses).

Contract (phi_impl.h header): ABSOLUTE error below 2.5e-16 over the whole range — the decoder only ever adds phi to
numbers of magnitude >= phi, so that is the accuracy class of the reference's own libm composition (whose
absolute error near d = 0 is 1.2e-16) — plus phi(0) == ln 2 exactly, 0 <= phi <= ln 2, exactly 0 beyond d = 37."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LN2 = 0.693147180559945309417232121458176568


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


def test_absolute_error_vs_mpmath(lib):
    import mpmath as mp
    mp.mp.prec = 120
    rng = np.random.default_rng(0)
    d = np.concatenate([rng.uniform(0, 40, 6000), rng.uniform(0, 1, 2500), 10 ** rng.uniform(-18, 0, 1000),
                        rng.uniform(30, 70, 500), rng.uniform(600, 800, 200),
                        # interval edges of the two tables
                        np.log(2) / 64 * (np.arange(1, 400) + rng.uniform(-1e-9, 1e-9, 399))])
    fast, libm = _run(lib, d)
    ef, el = [], []
    for x, o, l in zip(d, fast, libm):
        ex = mp.log1p(mp.exp(-mp.mpf(float(x))))
        ef.append(abs(float(mp.mpf(float(o)) - ex))); el.append(abs(float(mp.mpf(float(l)) - ex)))
    ef, el = np.array(ef), np.array(el)
    print(f"phi_fast: max abs err {ef.max():.3e}, mean {ef.mean():.3e}; glibc: max {el.max():.3e}, mean {el.mean():.3e}")
    assert ef.max() < 2.5e-16 and ef.mean() < 6e-17
    assert (fast >= 0).all() and (fast <= LN2).all()


def test_edge_cases(lib):
    d = np.array([0.0, -0.0, 1e-300, 1e-17, 5e-17, 1.1e-16, 2.2e-16, 36.0, 37.0, 38.0, 64.0, 65.0, 700, 745.2, 1400, 1e6, 1e300,
                  np.inf, -3.25, 3.25])
    fast, libm = _run(lib, d)
    assert fast[0] == LN2 and fast[1] == LN2                  # x == y case of np.logaddexp: x + ln 2
    assert (np.abs(fast[2:8] - libm[2:8]) < 2.5e-16).all()
    assert (fast[8:18] == 0.0).all()                          # below half an ulp of 1: adds nothing anywhere
    assert fast[18] == fast[19]                               # phi takes |d| itself


def test_monotone_on_a_fine_grid(lib):
    """phi is decreasing; the routine may wobble by its error bound but never more."""
    d = np.linspace(0, 40, 400001)
    fast, _ = _run(lib, d)
    assert (np.diff(fast) < 3e-16).all()


def test_f_as_psi_difference_matches_the_reference_form(lib):
    """The kernel forms f = logaddexp(a,b) - logaddexp(0,a+b) (rtwm/fastpolar.py:18-23) as psi(a-b) - psi(a+b) with
    psi(x) = |x|/2 + phi(|x|).  Against mpmath on LLR-like operands the absolute error stays within a few ulp of the larger
    operand — the error class of the reference's own two-logaddexp form — and psi is even, so f(-a,-b) == f(a,b) and
    f(-a,b) == -f(a,b) hold bit for bit."""
    import mpmath as mp
    mp.mp.prec = 120
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.uniform(-12, 12, 3000), rng.uniform(-60, 60, 1500), rng.uniform(-1e3, 1e3, 500)])
    b = np.concatenate([rng.uniform(-12, 12, 3000), rng.uniform(-60, 60, 1500), rng.uniform(-1e3, 1e3, 500)])

    def psi(x):
        x = np.ascontiguousarray(x, np.float64); o = np.empty_like(x)
        lib.psi_fast_array(x.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_int(x.size))
        return o
    f = psi(a - b) - psi(a + b)
    ref_form = np.logaddexp(a, b) - np.logaddexp(0.0, a + b)
    worst = worst_ref = 0.0
    for x, y, v, r in zip(a, b, f, ref_form):
        X, Y = mp.mpf(float(x)), mp.mpf(float(y))
        ex = mp.log((mp.exp(X) + mp.exp(Y)) / (1 + mp.exp(X + Y))) if abs(x) + abs(y) < 600 else None
        if ex is None:
            continue
        ulp = np.spacing(max(abs(x), abs(y), 1.0))
        worst = max(worst, abs(float(mp.mpf(float(v)) - ex)) / ulp)
        worst_ref = max(worst_ref, abs(float(mp.mpf(float(r)) - ex)) / ulp)
    print(f"f as psi difference: max error {worst:.2f} ulp of the larger operand; numpy two-logaddexp form: {worst_ref:.2f}")
    assert worst < 4.0
    assert (psi(-(a - b)) == psi(a - b)).all()
