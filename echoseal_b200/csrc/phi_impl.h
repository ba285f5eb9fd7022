// echoseal_b200/csrc/phi_impl.h — branch-free IEEE-double  phi(d) = log1p(exp(-|d|)).
//
// One source for the CUDA kernel (scl.cu), for the device-arithmetic model of the decoder
// (oracle/polar_oracle.c with -DORACLE_PHI_FAST) and for the accuracy test (tests/test_phi_accuracy.py
// builds it with gcc: every operation is an IEEE add/mul/fma on doubles plus table reads, so the host
// build is bit-identical to the device build).
//
// The reference evaluates this quantity as np.log1p(np.exp(-d)) inside np.logaddexp and _metric_penalty
// (rtwm/fastpolar.py:18-40) and only ever ADDS it to numbers of magnitude >= |phi|: to max(a,b) and
// max(0,a+b) in f = logaddexp(a,b) - logaddexp(0,a+b), whose difference cancels to an absolute error of a few
// 1e-16 whatever libm does, and to path metrics.  What the decoder needs from phi is therefore ABSOLUTE accuracy
// at the 1e-16 level, not relative accuracy of its tiny values.  This routine delivers |error| < 2.5e-16 over the
// whole range (measured max 1.9e-16, mean 0.5e-16, tests/test_phi_accuracy.py; glibc: 0.9e-16 near d = 0) in 18
// FP64 instructions and two table reads, 0 <= phi <= ln 2, phi(0) == ln 2 exactly, exactly 0 beyond d = 37:
//   t = exp(-d):  d clamped to 64; k = round(-d*64/ln2), r = -d - k*ln2/64, |r| <= ln2/128,
//                 t = 2^(k/64) * (1 + r + r^2/2 + .. + r^5/120) with a 64-entry table of 2^(j/64); the power of
//                 two goes straight into the exponent field (k/64 >= -93, never subnormal)
//   log1p(t):     u = 1 + t; i = top 9 mantissa bits of u; rr = u*invc[i] - 1 (fma), |rr| <= 2^-9;
//                 log u = logc[i] + rr - rr^2/2 + rr^3/3 - rr^4/4 + rr^5/5
//                 interval 0 is centred on 1 (invc = 1, logc = 0): a tiny t gives t back, never a negative number.
#pragma once
#include <stdint.h>
#include <string.h>

#ifndef PHI_FN
#ifdef __CUDACC__
#define PHI_FN __device__ __forceinline__
#else
#define PHI_FN static inline
#endif
#endif

#ifdef __CUDACC__
// tables live in shared memory; `tab` is a 32-bit shared-window address
typedef uint32_t phi_tab_t;
__device__ __forceinline__ double phi_lds(uint32_t addr)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
#define PHI_LD(tab, idx) phi_lds((tab) + 8u * (uint32_t)(idx))
__device__ __forceinline__ void phi_lds2(uint32_t addr, double& a, double& b)
{
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
// pair at doubles [2*idx, 2*idx+1] counted from double offset `off` (off even -> 16-byte aligned)
#define PHI_LD2(tab, off, idx, a, b) phi_lds2((tab) + 8u * (uint32_t)(off) + 16u * (uint32_t)(idx), a, b)
#define PHI_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PHI_HI(x) ((uint32_t)__double2hiint(x))
#define PHI_LO(x) ((uint32_t)__double2loint(x))
#define PHI_HILO(hi, lo) __hiloint2double((int)(hi), (int)(lo))
// hi + ((k div 64) << 20) as shift + multiply-add (two instructions; the compiler's own form is shift, mask, add)
__device__ __forceinline__ uint32_t phi_expadd(uint32_t hi, int32_t k)
{
    uint32_t r;
    asm("{\n.reg .s32 t;\nshr.s32 t, %1, 6;\nmad.lo.s32 %0, t, 1048576, %2;\n}" : "=r"(r) : "r"(k), "r"(hi));
    return r;
}
#define PHI_EXPADD(hi, k) phi_expadd((hi), (k))
#else
#include <math.h>
typedef const double* phi_tab_t;
#define PHI_LD(tab, idx) ((tab)[(idx)])
#define PHI_LD2(tab, off, idx, a, b) do { (a) = (tab)[(off) + 2 * (idx)]; (b) = (tab)[(off) + 2 * (idx) + 1]; } while (0)
#define PHI_FMA(a, b, c) fma((a), (b), (c))
static inline uint64_t PHI_D2U(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double PHI_U2D(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define PHI_HI(x) ((uint32_t)(PHI_D2U(x) >> 32))
#define PHI_LO(x) ((uint32_t)PHI_D2U(x))
#define PHI_HILO(hi, lo) PHI_U2D(((uint64_t)(uint32_t)(hi) << 32) | (uint32_t)(lo))
#define PHI_EXPADD(hi, k) ((uint32_t)(hi) + (uint32_t)(((k) >> 6) * 0x100000))
#endif

// scalar coefficients.  On the device they are read from __constant__ memory so that they are free
// instruction operands (c[bank][offset]) instead of immediates re-materialised inside hot loops.
#define PHI_NK 8
#define PHI_K_VALUES { \
    92.332482616893656768 /* 0: 64/ln2, exact bits set below */, 0.0 /* 1: ln2/64 hi */, 0.0 /* 2: ln2/64 lo */, \
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.2, 1.0 / 3.0 }
#ifdef __CUDACC__
__constant__ double c_phi_k[PHI_NK];
#define PHI_K(i) c_phi_k[i]
#else
static double g_phi_k[PHI_NK];
#define PHI_K(i) g_phi_k[i]
#endif
static inline void phi_fill_k(double* k)
{
    const double v[PHI_NK] = PHI_K_VALUES;
    for (int i = 0; i < PHI_NK; ++i) k[i] = v[i];
    const uint64_t b0 = PHI_INVLN2N_BITS, b1 = PHI_LN2HIN_BITS, b2 = PHI_LN2LON_BITS;
    memcpy(&k[0], &b0, 8); memcpy(&k[1], &b1, 8); memcpy(&k[2], &b2, 8);
}

// tables (doubles): [0,64) 2^(j/64); then 513 pairs (invc, logc) (entry 512 serves u == 2.0), read with one
// 16-byte load each.
#define PHI_OFF_EXP 0
#define PHI_OFF_LOGP 64
#define PHI_TAB_DOUBLES (64 + 2 * (PHI_NL + 1))   /* 1090, even */

PHI_FN double phi_fast(double d, phi_tab_t tab)
{
    const double SHIFT = 6755399441055744.0;                 // 1.5 * 2^52
    // ---- t = exp(-|d|); |d| clamped to [64, 64 + 2^-20) on the high word (phi(64) < 2e-28: zero either way)
    uint32_t dhi = PHI_HI(d) & 0x7fffffffu;
    dhi = dhi < 0x40500000u ? dhi : 0x40500000u;             // 0x40500000_00000000 = 64.0
    const double dd = PHI_HILO(dhi, PHI_LO(d));
    double kd = PHI_FMA(-dd, PHI_K(0), SHIFT);               // round(-d * 64/ln2) in the low word
    const int32_t ki = (int32_t)PHI_LO(kd);
    kd -= SHIFT;
    double r = PHI_FMA(kd, -PHI_K(1), -dd);
    r = PHI_FMA(kd, -PHI_K(2), r);
    double q = PHI_FMA(r, PHI_K(3), PHI_K(4));
    q = PHI_FMA(r, q, PHI_K(5));
    q = PHI_FMA(r, q, 0.5);
    const double p = PHI_FMA(r * r, q, r);                    // exp(r) - 1
    const double th0 = PHI_LD(tab, PHI_OFF_EXP + (ki & 63));
    // th = 2^(j/64) * 2^(k div 64): (k div 64) << 20 goes into the exponent field of the table value (>= -93: never
    // subnormal, so the scaling is exact and commutes with the fma below), off the critical path of the polynomial
    const double th = PHI_HILO(PHI_EXPADD(PHI_HI(th0), ki), PHI_LO(th0));
    const double t = PHI_FMA(th, p, th);                      // exp(-d)
    // ---- log1p(t), 0 <= t <= 1
    const double u = 1.0 + t;
    const int i = (int)(PHI_HI(u) >> 11) - 0x7fe00;           // 0..511, 512 iff u == 2.0
    double ic, lc;
    PHI_LD2(tab, PHI_OFF_LOGP, i, ic, lc);
    const double rr = PHI_FMA(u, ic, -1.0);
    double w = PHI_FMA(rr, PHI_K(6), -0.25);
    w = PHI_FMA(rr, w, PHI_K(7));
    w = PHI_FMA(rr, w, -0.5);
    return lc + PHI_FMA(rr * rr, w, rr);
}

// psi(x) = |x|/2 + phi(|x|) = ln(2 cosh(x/2)).  The decoder's f = logaddexp(a,b) - logaddexp(0,a+b) equals
// psi(a-b) - psi(a+b) (max(a,b) - max(0,a+b) = (|a-b| - |a+b|)/2), so no max/select terms travel with phi.
PHI_FN double psi_fast(double x, phi_tab_t tab)
{
    const double ax = PHI_HILO(PHI_HI(x) & 0x7fffffffu, PHI_LO(x));
    return PHI_FMA(ax, 0.5, phi_fast(x, tab));
}

// fill `tab` (PHI_TAB_DOUBLES doubles) from the generated bit patterns in phi_tables.h
#ifdef PHI_WANT_FILL
static void phi_fill_table(double* tab)
{
    for (int j = 0; j < 64; ++j) memcpy(&tab[PHI_OFF_EXP + j], &PHI_EXP[j], 8);
    for (int i = 0; i <= PHI_NL; ++i) {
        memcpy(&tab[PHI_OFF_LOGP + 2 * i], &PHI_INVC[i], 8);
        memcpy(&tab[PHI_OFF_LOGP + 2 * i + 1], &PHI_LOGC[i], 8);
    }
#ifndef __CUDACC__
    phi_fill_k(g_phi_k);
#endif
}
#endif
