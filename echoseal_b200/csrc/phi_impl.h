// echoseal_b200/csrc/phi_impl.h — branch-free IEEE-double  phi(d) = log1p(exp(-d)),  d >= 0.
//
// One source for the CUDA kernel (scl.cu) and for the host-side accuracy test
// (tests/test_phi_accuracy.py builds it with gcc: every operation is an IEEE add/mul/fma on
// doubles plus table reads, so the host build is bit-identical to the device build).
//
// The reference computes this quantity as np.log1p(np.exp(-d)) inside np.logaddexp and
// _metric_penalty (rtwm/fastpolar.py:18-40) with glibc / numpy-SIMD libm, i.e. to ~1 ulp.
// This routine is accurate to < 1 ulp as well (measured max 0.8 ulp vs mpmath) at about a
// quarter of the instruction count of exp()+log1p() and with no divergent branches:
//   t = exp(-d):  k = round(-d*64/ln2), r = -d - k*ln2/64, t = 2^(k/64) * (1 + r + r^2/2 + .. + r^5/120)
//                 with a 64-entry hi/lo table of 2^(j/64); subnormal results are rounded once.
//   log1p(t):     u = 1+t, c = t-(u-1) (exact); i = top 8 mantissa bits of u; r = u*invc[i]-1 (fma);
//                 log u = logc[i] + r - r^2/2 + .. + r^7/7 ;  log1p(t) = log u + c*invc[i]*(1-r)
//                 interval 0 is centred on 1 (invc=1, logc=0) so tiny t keeps full relative accuracy.
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
#define PHI_D2U(x) ((uint64_t)__double_as_longlong(x))
#define PHI_U2D(x) __longlong_as_double((long long)(x))
#define PHI_HI(x) ((uint32_t)__double2hiint(x))
#define PHI_LO(x) ((uint32_t)__double2loint(x))
#define PHI_HILO(hi, lo) __hiloint2double((int)(hi), (int)(lo))
#else
#include <math.h>
#include <string.h>
typedef const double* phi_tab_t;
#define PHI_LD(tab, idx) ((tab)[(idx)])
#define PHI_LD2(tab, off, idx, a, b) do { (a) = (tab)[(off) + 2 * (idx)]; (b) = (tab)[(off) + 2 * (idx) + 1]; } while (0)
#define PHI_FMA(a, b, c) fma((a), (b), (c))
static inline uint64_t PHI_D2U(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double PHI_U2D(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define PHI_HI(x) ((uint32_t)(PHI_D2U(x) >> 32))
#define PHI_LO(x) ((uint32_t)PHI_D2U(x))
#define PHI_HILO(hi, lo) PHI_U2D(((uint64_t)(uint32_t)(hi) << 32) | (uint32_t)(lo))
#endif

// scalar coefficients.  On the device they are read from __constant__ memory so that they are free
// instruction operands (c[bank][offset]) instead of immediates re-materialised inside hot loops.
#define PHI_NK 12
#define PHI_K_VALUES { \
    92.332482616893656768 /* 0: 64/ln2, exact bits set below */, 0.0 /* 1: ln2/64 hi */, 0.0 /* 2: ln2/64 lo */, \
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, \
    1.0 / 7.0, -1.0 / 6.0, 0.2, 1.0 / 3.0, 0.0 }
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
    const uint64_t b0 = 0x40571547652b82feULL, b1 = 0x3f862e42fef00000ULL, b2 = 0x3d7473de6af278edULL;
    memcpy(&k[0], &b0, 8); memcpy(&k[1], &b1, 8); memcpy(&k[2], &b2, 8);
}

// tables (doubles): [0,128) exp pairs (2^(j/64) hi, lo), j < 64; then 257 log pairs (invc, logc lo) and 257
// logc hi values (entry 256 serves u == 2.0).  Pairs are read with one 16-byte load.
#define PHI_NLT 257
#define PHI_OFF_EXP 0
#define PHI_OFF_LOGP 128
#define PHI_OFF_LOGC_HI (128 + 2 * PHI_NLT)
#define PHI_TAB_DOUBLES (128 + 3 * PHI_NLT + 1)   /* 900, even */

PHI_FN double phi_fast(double d, phi_tab_t tab)
{
    const double INVLN2N = PHI_K(0);                         // 64/ln2
    const double LN2HIN = PHI_K(1);                          // ln2/64, 33 significant bits
    const double LN2LON = PHI_K(2);
    const double SHIFT = 6755399441055744.0;                 // 1.5 * 2^52
    // ---- t = exp(-|d|); |d| clamped to < 1401 on the high word (exp(-1400) == 0 either way)
    uint32_t dhi = PHI_HI(d) & 0x7fffffffu;
    dhi = dhi < 0x4095e000u ? dhi : 0x4095e000u;             // 0x4095e000_00000000 = 1400.0
    const double dd = PHI_HILO(dhi, PHI_LO(d));
    double kd = PHI_FMA(-dd, INVLN2N, SHIFT);
    const int32_t ki = (int32_t)PHI_LO(kd);
    kd -= SHIFT;
    double r = PHI_FMA(kd, -LN2HIN, -dd);
    r = PHI_FMA(kd, -LN2LON, r);
    const int j = ki & 63;
    const int e = ki >> 6;                                    // -2020 .. 0
    const double r2 = r * r;
    double q = PHI_FMA(r, PHI_K(3), PHI_K(4));
    q = PHI_FMA(r, q, PHI_K(5));
    q = PHI_FMA(r, q, PHI_K(6));
    q = PHI_FMA(r, q, 0.5);
    const double p = PHI_FMA(r2, q, r);                       // exp(r) - 1
    double th, tl;
    PHI_LD2(tab, PHI_OFF_EXP, j, th, tl);
    const double tm = th + PHI_FMA(th, p, tl);                // 2^(j/64) * exp(r), in [1,2)
    // t = tm * 2^e in two exact steps (2^e1 on the exponent field, 2^e2 as a factor): the product with
    // 2^e2 is folded into the two fmas below, so a subnormal t is rounded exactly once
    const int e1 = e >> 1, e2 = e - e1;                       // both >= -1010
    const double t1 = PHI_HILO(PHI_HI(tm) + ((uint32_t)e1 << 20), PHI_LO(tm));
    const double s2 = PHI_HILO((uint32_t)(1023 + e2) << 20, 0u);
    // ---- log1p(t), 0 <= t <= 1
    const double u = PHI_FMA(t1, s2, 1.0);
    const double c = PHI_FMA(t1, s2, -(u - 1.0));
    const int i = (int)(PHI_HI(u) >> 12) - 0x3ff00;           // 0..255, 256 iff u == 2.0
    double ic, lclo;
    PHI_LD2(tab, PHI_OFF_LOGP, i, ic, lclo);
    const double rr = PHI_FMA(u, ic, -1.0);
    double w = PHI_FMA(rr, PHI_K(7), PHI_K(8));
    w = PHI_FMA(rr, w, PHI_K(9));
    w = PHI_FMA(rr, w, -0.25);
    w = PHI_FMA(rr, w, PHI_K(10));
    w = PHI_FMA(rr, w, -0.5);
    const double ci = c * ic;                                 // c/u to first order ...
    double tail = PHI_FMA(-ci, rr, ci) + lclo;               // ... times (1 - rr)
    tail = PHI_FMA(rr * rr, w, tail);
    return PHI_LD(tab, PHI_OFF_LOGC_HI + i) + (rr + tail);
}

// fill `tab` (PHI_TAB_DOUBLES doubles) from the generated bit patterns in phi_tables.h
#ifdef PHI_WANT_FILL
static void phi_fill_table(double* tab)
{
    for (int j = 0; j < 64; ++j) {
        memcpy(&tab[PHI_OFF_EXP + 2 * j], &PHI_EXP_HI[j], 8);
        memcpy(&tab[PHI_OFF_EXP + 2 * j + 1], &PHI_EXP_LO[j], 8);
    }
    for (int i = 0; i < PHI_NLT; ++i) {
        memcpy(&tab[PHI_OFF_LOGP + 2 * i], &PHI_INVC[i], 8);
        memcpy(&tab[PHI_OFF_LOGP + 2 * i + 1], &PHI_LOGC_LO[i], 8);
        memcpy(&tab[PHI_OFF_LOGC_HI + i], &PHI_LOGC_HI[i], 8);
    }
    tab[PHI_TAB_DOUBLES - 1] = 0.0;
#ifndef __CUDACC__
    phi_fill_k(g_phi_k);
#endif
}
#endif
