/*
 * oracle/polar_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's CRC-aided SCL polar decoder and encoder
 * (rtwm/fastpolar.py) in IEEE double arithmetic with glibc exp/log1p.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / CPU baseline.
 *
 * Parity status: PINNED against outputs of the reference itself run in the build
 * container (tests/golden/polar_golden.npz, made by tests/golden/make_polar_golden.py)
 * and against the reference's own polar tests' properties (tests/test_polar.py:40-107).
 *
 * What follows which reference lines:
 *   f-combine      rtwm/fastpolar.py:18-23   logaddexp(a,b) - logaddexp(0,a+b)   (LLR = log P1/P0)
 *   g-combine      rtwm/fastpolar.py:26-29   b + (1-2u)*a
 *   path penalty   rtwm/fastpolar.py:32-40   log1p(exp(-|l|)) (+|l| if bit != [l>=0])
 *   SC recursion   rtwm/fastpolar.py:127-154 (lazy tree there; standard per-level arrays here — same values)
 *   partial sums   rtwm/fastpolar.py:156-190 parent = (left^right | right)
 *   list step      rtwm/fastpolar.py:280-330 frozen bits penalised; 2|P| candidates appended in
 *                                            (path idx, bit 0, bit 1) order; STABLE sort by metric; keep L
 *   fast path      rtwm/fastpolar.py:261-276 hard decision -> transform -> zero frozen -> CRC
 *   final pick     rtwm/fastpolar.py:332-359 (done by the caller from the sorted list we return)
 *   CRC-8          rtwm/fastpolar.py:362-371 poly 0x07, init 0, MSB first
 *   transform      rtwm/fastpolar.py:376-389 x[i:i+h] ^= x[i+h:i+2h], natural order
 *   encode         rtwm/fastpolar.py:237-252
 *
 * np.logaddexp (numpy/core/src/npymath/npy_math_internal.h.src, npy_logaddexp):
 *   x==y -> x+ln2 ; d=x-y ; d>0 -> x+log1p(exp(-d)) ; d<=0 -> y+log1p(exp(d)) ; else NaN
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NMAX 1024
#define NLOG 10
#define LMAX 32

static const double LOGE2 = 0.693147180559945309417232121458176568;

/*
 * phi(d) = log1p(exp(-d)), d >= 0.  Default build: glibc, i.e. the reference's arithmetic.
 * -DORACLE_PHI_FAST builds the "device arithmetic model" instead (liboracle_fastphi.so): the
 * very same IEEE operation sequence the CUDA kernel runs (echoseal_b200/csrc/phi_impl.h), so
 * that library predicts the GPU's path lists bit-for-bit, ties included; the difference between
 * the two libraries isolates the effect of <1.4-ulp libm differences (hard part 1).
 */
#ifdef ORACLE_PHI_FAST
#include "../echoseal_b200/csrc/phi_tables.h"
#define PHI_WANT_FILL
#include "../echoseal_b200/csrc/phi_impl.h"
static double g_phi_tab[PHI_TAB_DOUBLES];
static int g_phi_init = 0;
static void phi_init(void) { phi_fill_table(g_phi_tab); g_phi_init = 1; }
static inline double phi(double d) { return phi_fast(d, g_phi_tab); }
#else
static inline double phi(double d) { return log1p(exp(-d)); }
#endif

static inline double np_logaddexp(double x, double y)
{
    if (x == y) return x + LOGE2;
    double d = x - y;
    if (d > 0) return x + phi(d);
    if (d <= 0) return y + phi(-d);
    return d; /* NaN */
}

static inline double f_comb(double a, double b)
{
#if defined(ORACLE_PHI_FAST) && defined(ORACLE_F_PSI)
    /* device model: the kernel's form of the same quantity, psi(a-b) - psi(a+b) with psi(x) = |x|/2 + phi(|x|)
     * (max(a,b) - max(0,a+b) = (|a-b| - |a+b|)/2); echoseal_b200/csrc/scl.cu fcomb2 */
    return psi_fast(a - b, g_phi_tab) - psi_fast(a + b, g_phi_tab);
#else
    return np_logaddexp(a, b) - np_logaddexp(0.0, a + b);
#endif
}

static inline double penalty(double l, int bit)
{
    double al = fabs(l);
    double p = phi(al);
    int pref = (l >= 0.0) ? 1 : 0;
    if (bit != pref) p += al;
    return p;
}

typedef struct {
    double metric;
    /* alpha[l] holds the current node of level l (size 2^(10-l)), l=1..10, packed */
    double alpha[NMAX];          /* offset(l) = NMAX - 2^(11-l) ... see AOFF */
    uint8_t bl[NMAX];            /* left-child partial sums per level, same packing */
    uint8_t u[NMAX];
    uint8_t xhat[NMAX];          /* root partial sums (codeword estimate), filled at the end */
} path_t;

/* level l (1..10) array of size s=2^(10-l) lives at [AOFF(l), AOFF(l)+s) ; sum of sizes = 1023 */
#define AOFF(l) (NMAX - (1 << (11 - (l))))

void es_oracle_polar_transform(uint8_t *x, int n)
{
    for (int h = 1; h < n; h <<= 1)
        for (int i = 0; i < n; i += 2 * h)
            for (int k = 0; k < h; k++) x[i + k] ^= x[i + h + k];
}

uint8_t es_oracle_crc8(const uint8_t *bits, int n)
{
    uint8_t reg = 0;
    for (int i = 0; i < n; i++) {
        reg ^= (uint8_t)((bits[i] & 1) << 7);
        if (reg & 0x80) reg = (uint8_t)((reg << 1) ^ 0x07);
        else reg = (uint8_t)(reg << 1);
    }
    return reg;
}

/* frozen: uint8[1024] (1 = frozen). info bits: K-8 values 0/1. out: codeword bits uint8[1024]. */
void es_oracle_polar_encode(const uint8_t *frozen, int K, const uint8_t *info, uint8_t *cw)
{
    int ninfo = K - 8;
    uint8_t crc = es_oracle_crc8(info, ninfo);
    memset(cw, 0, NMAX);
    int j = 0;
    for (int i = 0; i < NMAX; i++) {
        if (frozen[i]) continue;
        if (j < ninfo) cw[i] = info[j] & 1;
        else cw[i] = (crc >> (7 - (j - ninfo))) & 1;
        j++;
    }
    es_oracle_polar_transform(cw, NMAX);
}

static int crc_ok_u(const uint8_t *frozen, int K, const uint8_t *u, uint8_t *info_out)
{
    int ninfo = K - 8, j = 0;
    uint8_t crcbits = 0;
    for (int i = 0; i < NMAX; i++) {
        if (frozen[i]) continue;
        if (j < ninfo) info_out[j] = u[i];
        else crcbits = (uint8_t)((crcbits << 1) | (u[i] & 1));
        j++;
    }
    return es_oracle_crc8(info_out, ninfo) == crcbits;
}

/* Device model only: the kernel decodes the detector's sign-flipped variant -llr (rtwm/detector.py:405-413)
 * from +llr: f(-a,-b) = f(a,b), so the first half of the tree is that of +llr, and the level-1 g node of
 * -llr is minus the one of +llr.  g_neg selects that form; the reference build negates the input instead. */
static __thread int g_neg = 0;
static void sc_step_llr_to(path_t *p, const double *llr0, int i, int last);
static void sc_step_llr(path_t *p, const double *llr0, int i) { sc_step_llr_to(p, llr0, i, NLOG); }

static void sc_step_llr_to(path_t *p, const double *llr0, int i, int last)
{
    int l0;
    if (i == 0) l0 = 0;
    else l0 = NLOG - __builtin_ctz((unsigned)i);
    if (l0 >= 1) { /* g-node at level l0 */
        int s = 1 << (NLOG - l0);
        const double *par = (l0 == 1) ? llr0 : &p->alpha[AOFF(l0 - 1)];
        double *dst = &p->alpha[AOFF(l0)];
        const uint8_t *b = &p->bl[AOFF(l0)];
        for (int k = 0; k < s; k++)
            dst[k] = par[k + s] + (1.0 - 2.0 * (double)b[k]) * par[k];
        if (l0 == 1 && g_neg)
            for (int k = 0; k < s; k++) dst[k] = -dst[k];
    }
    for (int l = l0 + 1; l <= last; l++) {
        int s = 1 << (NLOG - l);
        const double *par = (l == 1) ? llr0 : &p->alpha[AOFF(l - 1)];
        double *dst = &p->alpha[AOFF(l)];
        for (int k = 0; k < s; k++) dst[k] = f_comb(par[k], par[k + s]);
    }
}

static void sc_extend(path_t *p, int i, int bit)
{
    uint8_t tmp[NMAX], tmp2[NMAX];
    p->u[i] = (uint8_t)bit;
    int l = NLOG, s = 1;
    tmp[0] = (uint8_t)bit;
    while (l > 0 && ((i >> (NLOG - l)) & 1)) {
        const uint8_t *left = &p->bl[AOFF(l)];
        for (int k = 0; k < s; k++) { tmp2[k] = left[k] ^ tmp[k]; tmp2[k + s] = tmp[k]; }
        s <<= 1; l--;
        memcpy(tmp, tmp2, (size_t)s);
    }
    if (l > 0) memcpy(&p->bl[AOFF(l)], tmp, (size_t)s);
    else memcpy(p->xhat, tmp, NMAX);
}

typedef struct { double m; int idx; int bit; } cand_t;

/*
 * Decode one codeword.
 *   llr        double[1024]  (LLR = log P1/P0; positive => bit 1)
 *   frozen     uint8[1024]
 *   K, L       info+crc bits, list size (<= LMAX)
 * Outputs:
 *   hard_info  uint8[K-8] hard-decision fast-path candidate, *hard_crc its CRC flag
 *   path_info  uint8[L*(K-8)] final paths in ascending-metric (stable) order
 *   path_metric double[L], path_crc int[L], *npaths
 *   stats[0] = min absolute prune gap (m[L]-m[L-1] over all full prune steps; +inf if none)
 *   stats[1] = min relative prune gap (gap / max(|m[L]|,1e-300))
 *   stats[2] = number of prune steps with gap == 0 (exact ties at the cut)
 *   stats[3] = number of exact ties anywhere among adjacent sorted candidates
 * The SCL stage is ALWAYS run (the caller applies the reference's early return).
 */
int es_oracle_scl_decode(const double *llr, const uint8_t *frozen, int K, int L,
                         uint8_t *hard_info, int *hard_crc,
                         uint8_t *path_info, double *path_metric, int *path_crc, int *npaths,
                         double *stats)
{
    if (L < 1 || L > LMAX) return -1;
#ifdef ORACLE_PHI_FAST
    if (!g_phi_init) phi_init();
#endif
    int ninfo = K - 8;
    /* fast path candidate */
    {
        uint8_t h[NMAX];
        for (int i = 0; i < NMAX; i++) h[i] = (g_neg ? -llr[i] : llr[i]) > 0.0;
        es_oracle_polar_transform(h, NMAX);
        for (int i = 0; i < NMAX; i++) if (frozen[i]) h[i] = 0;
        *hard_crc = crc_ok_u(frozen, K, h, hard_info);
    }
    path_t *P = (path_t *)malloc(sizeof(path_t) * (size_t)L);
    path_t *Q = (path_t *)malloc(sizeof(path_t) * (size_t)L);
    if (!P || !Q) { free(P); free(Q); return -2; }
    memset(&P[0], 0, sizeof(path_t));
    int np = 1;
    double min_gap = INFINITY, min_rel = INFINITY, nzero = 0, nties = 0;
    cand_t c[2 * LMAX], t;
    for (int i = 0; i < NMAX; i++) {
#ifdef ORACLE_PHI_FAST
        /* Device model only: the kernel's rate-0 node rule (echoseal_b200/csrc/scl.cu, r0_sum).  At a quad
         * boundary i > 0 that starts an aligned all-frozen node of s >= 4 bits (s maximal), every path adds
         * sum_k ln(1 + exp(a_k)) over the node's own LLRs -- mathematically the sum of the s leaf penalties
         * the reference walk below accumulates -- in the kernel's summation order, and decides s zeros. */
        if (i > 0 && (i & 3) == 0) {
            int s = 0;
            for (int c2 = 4; c2 <= NMAX / 2 && (i % c2) == 0 && i + c2 <= NMAX; c2 <<= 1) {
                int all = 1;
                for (int k = 0; k < c2 && all; k++) all = frozen[i + k] != 0;
                if (!all) break;
                s = c2;
            }
            if (s) {
                int lv = NLOG - __builtin_ctz((unsigned)s);
                for (int p = 0; p < np; p++) {
                    sc_step_llr_to(&P[p], llr, i, lv);
                    const double *a = &P[p].alpha[AOFF(lv)];
                    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#ifdef ORACLE_F_PSI
                    for (int k = 0; k < s; k++) acc[k & 3] += fma(a[k], 0.5, psi_fast(a[k], g_phi_tab));   /* ln(1 + e^a) = a/2 + psi(a) */
#else
                    for (int k = 0; k < s; k++) acc[k & 3] += phi(fabs(a[k])) + fmax(a[k], 0.0);
#endif
                    P[p].metric += (acc[0] + acc[1]) + (acc[2] + acc[3]);
                    for (int k = 0; k < s; k++) sc_extend(&P[p], i + k, 0);
                }
                i += s - 1;
                continue;
            }
        }
#endif
        for (int p = 0; p < np; p++) sc_step_llr(&P[p], llr, i);
        if (frozen[i]) {
            for (int p = 0; p < np; p++) {
                P[p].metric += penalty(P[p].alpha[AOFF(NLOG)], 0);
                sc_extend(&P[p], i, 0);
            }
            continue;
        }
        int nc = 0;
        for (int p = 0; p < np; p++) {
            double l = P[p].alpha[AOFF(NLOG)], base = P[p].metric;
            c[nc].m = base + penalty(l, 0); c[nc].idx = p; c[nc].bit = 0; nc++;
            c[nc].m = base + penalty(l, 1); c[nc].idx = p; c[nc].bit = 1; nc++;
        }
        /* stable insertion sort by metric */
        for (int a = 1; a < nc; a++) {
            t = c[a];
            int b = a - 1;
            while (b >= 0 && c[b].m > t.m) { c[b + 1] = c[b]; b--; }
            c[b + 1] = t;
        }
        for (int a = 1; a < nc; a++) if (c[a].m == c[a - 1].m) nties += 1;
        int ns = nc < L ? nc : L;
        if (nc > L) {
            double gap = c[L].m - c[L - 1].m;
            double den = fabs(c[L].m) > 1e-300 ? fabs(c[L].m) : 1e-300;
            if (gap < min_gap) min_gap = gap;
            if (gap / den < min_rel) min_rel = gap / den;
            if (gap == 0.0) nzero += 1;
        }
        for (int a = 0; a < ns; a++) {
            memcpy(&Q[a], &P[c[a].idx], sizeof(path_t));
            Q[a].metric = c[a].m;
            sc_extend(&Q[a], i, c[a].bit);
        }
        path_t *sw = P; P = Q; Q = sw;
        np = ns;
    }
    /* ascending metric, stable */
    int ord[LMAX];
    for (int a = 0; a < np; a++) ord[a] = a;
    for (int a = 1; a < np; a++) {
        int o = ord[a], b = a - 1;
        while (b >= 0 && P[ord[b]].metric > P[o].metric) { ord[b + 1] = ord[b]; b--; }
        ord[b + 1] = o;
    }
    for (int a = 0; a < np; a++) {
        path_metric[a] = P[ord[a]].metric;
        path_crc[a] = crc_ok_u(frozen, K, P[ord[a]].u, path_info + (size_t)a * (size_t)ninfo);
    }
    *npaths = np;
    if (stats) { stats[0] = min_gap; stats[1] = min_rel; stats[2] = nzero; stats[3] = nties; }
    free(P); free(Q);
    return 0;
}

/*
 * Batch wrapper, float32 LLR input (what rtwm/detector.py:405 and SURVEY §8d config 4 feed the
 * decoder; rtwm/fastpolar.py:258 widens to float64).  pthread workers over codewords.
 *   flags bit0: skip the list stage when the hard decision passes CRC (reference behaviour
 *               with validator=None, rtwm/fastpolar.py:269-276); outputs npaths=0 for those.
 *   nthreads <= 0 -> 1.
 */
typedef struct {
    const float *llr; int ncw; const uint8_t *frozen; int K, L, flags;
    uint8_t *hard_info; int *hard_crc; uint8_t *path_info; double *path_metric; int *path_crc; int *npaths;
    double *stats; int next; int rc;
} batch_t;

static void batch_one(batch_t *b, int w)
{
    int ninfo = b->K - 8, L = b->L;
    double d[NMAX];
    if (b->flags & 2) {     /* detector pairing: codeword w = (w odd ? - : +) row w/2 */
        const float *src = b->llr + (size_t)(w >> 1) * NMAX;
#ifdef ORACLE_PHI_FAST
        for (int i = 0; i < NMAX; i++) d[i] = (double)src[i];
        g_neg = w & 1;
#else
        const float sg = (w & 1) ? -1.0f : 1.0f;
        for (int i = 0; i < NMAX; i++) d[i] = (double)(sg * src[i]);
#endif
    } else {
        g_neg = 0;
        for (int i = 0; i < NMAX; i++) d[i] = (double)b->llr[(size_t)w * NMAX + i];
    }
    if (b->flags & 1) {
        uint8_t h[NMAX];
        for (int i = 0; i < NMAX; i++) h[i] = (g_neg ? -d[i] : d[i]) > 0.0;
        es_oracle_polar_transform(h, NMAX);
        for (int i = 0; i < NMAX; i++) if (b->frozen[i]) h[i] = 0;
        int ok = crc_ok_u(b->frozen, b->K, h, b->hard_info + (size_t)w * ninfo);
        if (ok) {
            b->hard_crc[w] = 1; b->npaths[w] = 0;
            if (b->stats) { double *s = b->stats + 4 * (size_t)w; s[0] = INFINITY; s[1] = INFINITY; s[2] = 0; s[3] = 0; }
            g_neg = 0;
            return;
        }
    }
    int r = es_oracle_scl_decode(d, b->frozen, b->K, L,
                                 b->hard_info + (size_t)w * ninfo, b->hard_crc + w,
                                 b->path_info + (size_t)w * L * ninfo, b->path_metric + (size_t)w * L,
                                 b->path_crc + (size_t)w * L, b->npaths + w,
                                 b->stats ? b->stats + 4 * (size_t)w : NULL);
    g_neg = 0;
    if (r) __atomic_store_n(&b->rc, r, __ATOMIC_RELAXED);
}

static void *batch_worker(void *arg)
{
    batch_t *b = (batch_t *)arg;
    for (;;) {
        int w = __atomic_fetch_add(&b->next, 1, __ATOMIC_RELAXED);
        if (w >= b->ncw) break;
        batch_one(b, w);
    }
    return NULL;
}

int es_oracle_scl_decode_batch(const float *llr, int ncw, const uint8_t *frozen, int K, int L, int flags,
                               int nthreads,
                               uint8_t *hard_info, int *hard_crc,
                               uint8_t *path_info, double *path_metric, int *path_crc, int *npaths,
                               double *stats)
{
    batch_t b = { llr, ncw, frozen, K, L, flags, hard_info, hard_crc, path_info, path_metric, path_crc,
                  npaths, stats, 0, 0 };
#ifdef ORACLE_PHI_FAST
    if (!g_phi_init) phi_init();
#endif
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1) { batch_worker(&b); return b.rc; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads; t++)
        if (pthread_create(&th[started], NULL, batch_worker, &b) == 0) started++;
    if (started == 0) batch_worker(&b);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return b.rc;
}
