/*
 * oracle/dsp_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restatement of scipy.signal.lfilter (scipy 1.18.1, third-party dependency of the reference,
 * not vendored under /root/reference): IIR filtering in direct-form II transposed,
 *     y[n]   = b[0] x[n] + z[0]
 *     z[i]   = b[i+1] x[n] + z[i+1] - a[i+1] y[n]      (i = 0 .. K-3)
 *     z[K-2] = b[K-1] x[n] - a[K-1] y[n]
 * with b, a normalised by a[0] and zero-padded to K = max(len(a), len(b)).
 * Call sites in the reference: rtwm/detector.py:60,67 (RX band-pass, preamble template),
 * rtwm/detector.py:277 (impulse response for the matched filter), rtwm/embedder.py:141-143 (TX).
 * Parity: pinned against scipy itself in tests/test_oracle_rx.py and against reference outputs
 * in tests/golden/rx_golden.npz.
 */
#include <stddef.h>

#define KMAX 64

/* z: in/out state, length K (last element unused / zero). */
void es_oracle_lfilter(const double *b, int nb, const double *a, int na,
                       const double *x, double *y, long n, double *z)
{
    int K = nb > na ? nb : na;
    if (K > KMAX) return;
    double bb[KMAX] = {0}, aa[KMAX] = {0};
    for (int i = 0; i < nb; i++) bb[i] = b[i] / a[0];
    for (int i = 0; i < na; i++) aa[i] = a[i] / a[0];
    for (long t = 0; t < n; t++) {
        double xn = x[t];
        double yn = z[0] + bb[0] * xn;
        for (int i = 0; i < K - 2; i++) z[i] = z[i + 1] + xn * bb[i + 1] - yn * aa[i + 1];
        if (K >= 2) z[K - 2] = xn * bb[K - 1] - yn * aa[K - 1];
        y[t] = yn;
    }
}
