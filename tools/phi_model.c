/* host build of phi_impl.h for accuracy testing (tests/test_phi_accuracy.py) */
#include <stdint.h>
#include <string.h>
#include "../echoseal_b200/csrc/phi_tables.h"
#define PHI_WANT_FILL
#include "../echoseal_b200/csrc/phi_impl.h"

static double g_tab[PHI_TAB_DOUBLES];
static int g_init = 0;
void phi_fast_array(const double* d, double* out, int n)
{
    if (!g_init) { phi_fill_table(g_tab); g_init = 1; }
    for (int k = 0; k < n; k++) out[k] = phi_fast(d[k], g_tab);
}
void phi_libm_array(const double* d, double* out, int n)
{
    for (int k = 0; k < n; k++) out[k] = log1p(exp(-d[k]));
}
void psi_fast_array(const double* d, double* out, int n) { if (!g_init) { phi_fill_table(g_tab); g_init = 1; } for (int i = 0; i < n; ++i) out[i] = psi_fast(d[i], g_tab); }
