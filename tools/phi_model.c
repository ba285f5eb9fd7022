/* host build of phi_impl.h for accuracy testing (tests/test_phi_accuracy.py) */
#include <stdint.h>
#include "../echoseal_b200/csrc/phi_tables.h"
#define PHI_FN static inline
#include "../echoseal_b200/csrc/phi_impl.h"

static double g_tab[PHI_TAB_DOUBLES];
static int g_init = 0;
static void init(void)
{
    for (int j = 0; j < 64; j++) { g_tab[PHI_OFF_EXP_HI + j] = PHI_U2D(PHI_EXP_HI[j]); g_tab[PHI_OFF_EXP_LO + j] = PHI_U2D(PHI_EXP_LO[j]); }
    for (int i = 0; i < 256; i++) {
        g_tab[PHI_OFF_INVC + i] = PHI_U2D(PHI_INVC[i]);
        g_tab[PHI_OFF_LOGC_HI + i] = PHI_U2D(PHI_LOGC_HI[i]);
        g_tab[PHI_OFF_LOGC_LO + i] = PHI_U2D(PHI_LOGC_LO[i]);
    }
    g_init = 1;
}
void phi_fast_array(const double* d, double* out, int n)
{
    if (!g_init) init();
    for (int k = 0; k < n; k++) out[k] = phi_fast(d[k], g_tab);
}
void phi_libm_array(const double* d, double* out, int n)
{
    for (int k = 0; k < n; k++) out[k] = log1p(exp(-d[k]));
}
