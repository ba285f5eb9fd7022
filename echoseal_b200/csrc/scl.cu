// echoseal_b200/csrc/scl.cu — batched CRC-aided SCL decoder for Polar(1024,K)+CRC-8 (list size <= 8 at full speed,
// 9..32 on the wide-list kernel), the hard-decision fast path and the matching encoder, hand-written for sm_100a.
//
// Replaces rtwm/fastpolar.py:254-359 (PolarCode.decode), :237-252 (encode), :362-389 (CRC-8, transform)
// behind rtwm/polar_fast.py:26-87.  Arithmetic is IEEE double with the reference's formulas:
//   f = logaddexp(a,b) - logaddexp(0,a+b)          (rtwm/fastpolar.py:18-23; LLR = log P1/P0), evaluated as
//       psi(a-b) - psi(a+b), psi(x) = |x|/2 + log1p(exp(-|x|))   (the same quantity)
//   g = b + (1-2u) a                                (rtwm/fastpolar.py:26-29)
//   penalty(l,bit) = log1p(exp(-|l|)) (+|l| if bit != [l>=0])   (rtwm/fastpolar.py:32-40)
//   frozen bits are penalised too; candidates are ranked by (metric, path order, bit) = the reference's
//   stable sort over candidates appended in (path idx, bit 0, bit 1) order (rtwm/fastpolar.py:288-299).
//
// Mapping of scl_list_kernel (B200-first):
//   * one THREAD per list path, 8 lanes per codeword, 4 codewords per warp, 16 warps per CTA, one persistent CTA per SM;
//     warps never wait on each other.
//   * LLR tree: level l (1..10) keeps only its CURRENT node (2^(10-l) doubles) per path slot.  Levels 9..10 live in
//     registers, levels S..8 in shared memory, levels < S in a global scratch [element][codeword(4)][slot(8)] (one
//     256-byte row per element, quarter-interleaved) that is streamed through a TMA bulk-copy ring (cp.async.bulk +
//     mbarrier); levels 1..2 carry an L2 evict_first policy.
//   * lazy copy without reference counts: every path rewrites a level at the same bit index, so a path
//     always writes its OWN slot and a clone only copies a 30-bit word of per-level slot pointers.
//   * partial sums are bit-packed: levels 6..10 in one register, levels 1..5 as pointer-indirected
//     32-bit words; the root (codeword estimate) is transformed back to u-hat at the end and its un-frozen bits are
//     bit-compressed into the payload.
//   * list pruning: 16 candidates ranked through unique 64-bit integer keys exchanged in shared memory;
//     survivors/clones matched by ballots.
#include "common.cuh"
#include <math_constants.h>
#include <string.h>
#include "phi_tables.h"
#define PHI_WANT_FILL
#include "phi_impl.h"

namespace es {

__constant__ uint32_t c_frozen[32];      // bit (i&31) of word (i>>5): 1 = frozen
__constant__ uint16_t c_datapos[1024];   // ascending un-frozen positions (K entries used)
__constant__ int c_K;                    // info + CRC bits
__device__ uint16_t d_datapos[1024];     // the same positions in global memory, for lane-indexed (divergent) reads
// rate-0 node map per quad (bits 4q..4q+3): 0 = ordinary quad; v in 1..8 = first quad of a maximal aligned
// all-frozen node of 4 << (v-1) bits; 255 = interior quad of such a node
__constant__ uint8_t c_r0[256];
__constant__ uint8_t c_crc8[256];        // CRC-8 (poly 0x07, MSB first) of one byte
__constant__ uint32_t c_cmp[32][5];      // per word of 32 positions: the five move masks that compress its un-frozen bits to the low end
__constant__ uint32_t c_cnt[32];         // un-frozen positions per word

// what each device's constant tables currently hold (es_polar_set_code)
struct CodeDev { int ready = 0; int K = 0; uint32_t frozen[32] = {0}; };
static CodeDev g_code[ES_MAX_DEVICES];
#define g_code_ready (g_code[current_device()].ready)
#define g_K (g_code[current_device()].K)

// ---------------------------------------------------------------------------------------------
// arithmetic
// ---------------------------------------------------------------------------------------------
// phi(d) = log1p(exp(-d)), d >= 0: branch-free table-driven double routine (phi_impl.h), < 1.4 ulp —
// the same accuracy class as the reference's libm composition, a quarter of the instructions.
//
// The routine exists exactly THREE times in the kernel image, as out-of-line functions of 4, 2 and 1
// interleaved evaluations.  The kernel is instruction-cache sensitive: with phi
// inlined, every phi-heavy loop cost 4 KB of SASS and the hot code sat on the 32 KB edge.  ptxas
// allocates registers across these calls (no marshalling: a call costs CALL + RET + one MOV).
// f = logaddexp(a,b) - logaddexp(0,a+b) (rtwm/fastpolar.py:18-23) is evaluated as psi(a-b) - psi(a+b) with
// psi(x) = |x|/2 + phi(|x|) (phi_impl.h): max(a,b) - max(0,a+b) = (|a-b| - |a+b|)/2, so no max/select terms travel
// with phi and the routines take one operand per evaluation.
struct D4 { double a, b, c, d; };
struct D2 { double a, b; };

// The routines take the shared-window address of the phi tables from the CTA's dynamic shared memory base themselves
// (a uniform register: the table reads are [index + base] with no address add per evaluation).
__device__ __forceinline__ uint32_t smem_base();
// N evaluations of psi_fast step by step in lock-step (the operation sequence of phi_impl.h per element, bit for bit):
// written interleaved so that the schedule does not depend on how ptxas happens to merge N separate call chains.
#ifndef ES_SCL_F_PSI
#define ES_SCL_F_PSI 1       // 1: f = psi(a-b) - psi(a+b); 0: the reference's two logaddexp terms with their own rounding points (15 % slower)
#endif
// PSI: out = |x|/2 + phi(|x|); otherwise out = mx + phi(|x|), the reference's logaddexp term (max + log1p(exp(-|d|))).
template <int N, bool PSI>
__device__ __forceinline__ void psi_lockstep(const double (&x)[N], const double (&mx)[N], double (&out)[N], uint32_t tab)
{
    const double SHIFT = 6755399441055744.0;
    double dd[N], ax[N], kd[N], r[N], q[N], p[N], th[N], t[N], u[N], ic[N], lc[N], rr[N], w[N];
    int32_t ki[N];
#pragma unroll
    for (int c = 0; c < N; ++c) {
        uint32_t dhi = PHI_HI(x[c]) & 0x7fffffffu;
        ax[c] = PHI_HILO(dhi, PHI_LO(x[c]));
        dhi = dhi < 0x40500000u ? dhi : 0x40500000u;
        dd[c] = PHI_HILO(dhi, PHI_LO(x[c]));
    }
#pragma unroll
    for (int c = 0; c < N; ++c) kd[c] = PHI_FMA(-dd[c], PHI_K(0), SHIFT);
#pragma unroll
    for (int c = 0; c < N; ++c) { ki[c] = (int32_t)PHI_LO(kd[c]); kd[c] -= SHIFT; }
#pragma unroll
    for (int c = 0; c < N; ++c) r[c] = PHI_FMA(kd[c], -PHI_K(1), -dd[c]);
#pragma unroll
    for (int c = 0; c < N; ++c) r[c] = PHI_FMA(kd[c], -PHI_K(2), r[c]);
#pragma unroll
    for (int c = 0; c < N; ++c) { const double th0 = PHI_LD(tab, PHI_OFF_EXP + (ki[c] & 63)); th[c] = PHI_HILO(PHI_EXPADD(PHI_HI(th0), ki[c]), PHI_LO(th0)); }
#pragma unroll
    for (int c = 0; c < N; ++c) q[c] = PHI_FMA(r[c], PHI_K(3), PHI_K(4));
#pragma unroll
    for (int c = 0; c < N; ++c) q[c] = PHI_FMA(r[c], q[c], PHI_K(5));
#pragma unroll
    for (int c = 0; c < N; ++c) q[c] = PHI_FMA(r[c], q[c], 0.5);
#pragma unroll
    for (int c = 0; c < N; ++c) p[c] = PHI_FMA(r[c] * r[c], q[c], r[c]);
#pragma unroll
    for (int c = 0; c < N; ++c) t[c] = PHI_FMA(th[c], p[c], th[c]);
#pragma unroll
    for (int c = 0; c < N; ++c) u[c] = 1.0 + t[c];
#pragma unroll
    for (int c = 0; c < N; ++c) { const int i = (int)(PHI_HI(u[c]) >> 11) - 0x7fe00; PHI_LD2(tab, PHI_OFF_LOGP, i, ic[c], lc[c]); }
#pragma unroll
    for (int c = 0; c < N; ++c) rr[c] = PHI_FMA(u[c], ic[c], -1.0);
#pragma unroll
    for (int c = 0; c < N; ++c) w[c] = PHI_FMA(rr[c], PHI_K(6), -0.25);
#pragma unroll
    for (int c = 0; c < N; ++c) w[c] = PHI_FMA(rr[c], w[c], PHI_K(7));
#pragma unroll
    for (int c = 0; c < N; ++c) w[c] = PHI_FMA(rr[c], w[c], -0.5);
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const double ph = lc[c] + PHI_FMA(rr[c] * rr[c], w[c], rr[c]);
        out[c] = PSI ? PHI_FMA(ax[c], 0.5, ph) : mx[c] + ph;
    }
}
__device__ __noinline__ D4 psi4(double x0, double x1, double x2, double x3)
{
    const double x[4] = {x0, x1, x2, x3};
    double o[4];
    psi_lockstep<4, true>(x, x, o, smem_base());
    D4 r;
    r.a = o[0]; r.b = o[1]; r.c = o[2]; r.d = o[3];
    return r;
}
// f of two element pairs in the reference's own form, numpy's npy_logaddexp terms max + log1p(exp(-|x-y|)) for (a,b) and
// (0,a+b) with the reference's rounding points (rtwm/fastpolar.py:18-23): operands in, results out, nothing else live
// across the call.  x == y needs no special case: phi_fast(0) == ln2 exactly (table entry 256).
__device__ __noinline__ D2 f2(double a0, double b0, double a1, double b1)
{
    const double d0 = a0 - b0, s0 = a0 + b0, d1 = a1 - b1, s1 = a1 + b1;
    const double x[4] = {d0, s0, d1, s1};
    const double m[4] = {(d0 > 0.0) ? a0 : b0, ((0.0 - s0) > 0.0) ? 0.0 : s0, (d1 > 0.0) ? a1 : b1, ((0.0 - s1) > 0.0) ? 0.0 : s1};
    double o[4];
    psi_lockstep<4, false>(x, m, o, smem_base());
    D2 r;
    r.a = o[0] - o[1];
    r.b = o[2] - o[3];
    return r;
}
// four phi evaluations (the rate-0 node sums)
__device__ __noinline__ D4 phi4(double x0, double x1, double x2, double x3)
{
    const double x[4] = {x0, x1, x2, x3}, z[4] = {0.0, 0.0, 0.0, 0.0};
    double o[4];
    psi_lockstep<4, false>(x, z, o, smem_base());
    D4 r;
    r.a = o[0]; r.b = o[1]; r.c = o[2]; r.d = o[3];
    return r;
}
__device__ __noinline__ D2 phi2(double d0, double d1)
{
    const uint32_t tab = smem_base();
    D2 r;
    r.a = phi_fast(d0, tab); r.b = phi_fast(d1, tab);
    return r;
}
__device__ __noinline__ double phi1(double d0) { return phi_fast(d0, smem_base()); }

// f of two element pairs at once (four independent phi chains).  x == y needs no special case (numpy's npy_logaddexp
// has one): phi_fast(0) == ln2 exactly (table entry 256).
__device__ __forceinline__ void fcomb2(double a0, double b0, double a1, double b1, uint32_t tab, double& r0, double& r1)
{
#if ES_SCL_F_PSI
    const D4 P = psi4(a0 - b0, a0 + b0, a1 - b1, a1 + b1);
    r0 = P.a - P.b;
    r1 = P.c - P.d;
#else
    const D2 Q = f2(a0, b0, a1, b1);
    r0 = Q.a;
    r1 = Q.b;
#endif
}
// same value as f(a, b), also handing out the two phi terms: fm = phi(|a-b|), fp = phi(|a+b|).
// They are exactly the phi values the reference's penalty needs for the NEXT (odd) leaf, whose LLR is
// b-a or b+a (rtwm/fastpolar.py:26-40) — so that penalty costs nothing.
__device__ __forceinline__ double fcomb_parts(double a, double b, uint32_t tab, double& fm, double& fp)
{
    const double d = a - b, s = a + b;
    const D2 P = phi2(d, s);
    fm = P.a;
    fp = P.b;
#if ES_SCL_F_PSI
    return __fma_rn(fabs(d), 0.5, fm) - __fma_rn(fabs(s), 0.5, fp);
#else
    const double A = ((d > 0.0) ? a : b) + fm;
    const double B = (((0.0 - s) > 0.0) ? 0.0 : s) + fp;
    return A - B;
#endif
}

// g = b + (1-2u) a (rtwm/fastpolar.py:26-29): (1-2u)*a is exact, so flipping the sign bit of a is the same number
__device__ __forceinline__ double gcomb(double a, double b, uint32_t u)
{
    const double sa = __hiloint2double(__double2hiint(a) ^ (int)(u << 31), __double2loint(a));
    return b + sa;
}

// ---------------------------------------------------------------------------------------------
// asynchronous staging: cp.async.bulk (TMA, SASS UBLKCP) global -> shared, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
#ifndef ES_SCL_L2HINT
#define ES_SCL_L2HINT 1      // 1: levels 1..2 (written once, read once, far apart) travel evict_first; 2: also levels 3..5 evict_last
#endif
__device__ __forceinline__ uint64_t l2_policy(int kind)     // 0 normal, 1 evict_first, 2 evict_last
{
    uint64_t p;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_hint(double* p, double v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// generic-proxy writes (st.global / st.shared) before, async-proxy accesses (bulk copies) after
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// ---------------------------------------------------------------------------------------------
// per-lane state
// ---------------------------------------------------------------------------------------------
struct SclParams {
    const float* llr;        // [nrows][1024] float32 channel LLRs
    const int32_t* index;    // optional list of codeword ids to decode (nullptr = 0..ncw-1)
    int ncw;                 // number of codewords to decode (length of index if given)
    int neg_mode;            // 0: codeword w = +row w ; 1: codeword w = (w&1 ? - : +) row (w>>1)
    int pair;                // neg_mode without index: one lane group decodes +row and -row (shared first half)
    int nunits;              // work units: rows when pair, else codewords
    int list_size;           // 1..8
    double* scratch;         // global alpha scratch, per warp
    size_t scratch_stride;   // doubles per warp
    const double* phi_tab;   // 896 doubles (phi_impl.h layout)
    uint8_t* path_payload;   // [ncw_total][list_size][ (K-8)/8 ]
    uint8_t* path_crc;       // [ncw_total][list_size]
    double* path_metric;     // [ncw_total][list_size]
    int32_t* npaths;         // [ncw_total]
    double* min_margin;      // [ncw_total] or nullptr: smallest relative gap between the last kept and the first dropped candidate
};

constexpr int SCL_S = 6;   // first LLR-tree level kept in shared memory
constexpr int RING_STAGES = 2;      // the passes toggle between two stages (st ^= 1)
#ifndef ES_SCL_RING_ALIAS
#define ES_SCL_RING_ALIAS 1  // stage 1 lives in the rows of level 7 (dead while any DRAM-level pass runs: those rewrite levels <= 6,
#endif                       // and levels 7..8 are recomputed after them); buys the shared memory for 20 warps per SM
constexpr int RING_STAGE_ROWS = 8;
constexpr int RING_STAGE_BYTES = RING_STAGE_ROWS * 256;

#ifndef ES_SCL_W
#define ES_SCL_W 16
#endif
#ifndef ES_SCL_MAXNREG
#define ES_SCL_MAXNREG ((65536 / (ES_SCL_W * 32)) / 8 * 8 > 255 ? 255 : (65536 / (ES_SCL_W * 32)) / 8 * 8)   // one CTA per SM owns the register file
#endif
template <int S> struct SclLayout {
    static constexpr int AROWS = (1 << (11 - S)) - 4;                   // levels S..8 (9 and 10 live in registers)
    static constexpr int BROWS_S = 7;                                    // beta words of levels 3..5 in shared memory
    static constexpr int BROWS_G = 24;                                   // beta words of levels 1..2 in global memory
    static constexpr int SNAP_BYTES = 32 * 32;                           // per-lane (metric, bptr, ord|active) + prune margin at bit 512
    static constexpr int ABYTES = (AROWS * 256 > 32 * 128) ? AROWS * 256 : 32 * 128;   // alpha rows; also holds the root partial sums
    static constexpr int RING_BYTES = (ES_SCL_RING_ALIAS ? 1 : RING_STAGES) * RING_STAGE_BYTES;   // TMA staging ring of the DRAM-level passes
    static constexpr int L7_OFF = ((1 << (11 - S)) - (1 << 4)) * 256;     // rows of level 7 inside the alpha area (8 rows = one stage)
    static constexpr int RING_OFF = ABYTES + BROWS_S * 128 + SNAP_BYTES;
    static constexpr int BAR_OFF = RING_OFF + RING_BYTES;                 // one 8-byte mbarrier per stage
    static constexpr int STASH_OFF = BAR_OFF + 16;                        // 48 bytes per lane: decode state parked across the LLR update
    static constexpr int WARP_BYTES = STASH_OFF + 32 * 48;
    static constexpr int NTH_OFF = PHI_TAB_DOUBLES * 8;                  // [mask 256][n 8] bytes: position of the n-th set bit
    static constexpr int CRC_OFF = NTH_OFF + 2048;                       // CRC-8 byte table
    static constexpr int TAB_BYTES = CRC_OFF + 256;
    static constexpr size_t G_ROWS = 1024 - (1 << (11 - S));           // global alpha rows, levels 1..S-1
    static constexpr size_t G_DOUBLES = G_ROWS * 32 + 1024 * 4 + BROWS_G * 16;   // + level-0 copy [1024][4] + beta rows
};

using SclLY = SclLayout<SCL_S>;

// shared-window address of the CTA's dynamic shared memory (the phi tables sit at its start)
__device__ __forceinline__ uint32_t smem_base()
{
    extern __shared__ __align__(16) unsigned char es_smem[];
    return (uint32_t)__cvta_generic_to_shared(es_smem);
}

// The per-lane addresses are all derived from three words (the warp's shared-memory window address, its global
// scratch pointer, the lane id) on the spot: a dozen pointer registers held across the whole decode were what
// kept ptxas from interleaving the four chains of phi4 at 128 registers per thread.
struct Lane {
    uint32_t wsm;        // shared-window address of this warp's region (alpha rows first)
    double* gw;          // this warp's global scratch: alpha rows of levels 1..S-1, level-0 copy, beta rows
    int lane;
    double m;
    uint32_t ptr, bptr, bs;
    int ord;
    bool active;
    bool neg;            // decoding the sign-flipped variant: level 0 is negated when bit 512 is reached
    uint32_t rphase;     // staging ring: bit st = parity the next wait on stage st has to see
    double mg_gap, mg_den;   // prune margin (kernels built with MG): running minimum of gap / den, kept as the pair

    __device__ __forceinline__ int p() const { return lane & 7; }               // list path = slot owned by this lane
    __device__ __forceinline__ int gbase() const { return lane & 24; }          // first lane of this lane's codeword
    __device__ __forceinline__ double* sa() const                               // shared alpha rows (+gbase)
    { return reinterpret_cast<double*>(__cvta_shared_to_generic(wsm)) + gbase(); }
    __device__ __forceinline__ uint32_t* sb() const                             // shared beta words, levels 3..5
    { return reinterpret_cast<uint32_t*>(__cvta_shared_to_generic(wsm + SclLY::ABYTES)); }
    __device__ __forceinline__ uint32_t* gb() const                             // global beta words, levels 1..2
    { return reinterpret_cast<uint32_t*>(gw + SclLY::G_ROWS * 32 + 1024 * 4); }
    __device__ __forceinline__ double* ga() const { return gw + gbase(); }      // global alpha rows (+gbase)
    __device__ __forceinline__ double* g0() const { return gw + SclLY::G_ROWS * 32 + (lane >> 3); }   // level-0 copy [k][4]
    __device__ __forceinline__ uint32_t tab() const { return smem_base(); }     // phi tables
    __device__ __forceinline__ uint32_t ring(int st) const                      // shared-window address of stage st
    {
#if ES_SCL_RING_ALIAS
        return wsm + (st ? (uint32_t)SclLY::L7_OFF : (uint32_t)SclLY::RING_OFF);
#else
        return wsm + SclLY::RING_OFF + (uint32_t)st * RING_STAGE_BYTES;
#endif
    }
    __device__ __forceinline__ uint32_t rbar() const { return wsm + SclLY::BAR_OFF; }    // its mbarrier at + 8 * st
};

// ---------------------------------------------------------------------------------------------
// addressing.  Levels >= S live in shared memory in natural element order.  Levels 0..S-1 live in global memory
// QUARTER-INTERLEAVED: element k = j*q + r of a node of n = 4q elements sits at position 4r + j, so the four
// elements (r, r+q, r+2q, r+3q) that every consumer of the node needs together are adjacent, and a pass reads
// its source node front to back in contiguous 8-row chunks: one bulk copy per stage.
// ---------------------------------------------------------------------------------------------
// first global row of level lv (1..S-1)
__device__ __forceinline__ int lvl_row0(int lv) { return 1024 - (1 << (11 - lv)); }
// position of element k inside the node of level lv (0..S-1)
__device__ __forceinline__ int ipos(int lv, int k)
{
    const int lq = 8 - lv;                                   // log2 of the quarter size
    return ((k & ((1 << lq) - 1)) << 2) | (k >> lq);
}

// row 0 of the left-child partial-sum words of level l (1..5): 2^(5-l) rows of 32 words
__device__ __forceinline__ uint32_t* beta_rows(const Lane& L, int l)
{
    // shared: level 5 -> row 0, level 4 -> rows 1..2, level 3 -> rows 3..6 ; global: level 2 -> rows 0..7, level 1 -> rows 8..23
    return (l >= 3) ? L.sb() + ((1 << (5 - l)) - 1) * 32 : L.gb() + ((l == 2) ? 0 : 8) * 32;
}

// element 0 of the shared-memory level lv (>= S) in `slot`; element k is k * 32 doubles further
template <int S> __device__ __forceinline__ double* slvl(const Lane& L, int lv, int slot)
{
    return L.sa() + (((1 << (11 - S)) - (1 << (11 - lv))) * 32) + slot;
}
// element k of any level in `slot` (level 0: the warp's copy of the channel LLRs, [position][codeword])
template <int S> __device__ __forceinline__ double* elem_ptr(const Lane& L, int lv, int slot, int k)
{
    if (lv == 0) return L.g0() + ipos(0, k) * 4;
    if (lv >= S) return slvl<S>(L, lv, slot) + k * 32;
    return L.ga() + (lvl_row0(lv) + ipos(lv, k)) * 32 + slot;
}

// f-combine of a whole shared-memory node (32-bit shared-window addresses, 256-byte element pitch):
// dst[k] = f(a[k], b[k]), k < count, count even.  Two elements per trip = four independent phi chains.
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __noinline__ void f_loop(uint32_t a, uint32_t b, uint32_t dst, int count, uint32_t tab)
{
#pragma unroll 1
    for (int k = 0; k + 2 <= count; k += 2) {
        const double a0 = lds_f64(a + k * 256), b0 = lds_f64(b + k * 256), a1 = lds_f64(a + k * 256 + 256), b1 = lds_f64(b + k * 256 + 256);
        double r0, r1;
        fcomb2(a0, b0, a1, b1, tab, r0, r1);
        sts_f64(dst + k * 256, r0);
        sts_f64(dst + k * 256 + 256, r1);
    }
}

// shared-window address of element 0 of the shared-memory level lv (>= S) in `slot`
template <int S> __device__ __forceinline__ uint32_t slvl_a(const Lane& L, int lv, int slot)
{
    return L.wsm + (uint32_t)(((1 << (11 - S)) - (1 << (11 - lv))) * 256 + (L.gbase() + slot) * 8);
}

// f node of the shared-memory level lv (S < lv <= 8) from level lv-1.  coop (bit 0): one path, slot 0, one element
// pair per lane; otherwise the lane's own slot, the whole node.
template <int S>
__device__ __forceinline__ void f_level(Lane& L, int lv, bool coop)
{
    const int s = 1 << (10 - lv);
    const int slot = coop ? 0 : L.p();
    const uint32_t off = coop ? (uint32_t)(2 * L.p() * 256) : 0u;
    const uint32_t src = slvl_a<S>(L, lv - 1, slot) + off;
    f_loop(src, src + s * 256, slvl_a<S>(L, lv, slot) + off, coop ? ((2 * L.p() < s) ? 2 : 0) : s, L.tab());
    L.ptr = (L.ptr & ~(7u << (3 * (lv - 1)))) | ((uint32_t)slot << (3 * (lv - 1)));
}

// g over one shared-memory level l0 (S..8): dst[k] = par[k+s] +- par[k], sign from the left-child partial sums in the
// bs register; the parent (slot ps of level l0-1) is the DRAM-resident level S-1 when l0 == S.  All loads first.
template <int S>
__device__ __forceinline__ void g_level(Lane& L, int l0)
{
    const int s = 1 << (10 - l0);                              // 16, 8 or 4
    const int ps = (L.ptr >> (3 * (l0 - 2))) & 7;
    double* dst = slvl<S>(L, l0, L.p());
    const uint32_t bits = L.bs >> s;
#pragma unroll 1
    for (int k0 = 0; k0 < s; k0 += 4) {
        double va[4], vb[4];
        if (l0 == S) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                va[kk] = *elem_ptr<S>(L, S - 1, ps, k0 + kk);
                vb[kk] = *elem_ptr<S>(L, S - 1, ps, k0 + kk + s);
            }
        } else {
            const double* pa = slvl<S>(L, l0 - 1, ps);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) { va[kk] = pa[(k0 + kk) * 32]; vb[kk] = pa[(k0 + kk + s) * 32]; }
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) dst[(k0 + kk) * 32] = gcomb(va[kk], vb[kk], (bits >> (k0 + kk)) & 1u);
    }
    L.ptr = (L.ptr & ~(7u << (3 * (l0 - 1)))) | ((uint32_t)L.p() << (3 * (l0 - 1)));
}

// level-0 copy of one codeword, this lane's share (positions p, p+8, ...): x -> -x.  Out of line: runs once per decode.
__device__ __noinline__ void negate_level0(double* l0)
{
#pragma unroll 4
    for (int k = 0; k < 128; ++k) l0[k * 32] = -l0[k * 32];
}

// ---------------------------------------------------------------------------------------------
// passes over the DRAM-resident levels (0..S-1).  Lane 0 streams the source node into the warp's ring with
// cp.async.bulk (TMA), one 8-row chunk per stage and two stages ahead, completion on one mbarrier per stage;
// every lane reads its own (codeword, slot) column of the landed rows.  Results go straight to their level
// (each store instruction writes one full 256-byte row).  The row pitch RS of the source is a template
// parameter (256 bytes; 32 for level 0, [position][codeword]) so that the eight loads of a chunk are one address
// plus immediates; the four phi chains of a chunk are inlined (fcomb2_inl), so nothing is called inside the loops.
// ---------------------------------------------------------------------------------------------
#ifndef ES_SCL_PASS_INLINE
#define ES_SCL_PASS_INLINE 1     // the two ring passes inlined into the kernel body (single call site each)
#endif
#if ES_SCL_PASS_INLINE
#define ES_PASS_INLINE __forceinline__
#else
#define ES_PASS_INLINE __noinline__
#endif

// lane 0: queue one stage
__device__ __forceinline__ void ring_issue(const Lane& L, int st, const void* src, uint32_t bytes, uint64_t pol)
{
    const uint32_t bar = L.rbar() + 8u * st;
    mbar_expect_tx(bar, bytes);
#if ES_SCL_L2HINT
    bulk_g2s_hint(L.ring(st), src, bytes, bar, pol);
#else
    bulk_g2s(L.ring(st), src, bytes, bar);
#endif
}

// All lanes hold the eight values of the current stage in registers: the stage may be refilled.  A plain
// __syncwarp() is not enough — the loads above are only ISSUED at that point, and a bulk copy queued right
// behind them was measured to overtake them (about one corrupted group in 3000).  A warp vote on the loaded
// words cannot issue before every lane's loads have returned; the branch it feeds is practically never taken
// and harmless when it is.
__device__ __forceinline__ void ring_release(int w)      // w: xor of the high words of the values just loaded
{
    if (__any_sync(0xffffffffu, w == 0x7ff80123)) __nanosleep(2000);
}
__device__ __forceinline__ int hi4(double a, double b, double c, double d)
{
    return __double2hiint(a) ^ __double2hiint(b) ^ __double2hiint(c) ^ __double2hiint(d);
}

// Ring state of one pass.  Every pass consumes an even number of chunks, so both stages always sit at the same
// mbarrier parity between passes: one bit (ph) is carried from pass to pass.
template <int RS> struct Ring {
    const unsigned char* next;   // global address of the next chunk to queue (warp-uniform)
    uint32_t col0, col1;         // shared-window address of this lane's column in stage 0 / 1
    uint32_t bar;                // mbarrier of stage 0 (stage 1: + 8)
    uint32_t ph;                 // parity the next wait has to see
    int st;                      // stage of the next chunk
    int left;                    // chunks not queued yet
    uint64_t pol;                // L2 policy of the source node (ES_SCL_L2HINT)

    __device__ __forceinline__ void begin(const Lane& L, const unsigned char* base, uint32_t col, int nch, uint32_t ph0, int srclvl)
    {
        pol = 0;
#if ES_SCL_L2HINT
        pol = l2_policy((srclvl >= 1 && srclvl <= 2) ? 1 : ((ES_SCL_L2HINT >= 2 && srclvl >= 3) ? 2 : 0));
#endif
        col0 = L.ring(0) + col; col1 = L.ring(1) + col; bar = L.rbar(); ph = ph0; st = 0;
        fence_proxy_async();      // the node was written with ordinary stores (this warp, earlier passes)
        __syncwarp();
        if (L.lane == 0) {
            ring_issue(L, 0, base, 8u * RS, pol);
            ring_issue(L, 1, base + 8u * RS, 8u * RS, pol);
        }
        __syncwarp();
        next = base + 16u * RS;
        left = nch - 2;
    }
    // wait for the next chunk; returns the shared-window address of this lane's column of its row 0
    __device__ __forceinline__ uint32_t wait()
    {
        mbar_wait(bar + 8u * (uint32_t)st, ph);
        return st ? col1 : col0;
    }
    // every lane holds the values of the stage in registers (w covers the loads not consumed yet): refill it
    __device__ __forceinline__ void refill(const Lane& L, int w)
    {
        ring_release(w);
        if (L.lane == 0 && left > 0) ring_issue(L, st, next, 8u * RS, pol);
        next += 8u * RS;
        --left;
        ph ^= (uint32_t)st;
        st ^= 1;
    }
};
template <int RS> __device__ __forceinline__ void ring_ld4(uint32_t at, int i0, double& a, double& b, double& c, double& d)
{
    a = lds_f64(at + (uint32_t)(i0 * RS)); b = lds_f64(at + (uint32_t)((i0 + 1) * RS));
    c = lds_f64(at + (uint32_t)((i0 + 2) * RS)); d = lds_f64(at + (uint32_t)((i0 + 3) * RS));
}

// g node of level l0 (1..S-1) from its parent (slot ps of level l0-1; level 0 = the channel LLRs), fused with the
// f node of level l0+1 below it when do_f: chunk c holds the parent elements k, k+h, k+s, k+s+h for k = 2c, 2c+1
// and gives g[k], g[k+h] (stored: the g node of level l0+1 needs them later) and f(g[k], g[k+h]) — the g node is
// never read back for its f child.
template <int S, int RS>
__device__ __forceinline__ uint32_t pass_gf_body(const Lane& L, int l0, bool do_f)
{
    const int s = 1 << (10 - l0), h = s >> 1;
    const int nch = h >> 1;
    Ring<RS> R;
    if (RS == 32) R.begin(L, reinterpret_cast<const unsigned char*>(L.gw + SclLY::G_ROWS * 32), (uint32_t)(L.lane >> 3) * 8u, nch, L.rphase, 0);
    else R.begin(L, reinterpret_cast<const unsigned char*>(L.gw + (size_t)lvl_row0(l0 - 1) * 32),
                 (uint32_t)(L.gbase() + ((L.ptr >> (3 * (l0 - 2))) & 7)) * 8u, nch, L.rphase, l0 - 1);
#if ES_SCL_L2HINT
    const uint64_t gpol = l2_policy((l0 <= 2) ? 1 : ((ES_SCL_L2HINT >= 2) ? 2 : 0));
    const uint64_t fpol = l2_policy((l0 + 1 <= 2) ? 1 : ((ES_SCL_L2HINT >= 2) ? 2 : 0));
#endif
    // outputs, as running pointers.  g[k] sits at row 8c (c < nch/2) or 8(c - nch/2) + 1, g[k+h] 2 rows, g[k+1] 4 rows,
    // g[k+1+h] 6 rows further.  f[k]: natural order in shared memory (row k, f[k+1] one row further), or quarter-
    // interleaved in global memory: quarter j = c / (nch/4), row 8(c mod nch/4) + j, f[k+1] four rows further.
    double* const gbase = L.ga() + lvl_row0(l0) * 32 + L.p();
    double* gd = gbase;
    const bool nat = (l0 + 1 >= S);
    double* const fbase = nat ? slvl<S>(L, l0 + 1, L.p()) : L.ga() + lvl_row0(l0 + 1) * 32 + L.p();
    double* fd = fbase;
    const int fstep = nat ? 2 * 32 : 8 * 32, f1 = nat ? 32 : 4 * 32;
    const int lnq = 31 - __clz(nch) - 2;                      // log2(nch / 4) >= 1
    const int bmask = ((lnq < 4) ? (1 << lnq) : 16) - 1;      // chunks between two looks at the block below
    // left-child partial sums of the g node: bit k of the lane's beta words; 32 consecutive k per word, h is a
    // multiple of 16: the two words in use are reloaded every 16 chunks only
    const uint32_t* bw = beta_rows(L, l0) + L.gbase() + ((L.bptr >> (3 * (l0 - 1))) & 7);
    uint32_t w0 = 0, w1 = 0;
    const uint32_t tab = L.tab();
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        double v0, v1, v2, v3, v4, v5, v6, v7;     // A(k) A(k+h) B(k) B(k+h) A(k+1) A(k+1+h) B(k+1) B(k+1+h)
        {
            const uint32_t at = R.wait();
            ring_ld4<RS>(at, 0, v0, v1, v2, v3);
            ring_ld4<RS>(at, 4, v4, v5, v6, v7);
            R.refill(L, hi4(v0, v1, v2, v3) ^ hi4(v4, v5, v6, v7));
        }
        if ((c & bmask) == 0) {
            if ((c & 15) == 0) {
                const int k = 2 * c;
                w0 = bw[(k >> 5) * 32];
                w1 = bw[((k + h) >> 5) * 32] >> (h & 31);         // h = 16 (level 5): the upper half of the same word
            }
            if ((c & ((1 << lnq) - 1)) == 0) {
                const int j = c >> lnq;
                if (!nat) fd = fbase + j * 32;
                if (j == 2) gd = gbase + 32;
            }
        }
        const double g0 = gcomb(v0, v2, w0 & 1u), gh0 = gcomb(v1, v3, w1 & 1u);
        const double g1 = gcomb(v4, v6, (w0 >> 1) & 1u), gh1 = gcomb(v5, v7, (w1 >> 1) & 1u);
        w0 >>= 2; w1 >>= 2;
#if ES_SCL_L2HINT
        stg_hint(gd, g0, gpol); stg_hint(gd + 2 * 32, gh0, gpol); stg_hint(gd + 4 * 32, g1, gpol); stg_hint(gd + 6 * 32, gh1, gpol);
#else
        gd[0] = g0; gd[2 * 32] = gh0; gd[4 * 32] = g1; gd[6 * 32] = gh1;
#endif
        gd += 8 * 32;
        if (do_f) {
            double r0, r1;
            fcomb2(g0, gh0, g1, gh1, tab, r0, r1);
#if ES_SCL_L2HINT
            if (!nat) { stg_hint(fd, r0, fpol); stg_hint(fd + f1, r1, fpol); }
            else
#endif
            { fd[0] = r0; fd[f1] = r1; }
            fd += fstep;
        }
    }
    return R.ph;
}

// Out of line on purpose (one copy, and its working set does not add to the register pressure of the
// decode loop): only the words the pass needs travel; the caller re-points the level's slot.
template <int S>
__device__ ES_PASS_INLINE uint32_t pass_gf_fn(uint32_t wsm, double* gw, int lane, uint32_t rphase, uint32_t ptr, uint32_t bptr,
                                            int l0, bool do_f)
{
    Lane T;
    T.wsm = wsm; T.gw = gw; T.lane = lane; T.rphase = rphase; T.ptr = ptr; T.bptr = bptr;
    return (l0 == 1) ? pass_gf_body<S, 32>(T, l0, do_f) : pass_gf_body<S, 256>(T, l0, do_f);
}
template <int S>
__device__ __forceinline__ void pass_gf(Lane& L, int l0, bool do_f)
{
    L.rphase = pass_gf_fn<S>(L.wsm, L.gw, L.lane, L.rphase, L.ptr, L.bptr, l0, do_f);
    uint32_t ptr = (L.ptr & ~(7u << (3 * (l0 - 1)))) | ((uint32_t)L.p() << (3 * (l0 - 1)));
    if (do_f) ptr = (ptr & ~(7u << (3 * l0))) | ((uint32_t)L.p() << (3 * l0));
    L.ptr = ptr;
}

// f node of level lv (3..S) from the node of level lv-1 in the lane's own slot (written by the pass above it): chunk c
// holds the source elements r, r+q, r+2q, r+3q for r = 2c, 2c+1 (q = quarter of the source node) and gives f[r],
// f[r+q], f[r+1], f[r+1+q].
template <int S>
__device__ ES_PASS_INLINE uint32_t pass_f_fn(uint32_t wsm, double* gw, int lane, uint32_t rphase, int lv)
{
    Lane L;
    L.wsm = wsm; L.gw = gw; L.lane = lane;
    const int s = 1 << (10 - lv), q = s >> 1;
    const int nch = s >> 2;
    Ring<256> R;
    R.begin(L, reinterpret_cast<const unsigned char*>(gw + (size_t)lvl_row0(lv - 1) * 32), (uint32_t)lane * 8u, nch, rphase, lv - 1);
    // f[r]: natural order in shared memory (row r; f[r+q] q rows, f[r+1] one row further) or quarter-interleaved in
    // global memory: row 8c (c < nch/2) or 8(c - nch/2) + 1, f[r+q] two rows, f[r+1] four rows further
    const bool nat = lv >= S;
    double* const fbase = nat ? slvl<S>(L, lv, L.p()) : L.ga() + lvl_row0(lv) * 32 + L.p();
    double* fd = fbase;
    const int fstep = nat ? 2 * 32 : 8 * 32, dq = nat ? q * 32 : 2 * 32, d1 = nat ? 32 : 4 * 32;
    const int half = nat ? -1 : (nch >> 1);
    const uint32_t tab = L.tab();
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        // rows of the chunk: e(r) e(r+q) e(r+2q) e(r+3q) | e(r+1) e(r+1+q) e(r+1+2q) e(r+1+3q)
        double v0, v1, v2, v3, r0, r1;
        const uint32_t at = R.wait();
        if (c == half) fd = fbase + 32;
        ring_ld4<256>(at, 0, v0, v1, v2, v3);
        fcomb2(v0, v2, v1, v3, tab, r0, r1);               // f[r], f[r+q]
        fd[0] = r0; fd[dq] = r1;
        ring_ld4<256>(at, 4, v0, v1, v2, v3);
        R.refill(L, hi4(v0, v1, v2, v3));
        fcomb2(v0, v2, v1, v3, tab, r0, r1);               // f[r+1], f[r+1+q]
        fd[d1] = r0; fd[d1 + dq] = r1;
        fd += fstep;
    }
    return R.ph;
}

// The bit-0 spine (one path): f node of level lv (1..S) from level lv-1, the 8 lanes of a codeword sharing the chunks
// of slot 0, read straight from memory.  Runs once per decode.
template <int S>
__device__ __noinline__ void pass_f_coop_fn(uint32_t wsm, double* gw, int lane, int lv)
{
    Lane L;
    L.wsm = wsm; L.gw = gw; L.lane = lane;
    const int s = 1 << (10 - lv), q = s >> 1;
    const int nch = s >> 2;
    const unsigned char* base;
    uint32_t rs, col;
    if (lv == 1) {
        base = reinterpret_cast<const unsigned char*>(gw + SclLY::G_ROWS * 32);
        rs = 32u;
        col = (uint32_t)(lane >> 3) * 8u;
    } else {
        base = reinterpret_cast<const unsigned char*>(gw + (size_t)lvl_row0(lv - 1) * 32);
        rs = 256u;
        col = (uint32_t)L.gbase() * 8u;
    }
    const bool nat = lv >= S;
    double* fdst = nat ? slvl<S>(L, lv, 0) : L.ga() + lvl_row0(lv) * 32;
    const int half = nch >> 1;
#pragma unroll 1
    for (int c = L.p(); c < nch; c += 8) {
        double* fd = nat ? fdst + 2 * c * 32 : fdst + ((c < half) ? 8 * c : 8 * (c - half) + 1) * 32;   // f[r]
        const int dq = nat ? q * 32 : 2 * 32, d1 = nat ? 32 : 4 * 32;                                        // f[r+q], f[r+1]
        const unsigned char* src = base + (size_t)c * 8u * rs + col;
        double v[8], r0, r1;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const double*>(src + (size_t)i * rs);
        fcomb2(v[0], v[2], v[1], v[3], L.tab(), r0, r1);       // f[r], f[r+q]
        fd[0] = r0; fd[dq] = r1;
        fcomb2(v[4], v[6], v[5], v[7], L.tab(), r0, r1);       // f[r+1], f[r+1+q]
        fd[d1] = r0; fd[d1 + dq] = r1;
    }
}
template <int S>
__device__ __forceinline__ void pass_f(Lane& L, int lv, bool coop)
{
    if (coop) pass_f_coop_fn<S>(L.wsm, L.gw, L.lane, lv);
    else L.rphase = pass_f_fn<S>(L.wsm, L.gw, L.lane, L.rphase, lv);
    L.ptr = (L.ptr & ~(7u << (3 * (lv - 1)))) | ((uint32_t)(coop ? 0 : L.p()) << (3 * (lv - 1)));
}

// levels 1..8 for the quad starting at bit i (i % 4 == 0): one g node, then f nodes down to level `last`.
// Bit 0 has no g node and only one path: the 8 lanes of the codeword share the f chain of slot 0 (coop).
// Levels 9 and 10 never touch memory: the quad routine in the kernel keeps them in registers.
template <int S>
__device__ __forceinline__ void llr_update8(Lane& L, int i, int last)   // l0 <= last <= 8
{
    const bool coop = (i == 0);
    int lv = 1;
    if (!coop) {
        const int l0 = 11 - __ffs(i);    // <= 8
        if (l0 < S) {
            const bool fuse = last > l0;
            pass_gf<S>(L, l0, fuse);
            lv = l0 + (fuse ? 2 : 1);
        } else {
            g_level<S>(L, l0);
            lv = l0 + 1;
        }
    }
#pragma unroll 1
    for (; lv <= last && lv <= S; ++lv) {
        pass_f<S>(L, lv, coop);
        if (coop) __syncwarp();
    }
#pragma unroll 1
    for (; lv <= last; ++lv) {
        f_level<S>(L, lv, coop);
        if (coop) __syncwarp();
    }
    if (coop) L.ptr = 0;
}

// Rate-0 node: all `count` (multiple of 4) bits below the node are frozen, so every path's decisions there
// are zeros and its metric grows by -ln P(x = 0 | node LLRs) = sum_k ln(1 + exp(a_k)) -- the same quantity the
// leaf-by-leaf walk (rtwm/fastpolar.py:305-312) accumulates, summed in another order (DESIGN.md section 4,
// "tie contract").  Four interleaved partial sums over the elements in natural order, combined as
// (s0 + s1) + (s2 + s3).  a = position 0 of the node in the lane's slot; lq >= 0: quarter-interleaved node
// with quarters of 2^lq elements (>= 8), lq < 0: natural order.
__device__ __noinline__ double r0_sum(const double* a, int lq, int count, uint32_t tab)
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const int step = (lq < 0) ? 32 : 128;
#pragma unroll 1
    for (int k = 0; k < count; k += 4) {
        const double* e = a + ((lq < 0) ? k : (((k & ((1 << lq) - 1)) << 2) | (k >> lq))) * 32;
        const double a0 = e[0], a1 = e[step], a2 = e[2 * step], a3 = e[3 * step];
#if ES_SCL_F_PSI
        const D4 P = psi4(a0, a1, a2, a3);            // ln(1 + e^a) = a/2 + psi(a)
        s0 += __fma_rn(a0, 0.5, P.a);
        s1 += __fma_rn(a1, 0.5, P.b);
        s2 += __fma_rn(a2, 0.5, P.c);
        s3 += __fma_rn(a3, 0.5, P.d);
#else
        const D4 P = phi4(a0, a1, a2, a3);            // ln(1 + e^a) = max(a, 0) + phi(|a|)
        s0 += P.a + fmax(a0, 0.0);
        s1 += P.b + fmax(a1, 0.0);
        s2 += P.c + fmax(a2, 0.0);
        s3 += P.d + fmax(a3, 0.0);
#endif
    }
    return (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ int nth_set8(uint32_t mask, int n)
{
    int r = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool set = (mask >> b) & 1u;
        if (set && n == 0) r = b;
        n -= set ? 1 : 0;
    }
    return r;
}

// quad-local values that must follow a path through a clone
struct Carry { double c0, c1, c2, c3; uint32_t qb; };

// information-bit step: rank the 2*np candidates, keep list_size, clone into free lanes.
// pen0/pen1 = penalties of deciding 0 / 1 (rtwm/fastpolar.py:32-40).  Metrics are non-negative finite
// doubles, so their bit patterns order like unsigned integers; the reference's stable tie-break
// (candidate index 2*ord+bit) folds into the comparison as  (kj < k) + (kj == k && c) == (kj < k + c).
// Ranking of the 16 candidates of a codeword by (metric, path order, bit) = the reference's stable sort over candidates
// appended in (path, 0, 1) order (rtwm/fastpolar.py:288-299), as ONE unsigned 64-bit key per candidate.  A path metric is a sum of penalties that are each 0 or
// >= 2^-54 (phi_fast is exactly 0 or >= 2^-53; |leaf| is only added on top of phi), so it is 0 or >= 2^-54 and < 2^64: its
// exponent field spans fewer than 128 values above 960 and the bit pattern, rebased there, leaves four low bits for the
// candidate index.  Keys are unique, so the rank is a plain count of smaller keys.
__device__ __forceinline__ unsigned long long rank_key(double m, uint32_t cand)
{
    uint32_t hi = (uint32_t)__double2hiint(m), lo = (uint32_t)__double2loint(m);
    hi = max(hi, 0x3C000000u) - 0x3C000000u;                    // exact zero stays zero (its low word is zero)
    return ((((unsigned long long)hi << 32) | lo) << 4) | cand;
}

// carry: the quad-local values have to follow a clone only after an EVEN leaf (the odd leaf right after it reads
// c0, c1 and - a clone always took bit 1 - c2); after an odd leaf nothing of them is read again.
template <bool MG>
__device__ __forceinline__ int info_step(Lane& L, int list_size, double pen0, double pen1, Carry& cy, bool carry)
{
    const unsigned full = 0xffffffffu;
    const double m0 = L.m + pen0, m1 = L.m + pen1;
    int r0 = 0, r1 = 0;
    {
        const unsigned long long k0 = L.active ? rank_key(m0, 2u * (uint32_t)L.ord) : ~0ull;
        const unsigned long long k1 = L.active ? rank_key(m1, 2u * (uint32_t)L.ord + 1u) : ~0ull;
        // exchange through the warp's stash area (idle outside the LLR update): 16 bytes per lane, broadcast reads
        const uint32_t xa = L.wsm + (uint32_t)SclLY::STASH_OFF;
        asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(xa + (uint32_t)L.lane * 16u), "l"(k0), "l"(k1) : "memory");
        __syncwarp();
        const uint32_t xg = xa + (uint32_t)L.gbase() * 16u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            unsigned long long k0j, k1j;
            asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(k0j), "=l"(k1j) : "r"(xg + 16u * j) : "memory");
            r0 += (int)(k0j < k0) + (int)(k1j < k0);
            r1 += (int)(k0j < k1) + (int)(k1j < k1);
        }
    }
    if (MG) {
        // prune margin (rtwm/fastpolar.py:288-299 sorts the 2|P| candidates and keeps L): relative gap between rank L-1
        // and rank L, minimum over the decode.  gap/den is compared by cross-multiplication; one division at the end.
        const bool hasA = L.active && (r0 == list_size - 1 || r1 == list_size - 1);
        const bool hasB = L.active && (r0 == list_size || r1 == list_size);
        const double vA = (r0 == list_size - 1) ? m0 : m1, vB = (r0 == list_size) ? m0 : m1;
        const uint32_t bA = (__ballot_sync(full, hasA) >> L.gbase()) & 0xffu;
        const uint32_t bB = (__ballot_sync(full, hasB) >> L.gbase()) & 0xffu;
        const double mA = __shfl_sync(full, vA, __ffs(bA) - 1, 8), mB = __shfl_sync(full, vB, __ffs(bB) - 1, 8);
        if (bB) {        // more than L candidates: this step prunes
            const double gap = mB - mA;
            if (gap * L.mg_den < L.mg_gap * mB) { L.mg_gap = gap; L.mg_den = mB; }
        }
    }
    const bool s0 = L.active && (r0 < list_size);
    const bool s1 = L.active && (r1 < list_size);
    const uint32_t cm = (__ballot_sync(full, s0 && s1) >> L.gbase()) & 0xffu;
    const uint32_t fm = (__ballot_sync(full, !(s0 || s1)) >> L.gbase()) & 0xffu;
    // clone source for free lanes (j-th free lane takes the j-th clone)
    const int jfree = __popc(fm & ((1u << L.p()) - 1u));
    const bool take = !(s0 || s1) && (jfree < __popc(cm));
    int src = L.p();                 // clone source: the jfree-th set bit of the clone mask, from a 2 KB shared table
    if (take) {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(smem_base() + (uint32_t)SclLY::NTH_OFF + cm * 8u + (uint32_t)jfree) : "memory");
        src = (int)v;
    }

    const double cm1 = __shfl_sync(full, m1, src, 8);
    const int cr1 = __shfl_sync(full, r1, src, 8);
    const uint32_t cptr = __shfl_sync(full, L.ptr, src, 8);
    const uint32_t cbptr = __shfl_sync(full, L.bptr, src, 8);
    const uint32_t cbs = __shfl_sync(full, L.bs, src, 8);
    const uint32_t cqb = __shfl_sync(full, cy.qb, src, 8);
    if (carry) {
        const double q0 = __shfl_sync(full, cy.c0, src, 8), q1 = __shfl_sync(full, cy.c1, src, 8);
        const double q2 = __shfl_sync(full, cy.c2, src, 8);
        if (take) { cy.c0 = q0; cy.c1 = q1; cy.c2 = q2; }
    }
    int bit = 0;
    if (s0) { L.m = m0; L.ord = r0; bit = 0; }
    else if (s1) { L.m = m1; L.ord = r1; bit = 1; }
    else if (take) {
        L.m = cm1; L.ord = cr1; L.ptr = cptr; L.bptr = cbptr; L.bs = cbs; bit = 1; L.active = true;
        cy.qb = cqb;
    } else { L.active = false; }
    return bit;
}

// one decision (frozen or information bit) given the leaf LLR and phi(|leaf|)
template <bool MG>
__device__ __forceinline__ int decide(Lane& L, bool frozen, double leaf, double ph, int list_size, Carry& cy, bool carry)
{
    const double al = fabs(leaf);
    const bool pref1 = (leaf >= 0.0);
    const double pen0 = pref1 ? (ph + al) : ph;    // deciding 0 against a non-negative LLR costs |l| more
    const double pen1 = pref1 ? ph : (ph + al);
    if (frozen) {
        if (L.active) L.m += pen0;
        return 0;
    }
    return info_step<MG>(L, list_size, pen0, pen1, cy, carry);
}

// partial-sum update after a quad (bits 4q..4q+3): X = the 4 partial sums of the finished level-8 node
// (see tests/model_scl_lanes.py for the bit-level model of the chain)
__device__ __forceinline__ void beta_update_quad(Lane& L, int q, uint32_t X, uint32_t* xroot)
{
    if ((q & 1) == 0) { L.bs = (L.bs & ~0xf0u) | (X << 4); return; }      // level-8 left child
    const int t1 = __ffs(~q) - 1;            // trailing ones of q, 1..8
    const int tr = t1 < 3 ? t1 : 3;
    for (int j = 0; j < tr; ++j) {           // levels 8, 7, 6 live in the bs register
        const int s = 4 << j;
        const uint32_t Lb = (L.bs >> s) & ((1u << s) - 1u);
        X = (Lb ^ X) | (X << s);
    }
    if (t1 <= 2) {
        const int s = 4 << t1;
        const uint32_t mask = (s == 16) ? 0xffff0000u : (((1u << s) - 1u) << s);
        L.bs = (L.bs & ~mask) | (X << s);
    } else if (t1 == 3) {
        L.sb()[L.lane] = X;                       // level 5, word offset 0
        L.bptr = (L.bptr & ~(7u << 12)) | ((uint32_t)L.p() << 12);
    } else {
        const int lstar = 8 - t1;               // 4..0
        uint32_t* D = (lstar == 0) ? (xroot + L.lane) : (beta_rows(L, lstar) + L.lane);
        D[0] = X;
        int n = 1;
#pragma unroll 1                                 // one quad in sixteen gets here: keep it small, not unrolled
        for (int l = 5; l > lstar; --l) {
            const int bsl = (L.bptr >> (3 * (l - 1))) & 7;
            const uint32_t* Lp = beta_rows(L, l) + L.gbase() + bsl;
#pragma unroll 1
            for (int w = 0; w < n; ++w) {
                const uint32_t x = D[w * 32];
                D[(n + w) * 32] = x;
                D[w * 32] = x ^ Lp[w * 32];
            }
            n <<= 1;
        }
        if (lstar > 0) L.bptr = (L.bptr & ~(7u << (3 * (lstar - 1)))) | ((uint32_t)L.p() << (3 * (lstar - 1)));
    }
}

__device__ __forceinline__ uint32_t xform_word(uint32_t x)
{
    x ^= (x >> 1) & 0x55555555u;
    x ^= (x >> 2) & 0x33333333u;
    x ^= (x >> 4) & 0x0f0f0f0fu;
    x ^= (x >> 8) & 0x00ff00ffu;
    x ^= (x >> 16) & 0x0000ffffu;
    return x;
}

__device__ __forceinline__ uint8_t crc8_step_bit(uint8_t reg, uint32_t bit)
{
    reg ^= (uint8_t)(bit << 7);
    return (reg & 0x80) ? (uint8_t)((reg << 1) ^ 0x07) : (uint8_t)(reg << 1);
}

// ---------------------------------------------------------------------------------------------
// list decoder kernel: W warps per CTA share the phi tables; each warp decodes 4 codewords at a time
// ---------------------------------------------------------------------------------------------
template <int S, int W, bool MG>
__global__ void __maxnreg__(ES_SCL_MAXNREG) scl_list_kernel(SclParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using LY = SclLayout<S>;
    const unsigned full = 0xffffffffu;
    // phi tables -> shared
    {
        double* st = reinterpret_cast<double*>(smem_raw);
        for (int q = threadIdx.x; q < PHI_TAB_DOUBLES; q += W * 32) st[q] = P.phi_tab[q];
        for (int q = threadIdx.x; q < 2048; q += W * 32) smem_raw[LY::NTH_OFF + q] = (unsigned char)nth_set8((uint32_t)(q >> 3), q & 7);
        for (int q = threadIdx.x; q < 256; q += W * 32) smem_raw[LY::CRC_OFF + q] = c_crc8[q];
    }
    const int warp = threadIdx.x >> 5;
    Lane L;
    L.lane = threadIdx.x & 31;
    L.wsm = smem_base() + (uint32_t)LY::TAB_BYTES + (uint32_t)warp * (uint32_t)LY::WARP_BYTES;
    L.rphase = 0;
    {   // the warp's staging ring: one mbarrier per stage, a single arrival (lane 0's expect_tx) per phase
        if (L.lane == 0) {
#pragma unroll
            for (int st = 0; st < RING_STAGES; ++st) mbar_init(L.rbar() + 8u * st, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        fence_proxy_async();
        __syncthreads();
    }
    L.gw = P.scratch + (size_t)(blockIdx.x * W + warp) * P.scratch_stride;
    // per-lane (metric, bptr, ord|active) saved at bit 512 for the second decode of a +/- pair
#define snap (reinterpret_cast<uint4*>(__cvta_shared_to_generic(L.wsm + LY::ABYTES + LY::BROWS_S * 128)) + L.lane)
#define xroot (reinterpret_cast<uint32_t*>(__cvta_shared_to_generic(L.wsm)))   /* root partial sums: aliases alpha (dead by then) */
#define gscr (L.gw)

#pragma unroll 1
    for (int grp = blockIdx.x * W + warp; grp < ((P.nunits + 3) >> 2); grp += gridDim.x * W) {
        const int j = grp * 4 + (L.lane >> 3);
        bool valid = j < P.nunits;
        const int jj = valid ? j : (P.nunits - 1);
        int w = P.pair ? 2 * jj : (P.index ? P.index[jj] : jj);
        {   // level 0: widen this warp's 4 rows to double, [position][4] (quarter-interleaved like the levels below).  Always +row: f(-a,-b) = f(a,b), so the first
            // half of the tree of -row is that of +row; the copy is negated at bit 512 (below) for the second half.
            const int row = P.neg_mode ? (w >> 1) : w;
            const float* src = P.llr + (size_t)row * 1024;
            double* dst = gscr + LY::G_ROWS * 32 + (L.lane >> 3);
            if ((reinterpret_cast<uintptr_t>(P.llr) & 15u) == 0) {          // 16-byte loads, eight in flight per lane
                const float4* src4 = reinterpret_cast<const float4*>(src);
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const int k = 4 * (L.p() + 8 * i);
                    const float4 v = __ldg(src4 + (k >> 2));
                    double* d0 = dst + ipos(0, k) * 4;                      // k .. k+3 share their 256-block: positions 4 apart
                    d0[0] = (double)v.x; d0[16] = (double)v.y; d0[32] = (double)v.z; d0[48] = (double)v.w;
                }
            } else {
#pragma unroll 4
                for (int k = L.p(); k < 1024; k += 8) dst[ipos(0, k) * 4] = (double)__ldg(src + k);
            }
        }
        L.m = 0.0; L.ptr = 0; L.bptr = 0; L.bs = 0; L.ord = 0;
        L.mg_gap = CUDART_INF; L.mg_den = 1.0;
        L.active = (L.p() == 0);
        L.neg = P.neg_mode && (w & 1);
        __syncwarp();

        int qfirst = 0;
#pragma unroll 1
        for (int pass = 0; ; ++pass) {
#pragma unroll 1
        for (int q = qfirst; q < 256; ++q) {
            const int i = q << 2;
            if (q == 128 && P.pair && pass == 0) {
                // bit 512: everything the second half reads from the first half is the level-1 partial sums
                // (global rows, never rewritten) plus these per-lane words
                *snap = make_uint4((uint32_t)__double2loint(L.m), (uint32_t)__double2hiint(L.m), L.bptr,
                                   ((uint32_t)L.ord << 1) | (L.active ? 1u : 0u));
                if (MG) *reinterpret_cast<double2*>(snap + 32) = make_double2(L.mg_gap, L.mg_den);
            }
            if (q == 128) {
                if (L.neg) negate_level0(gscr + LY::G_ROWS * 32 + (L.lane >> 3) + L.p() * 4);
                __syncwarp();
            }
            // rate-0 nodes: the first quad adds the whole node's penalty, every quad feeds zeros upward
            const int r0 = c_r0[q];
            const int last = (r0 == 0) ? 8 : (9 - r0);
            if (r0 != 255) {                       // (quad 0 is never a rate-0 node)
                // The update below is where the register pressure peaks (pass state + four interleaved phi chains):
                // the words it does not touch wait in shared memory meanwhile.
                const uint32_t stash = L.wsm + (uint32_t)LY::STASH_OFF + (uint32_t)L.lane * 48u;
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(stash), "d"(L.m), "d"(L.mg_gap) : "memory");
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(stash + 16u), "d"(L.mg_den),
                             "d"(__hiloint2double((int)L.bs, L.ord)) : "memory");
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(stash + 32u), "r"(w),
                             "r"((valid ? 1 : 0) | (pass << 1) | (qfirst << 2) | ((int)L.active << 10) | ((int)L.neg << 11)) : "memory");
                llr_update8<S>(L, i, last);
                {
                    int t2;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w), "=r"(t2) : "r"(stash + 32u) : "memory");
                    valid = (t2 & 1) != 0; pass = (t2 >> 1) & 1; qfirst = (t2 >> 2) & 255;
                    L.active = ((t2 >> 10) & 1) != 0; L.neg = ((t2 >> 11) & 1) != 0;
                }
                {
                    double t;
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(L.m), "=d"(L.mg_gap) : "r"(stash) : "memory");
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(L.mg_den), "=d"(t) : "r"(stash + 16u) : "memory");
                    L.bs = (uint32_t)__double2hiint(t);
                    L.ord = __double2loint(t);
                }
                __syncwarp();
            }
            Carry cy; cy.qb = 0;
            if (r0 != 0) {
                if (r0 != 255) {
                    const double* nd = (last >= S) ? slvl<S>(L, last, L.p()) : L.ga() + lvl_row0(last) * 32 + L.p();
                    const double pen = r0_sum(nd, (last >= S) ? -1 : 8 - last, 4 << (r0 - 1), L.tab());
                    if (L.active) L.m += pen;
                }
            } else {
            const uint32_t fz = (c_frozen[i >> 5] >> (i & 31)) & 15u;       // frozen flags of bits i..i+3
            double fm = 0.0, fp = 0.0;
            int prev = 0;
#pragma unroll 1
            for (int t = 0; t < 4; ++t) {
                double leaf, ph;
                if ((t & 1) == 0) {
                    // level-9 node t/2 of the current level-8 node, read through the (possibly re-pointed) slot
                    const double* l8 = slvl<S>(L, 8, (L.ptr >> 21) & 7);
                    const double a0 = l8[0], a1 = l8[32], a2 = l8[64], a3 = l8[96];
                    double x0, x1;
                    if (t == 0) {
                        fcomb2(a0, a2, a1, a3, L.tab(), x0, x1);
                    } else {
                        const uint32_t u0 = cy.qb & 1u, u1 = (cy.qb >> 1) & 1u;
                        x0 = (u0 ^ u1) ? (a2 - a0) : (a2 + a0);
                        x1 = u1 ? (a3 - a1) : (a3 + a1);
                    }
                    // even leaf: f of the pair; its two phi terms are the odd leaf's penalty terms
                    leaf = fcomb_parts(x0, x1, L.tab(), fm, fp);
                    ph = phi1(leaf);
                    cy.c0 = x0; cy.c1 = x1; cy.c2 = fm; cy.c3 = fp;
                } else {
                    // odd leaf: g of the pair
                    leaf = prev ? (cy.c1 - cy.c0) : (cy.c1 + cy.c0);
                    ph = prev ? cy.c2 : cy.c3;
                }
                prev = decide<MG>(L, (fz >> t) & 1u, leaf, ph, P.list_size, cy, (t & 1) == 0);
                cy.qb |= (uint32_t)prev << t;
            }
            }
            uint32_t X = cy.qb;
            X ^= (X >> 1) & 0x5u;
            X ^= (X >> 2) & 0x3u;
            beta_update_quad(L, q, X, xroot);
            __syncwarp();
        }

        // ---- output: rank by (metric, order) = sorted(paths, key=metric) (rtwm/fastpolar.py:335)
        int rank = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int aj = __shfl_sync(full, (int)L.active, q, 8);
            const double mj = __shfl_sync(full, L.m, q, 8);
            const int oj = __shfl_sync(full, L.ord, q, 8);
            if (aj && (mj < L.m || (mj == L.m && oj < L.ord))) ++rank;
        }
        const uint32_t am = (__ballot_sync(full, L.active) >> L.gbase()) & 0xffu;
        const int np = __popc(am);
        if (!L.active) rank = np + __popc((~am & 0xffu) & ((1u << L.p()) - 1u));

        // u-hat = transform(x-hat), in registers
        uint32_t x[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) x[q] = xform_word(xroot[q * 32 + L.lane]);
#pragma unroll
        for (int h = 1; h < 32; h <<= 1)
#pragma unroll
            for (int q = 0; q < 32; ++q)
                if (!(q & h)) x[q] ^= x[q + h];
#pragma unroll
        for (int q = 0; q < 32; ++q) xroot[q * 32 + L.lane] = x[q];

        const bool wr = valid && (rank < P.list_size);
        const size_t orow = (size_t)w * P.list_size + (size_t)(wr ? rank : 0);
        const int K = c_K;
        const int nbytes = (K - 8) >> 3;
        uint8_t* out = P.path_payload + orow * nbytes;
        uint32_t crcreg = 0, crcbits = 0;
        {
            // The K un-frozen bits of u-hat in ascending position are the payload + CRC stream, first bit = MSB of byte 0
            // (rtwm/fastpolar.py:340-349).  Per 32-bit word: compress the un-frozen bits to the low end (five shift steps
            // with masks precomputed by es_polar_set_code), append them to a 64-bit queue, drain whole bytes.
            unsigned long long acc = 0;
            int fill = 0, nb = 0;
            const uint32_t crc_tab = smem_base() + (uint32_t)LY::CRC_OFF;
#pragma unroll 1
            for (int j = 0; j < 32; ++j) {
                uint32_t v = xroot[j * 32 + L.lane] & ~c_frozen[j];
#pragma unroll
                for (int s5 = 0; s5 < 5; ++s5) {
                    const uint32_t tt = v & c_cmp[j][s5];
                    v = (v ^ tt) | (tt >> (1 << s5));
                }
                acc |= (unsigned long long)v << fill;
                fill += c_cnt[j];
#pragma unroll 1
                while (fill >= 8) {
                    const uint32_t byte = __brev((uint32_t)acc) >> 24;
                    acc >>= 8; fill -= 8;
                    if (nb < nbytes) {
                        uint32_t t8;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t8) : "r"(crc_tab + (crcreg ^ byte)) : "memory");
                        crcreg = t8;
                        if (wr && L.active) out[nb] = (uint8_t)byte;
                    } else {
                        crcbits = byte;
                    }
                    ++nb;
                }
            }
        }
        if (wr) {
            P.path_crc[orow] = (L.active && crcreg == crcbits) ? 1 : 0;
            P.path_metric[orow] = L.active ? L.m : CUDART_INF;
        }
        if (valid && L.p() == 0) {
            P.npaths[w] = np;
            if (MG) P.min_margin[w] = L.mg_gap / L.mg_den;      // +inf when no step pruned
        }
        __syncwarp();
        if (!P.pair || pass == 1) break;
        // second pass: the sign-flipped variant of the same row, restarted at bit 512
        const uint4 sn = *snap;
        L.m = __hiloint2double((int)sn.y, (int)sn.x);
        L.bptr = sn.z;
        L.ord = (int)(sn.w >> 1);
        L.active = (sn.w & 1u) != 0u;
        if (MG) { const double2 mg = *reinterpret_cast<const double2*>(snap + 32); L.mg_gap = mg.x; L.mg_den = mg.y; }
        L.neg = true;
        w += 1;
        qfirst = 128;
        }
    }
}

#undef snap
#undef xroot
#undef gscr

// ---------------------------------------------------------------------------------------------
// wide-list decoder: list sizes 9..32 (rtwm/detector.py:27 accepts any list_size; its quick test uses 32).  Same
// arithmetic and rules as scl_list_kernel (f as a psi difference, rate-0 node sums, the -row variant decoded from +row
// with the level-1 g node negated, stable ranking by (metric, path order, bit)), in the plain form of the device-
// arithmetic model (oracle/polar_oracle.c): one WARP per codeword, one lane per path, every tree level in a global
// scratch [element][lane], bit-by-bit leaves, partial sums as bytes.  Lazy copy as in the fast kernel: every path
// rewrites a level at the same bit index into its OWN slot, so a clone copies two 64-bit words of per-level slot
// pointers, its 1024 decision bits and the metric.  Built for exactness, not speed (about a tenth of the SCL-8 rate).
// ---------------------------------------------------------------------------------------------
constexpr int WIDE_W = 8;                                  // warps per CTA
constexpr size_t WIDE_A = 1024u * 32u;                     // doubles: alpha [AOFF(l) + k][lane]
constexpr size_t WIDE_SCRATCH_BYTES = WIDE_A * 8 + 2 * 1024u * 32u;      // + bl bytes [AOFF(l) + k][lane] + tmp bytes [k][lane]

struct WideParams {
    const float* llr; int ncw; int neg_mode; int list_size;
    unsigned char* scratch;
    const double* phi_tab;
    uint8_t* path_payload; uint8_t* path_crc; double* path_metric; int32_t* npaths;
};

__device__ __forceinline__ int wide_aoff(int l) { return 1024 - (1 << (11 - l)); }
__device__ __forceinline__ int wptr_get(unsigned long long w, int l) { return (int)((w >> (5 * (l - 1))) & 31ull); }
__device__ __forceinline__ unsigned long long wptr_set(unsigned long long w, int l, int v)
{
    return (w & ~(31ull << (5 * (l - 1)))) | ((unsigned long long)v << (5 * (l - 1)));
}

struct WideLane {
    double* A; uint8_t* B; uint8_t* T; const float* row; uint32_t* U;      // U: decision bits [word][lane] in shared memory
    int lane; bool neg;
    unsigned long long aptr, bptr;
    double m; int ord; bool active;
};

// LLR update for bit i down to level `last` (oracle sc_step_llr_to)
__device__ __noinline__ unsigned long long wide_step(double* A, const uint8_t* B, const float* row, int lane, bool neg,
                                                    unsigned long long aptr, unsigned long long bptr, int i, int last)
{
    const int l0 = (i == 0) ? 0 : 10 - (__ffs(i) - 1);
    if (l0 >= 1) {
        const int s = 1 << (10 - l0);
        const int ps = (l0 == 1) ? 0 : wptr_get(aptr, l0 - 1);
        const double* par = A + (size_t)wide_aoff(l0 - 1) * 32 + ps;            // unused for l0 == 1
        double* dst = A + (size_t)wide_aoff(l0) * 32 + lane;
        const uint8_t* bl = B + (size_t)wide_aoff(l0) * 32 + wptr_get(bptr, l0);
#pragma unroll 2
        for (int k = 0; k < s; ++k) {
            const double a = (l0 == 1) ? (double)__ldg(row + k) : par[(size_t)k * 32];
            const double b = (l0 == 1) ? (double)__ldg(row + k + s) : par[(size_t)(k + s) * 32];
            double g = gcomb(a, b, bl[(size_t)k * 32]);
            if (l0 == 1 && neg) g = -g;
            dst[(size_t)k * 32] = g;
        }
        aptr = wptr_set(aptr, l0, lane);
    }
    for (int l = l0 + 1; l <= last; ++l) {
        const int s = 1 << (10 - l);
        const int ps = (l == 1) ? 0 : wptr_get(aptr, l - 1);
        const double* par = A + (size_t)wide_aoff(l - 1) * 32 + ps;
        double* dst = A + (size_t)wide_aoff(l) * 32 + lane;
        if (s == 1) {
            const double a = par[0], b = par[32];
            const D4 P = psi4(a - b, a + b, a - b, a + b);
            dst[0] = P.a - P.b;
        } else {
#pragma unroll 1
            for (int k = 0; k < s; k += 2) {
                const double a0 = (l == 1) ? (double)__ldg(row + k) : par[(size_t)k * 32];
                const double b0 = (l == 1) ? (double)__ldg(row + k + s) : par[(size_t)(k + s) * 32];
                const double a1 = (l == 1) ? (double)__ldg(row + k + 1) : par[(size_t)(k + 1) * 32];
                const double b1 = (l == 1) ? (double)__ldg(row + k + 1 + s) : par[(size_t)(k + 1 + s) * 32];
                const D4 P = psi4(a0 - b0, a0 + b0, a1 - b1, a1 + b1);
                dst[(size_t)k * 32] = P.a - P.b;
                dst[(size_t)(k + 1) * 32] = P.c - P.d;
            }
        }
        aptr = wptr_set(aptr, l, lane);
    }
    return aptr;
}

// partial-sum update after deciding `bit` at position i (oracle sc_extend); returns the new bptr
__device__ __noinline__ unsigned long long wide_extend(uint8_t* B, uint8_t* T, int lane, unsigned long long bptr, int i, int bit)
{
    uint8_t* tmp = T + lane;
    tmp[0] = (uint8_t)bit;
    int l = 10, s = 1;
    while (l > 0 && ((i >> (10 - l)) & 1)) {
        const uint8_t* left = B + (size_t)wide_aoff(l) * 32 + wptr_get(bptr, l);
        for (int k = 0; k < s; ++k) {
            const uint8_t t = tmp[(size_t)k * 32];
            tmp[(size_t)(k + s) * 32] = t;
            tmp[(size_t)k * 32] = (uint8_t)(left[(size_t)k * 32] ^ t);
        }
        s <<= 1; --l;
    }
    if (l > 0) {
        uint8_t* dst = B + (size_t)wide_aoff(l) * 32 + lane;
        for (int k = 0; k < s; ++k) dst[(size_t)k * 32] = tmp[(size_t)k * 32];
        bptr = wptr_set(bptr, l, lane);
    }
    return bptr;
}

__global__ void __launch_bounds__(WIDE_W * 32) scl_wide_kernel(WideParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using LY = SclLayout<SCL_S>;
    const unsigned full = 0xffffffffu;
    {
        double* st = reinterpret_cast<double*>(smem_raw);
        for (int q = threadIdx.x; q < PHI_TAB_DOUBLES; q += WIDE_W * 32) st[q] = P.phi_tab[q];
        for (int q = threadIdx.x; q < 256; q += WIDE_W * 32) smem_raw[LY::CRC_OFF + q] = c_crc8[q];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* U = reinterpret_cast<uint32_t*>(smem_raw + LY::TAB_BYTES) + (size_t)warp * 32 * 32;
    unsigned char* scr = P.scratch + (size_t)(blockIdx.x * WIDE_W + warp) * WIDE_SCRATCH_BYTES;
    double* A = reinterpret_cast<double*>(scr);
    uint8_t* B = scr + WIDE_A * 8;
    uint8_t* T = B + 1024u * 32u;
    const int Lsz = P.list_size;
    const int K = c_K;
    const int nbytes = (K - 8) >> 3;
#pragma unroll 1
    for (int w = blockIdx.x * WIDE_W + warp; w < P.ncw; w += gridDim.x * WIDE_W) {
        const int rowi = P.neg_mode ? (w >> 1) : w;
        const bool neg = P.neg_mode && (w & 1);
        const float* row = P.llr + (size_t)rowi * 1024;
        double m = 0.0;
        int ord = 0;
        bool active = (lane == 0);
        unsigned long long aptr = 0, bptr = 0;
        for (int j = 0; j < 32; ++j) U[j * 32 + lane] = 0;
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < 1024; ++i) {
            const int q = i >> 2;
            const int r0 = ((i & 3) == 0) ? c_r0[q] : 0;
            if (r0 != 0 && r0 != 255) {
                // rate-0 node of s = 4 << (r0-1) bits: sum_k ln(1 + e^{a_k}) over the node, zeros decided
                const int s = 4 << (r0 - 1);
                const int lv = 10 - (31 - __clz(s));
                aptr = wide_step(A, B, row, lane, neg, aptr, bptr, i, lv);
                const double* a = A + (size_t)wide_aoff(lv) * 32 + wptr_get(aptr, lv);
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 1
                for (int k = 0; k < s; k += 4) {
                    const double a0 = a[(size_t)k * 32], a1 = a[(size_t)(k + 1) * 32], a2 = a[(size_t)(k + 2) * 32], a3 = a[(size_t)(k + 3) * 32];
                    const D4 Q4 = psi4(a0, a1, a2, a3);
                    s0 += __fma_rn(a0, 0.5, Q4.a); s1 += __fma_rn(a1, 0.5, Q4.b);
                    s2 += __fma_rn(a2, 0.5, Q4.c); s3 += __fma_rn(a3, 0.5, Q4.d);
                }
                if (active) m += (s0 + s1) + (s2 + s3);
                for (int k = 0; k < s; ++k) bptr = wide_extend(B, T, lane, bptr, i + k, 0);
                i += s - 1;
                continue;
            }
            aptr = wide_step(A, B, row, lane, neg, aptr, bptr, i, 10);
            const double leaf = A[(size_t)wide_aoff(10) * 32 + wptr_get(aptr, 10)];
            const double al = fabs(leaf);
            const double ph = phi1(leaf);
            const bool pref1 = (leaf >= 0.0);
            const double pen0 = pref1 ? (ph + al) : ph, pen1 = pref1 ? ph : (ph + al);
            const bool frozen = (c_frozen[i >> 5] >> (i & 31)) & 1u;
            int bit = 0;
            if (frozen) {
                if (active) m += pen0;
            } else {
                const double m0 = m + pen0, m1 = m + pen1;
                const double k0 = active ? m0 : CUDART_INF, k1 = active ? m1 : CUDART_INF;
                int r0c = 0, r1c = 0;
#pragma unroll 4
                for (int j = 0; j < 32; ++j) {
                    const double k0j = __shfl_sync(full, k0, j);
                    const double k1j = __shfl_sync(full, k1, j);
                    const int oj = __shfl_sync(full, ord, j);
                    const bool lt = oj < ord, le = oj <= ord;
                    r0c += (int)((k0j < k0) || (lt && k0j == k0)) + (int)((k1j < k0) || (lt && k1j == k0));
                    r1c += (int)((k0j < k1) || (le && k0j == k1)) + (int)((k1j < k1) || (lt && k1j == k1));
                }
                const bool s0 = active && (r0c < Lsz), s1 = active && (r1c < Lsz);
                const uint32_t cm = __ballot_sync(full, s0 && s1);
                const uint32_t fm = __ballot_sync(full, !(s0 || s1));
                const int jfree = __popc(fm & ((1u << lane) - 1u));
                const bool take = !(s0 || s1) && (jfree < __popc(cm));
                const int src = take ? (int)__fns(cm, 0, jfree + 1) : lane;
                const double cm1 = __shfl_sync(full, m1, src);
                const int cr1 = __shfl_sync(full, r1c, src);
                const unsigned long long ca = __shfl_sync(full, aptr, src), cb = __shfl_sync(full, bptr, src);
                for (int j = 0; j < 32; ++j) {              // decision bits follow the clone
                    const uint32_t uw = U[j * 32 + src];
                    __syncwarp();
                    if (take) U[j * 32 + lane] = uw;
                }
                if (s0) { m = m0; ord = r0c; bit = 0; }
                else if (s1) { m = m1; ord = r1c; bit = 1; }
                else if (take) { m = cm1; ord = cr1; aptr = ca; bptr = cb; bit = 1; active = true; }
                else active = false;
                if (bit) U[(i >> 5) * 32 + lane] |= 1u << (i & 31);
            }
            bptr = wide_extend(B, T, lane, bptr, i, bit);
            __syncwarp();
        }
        // ---- output in ascending (metric, order): sorted(paths, key=metric) (rtwm/fastpolar.py:335)
        int rank = 0;
        for (int j = 0; j < 32; ++j) {
            const int aj = __shfl_sync(full, (int)active, j);
            const double mj = __shfl_sync(full, m, j);
            const int oj = __shfl_sync(full, ord, j);
            if (aj && (mj < m || (mj == m && oj < ord))) ++rank;
        }
        const uint32_t am = __ballot_sync(full, active);
        const int np = __popc(am);
        const bool wr = active && rank < Lsz;
        const size_t orow = (size_t)w * Lsz + (size_t)(wr ? rank : 0);
        uint8_t* out = P.path_payload + orow * nbytes;
        uint32_t crcreg = 0, crcbits = 0;
        {
            unsigned long long acc = 0;
            int fill = 0, nb = 0;
            const uint32_t crc_tab = smem_base() + (uint32_t)LY::CRC_OFF;
#pragma unroll 1
            for (int j = 0; j < 32; ++j) {
                uint32_t v = U[j * 32 + lane] & ~c_frozen[j];
#pragma unroll
                for (int s5 = 0; s5 < 5; ++s5) {
                    const uint32_t tt = v & c_cmp[j][s5];
                    v = (v ^ tt) | (tt >> (1 << s5));
                }
                acc |= (unsigned long long)v << fill;
                fill += c_cnt[j];
#pragma unroll 1
                while (fill >= 8) {
                    const uint32_t byte = __brev((uint32_t)acc) >> 24;
                    acc >>= 8; fill -= 8;
                    if (nb < nbytes) {
                        uint32_t t8;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t8) : "r"(crc_tab + (crcreg ^ byte)) : "memory");
                        crcreg = t8;
                        if (wr) out[nb] = (uint8_t)byte;
                    } else {
                        crcbits = byte;
                    }
                    ++nb;
                }
            }
        }
        if (wr) {
            P.path_crc[orow] = (crcreg == crcbits) ? 1 : 0;
            P.path_metric[orow] = m;
        }
        // unused list slots of this codeword: metric +inf, crc 0 (as the fast kernel leaves them)
        if (lane >= np && lane < Lsz) {
            P.path_crc[(size_t)w * Lsz + lane] = 0;
            P.path_metric[(size_t)w * Lsz + lane] = CUDART_INF;
        }
        if (lane == 0) P.npaths[w] = np;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// hard-decision fast path (rtwm/fastpolar.py:261-276): one warp per codeword
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_transform(uint32_t x, int lane)
{
    x = xform_word(x);
#pragma unroll
    for (int h = 1; h < 32; h <<= 1) {
        const uint32_t v = __shfl_down_sync(0xffffffffu, x, h);
        if (!(lane & h)) x ^= v;
    }
    return x;
}

// One warp per LLR row.  In neg_mode the row yields two codewords (+row: bit = llr > 0, -row: bit = llr < 0) from
// one set of loads; all 32 coalesced row loads are issued before the first ballot.  Lane j ends up with word j of u-hat
// (positions 32j .. 32j+31); its un-frozen bits are compressed to the low end with the five move masks of that word
// (d_hard_tab, filled by es_polar_set_code) and OR-ed into the K-bit stream at the word's stream offset; lanes then emit
// four stream bytes each.  CRC-8 is linear: every lane runs its own bytes through the table, advances the result over
// the payload bytes that follow its group (d_hard_shift: the state after r more zero bytes) and the warp XORs the parts.
struct HardTab { uint32_t cmp[5]; uint32_t unfrozen; uint32_t cnt; uint32_t off; };
__device__ HardTab d_hard_tab[32];
__device__ uint8_t d_hard_shift[32][256];       // [lane][crc state]: the state advanced over the payload bytes after the lane's group
__device__ uint8_t d_crc8_g[256];               // CRC-8 byte table in global memory (per-lane index)

__global__ void __launch_bounds__(128) scl_hard_kernel(const float* __restrict__ llr, int nrows, int neg_mode,
                                                       uint8_t* __restrict__ hard_payload,
                                                       uint8_t* __restrict__ hard_crc)
{
    __shared__ uint32_t sbuf[4][2][34];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int row = blockIdx.x * (blockDim.x >> 5) + wl;
    if (row >= nrows) return;
    const float* src = llr + (size_t)row * 1024;
    const int K = c_K;
    const int nbytes = (K - 8) >> 3;
    float v[32];
#pragma unroll
    for (int it = 0; it < 32; ++it) v[it] = __ldg(src + it * 32 + lane);
    const HardTab T = d_hard_tab[lane];
    uint32_t xp = 0, xn = 0;
#pragma unroll
    for (int it = 0; it < 32; ++it) {
        const uint32_t bp = __ballot_sync(full, v[it] > 0.0f);
        const uint32_t bn = __ballot_sync(full, v[it] < 0.0f);      // -llr > 0
        if (lane == it) { xp = bp; xn = bn; }
    }
    const int nvar = neg_mode ? 2 : 1;
    sbuf[wl][0][lane] = 0; sbuf[wl][1][lane] = 0;
    if (lane < 2) { sbuf[wl][0][32 + lane] = 0; sbuf[wl][1][32 + lane] = 0; }
    __syncwarp();
#pragma unroll 1
    for (int var = 0; var < nvar; ++var) {
        const int w = neg_mode ? (2 * row + var) : row;
        uint32_t u = warp_transform(var ? xn : xp, lane) & T.unfrozen;
#pragma unroll
        for (int s5 = 0; s5 < 5; ++s5) {
            const uint32_t tt = u & T.cmp[s5];
            u = (u ^ tt) | (tt >> (1 << s5));
        }
        uint32_t* sb = sbuf[wl][var];
        const unsigned long long piece = (unsigned long long)u << (T.off & 31);
        if (T.cnt) {
            atomicOr(&sb[T.off >> 5], (uint32_t)piece);
            if ((uint32_t)(piece >> 32)) atomicOr(&sb[(T.off >> 5) + 1], (uint32_t)(piece >> 32));
        }
        __syncwarp();
        // stream bit q sits at bit q & 31 of word q >> 5; byte i = bits 8i .. 8i+7, first bit = MSB
        const uint32_t rev = __brev(sb[lane]);       // byte 4*lane + k = (rev >> (24 - 8k)) & 255
        uint8_t* out = hard_payload + (size_t)w * nbytes;
        uint32_t crc = 0, crcbits = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = 4 * lane + k;
            const uint32_t byte = (rev >> (24 - 8 * k)) & 255u;
            if (i < nbytes) { out[i] = (uint8_t)byte; crc = d_crc8_g[crc ^ byte]; }
            else if (i == nbytes) crcbits = byte;
        }
        crc = d_hard_shift[lane][crc];               // lanes past the payload hold 0 -> 0
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { crc ^= __shfl_xor_sync(full, crc, o); crcbits |= __shfl_xor_sync(full, crcbits, o); }
        if (lane == 0) hard_crc[w] = (crc == crcbits) ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// compaction of the CRC-passing candidates (the only decoder outputs the host validator needs):
// (codeword, slot, payload) with slot 0 = hard decision, 1.. = list rank + 1.  Unordered (atomic
// append); the host sorts by (codeword, slot).  One thread per (codeword, slot).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scl_collect_hits_kernel(const uint8_t* __restrict__ hard_crc,
                                                               const uint8_t* __restrict__ path_crc,
                                                               const uint8_t* __restrict__ hard_payload,
                                                               const uint8_t* __restrict__ path_payload,
                                                               long long ncw, int L, int nbytes, int cap,
                                                               int32_t* __restrict__ counter, int64_t* __restrict__ out_cw,
                                                               int32_t* __restrict__ out_slot, uint8_t* __restrict__ out_payload)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nslot = L + 1;
    if (g >= ncw * nslot) return;
    const long long w = g / nslot;
    const int slot = (int)(g - w * nslot);
    const bool hit = slot == 0 ? (hard_crc[w] != 0) : (path_crc[w * L + (slot - 1)] != 0);
    if (!hit) return;
    const int idx = atomicAdd(counter, 1);
    if (idx >= cap) return;
    out_cw[idx] = w;
    out_slot[idx] = slot;
    const uint8_t* src = slot == 0 ? hard_payload + w * nbytes : path_payload + (w * L + (slot - 1)) * nbytes;
    uint8_t* dst = out_payload + (long long)idx * nbytes;
    for (int b = 0; b < nbytes; ++b) dst[b] = src[b];
}

// ---------------------------------------------------------------------------------------------
// encoder (rtwm/fastpolar.py:237-252): one warp per payload, output one byte per code bit
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) polar_encode_kernel(const uint8_t* __restrict__ payload, int n,
                                                           uint8_t* __restrict__ cw_bits,
                                                           uint32_t* __restrict__ cw_words)
{
    __shared__ uint32_t su[4][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int w = blockIdx.x * (blockDim.x >> 5) + wl;
    if (w >= n) return;
    const int K = c_K;
    const int nbytes = (K - 8) >> 3;
    const uint8_t* src = payload + (size_t)w * nbytes;
    // CRC-8 over the payload bits, MSB first (every lane computes it; 55 bytes)
    uint8_t crc = 0;
    for (int b = 0; b < nbytes; ++b) {
        const uint8_t byte = __ldg(src + b);
#pragma unroll
        for (int t = 7; t >= 0; --t) crc = crc8_step_bit(crc, (byte >> t) & 1u);
    }
    su[wl][lane] = 0;
    __syncwarp();
    for (int q = lane; q < K; q += 32) {
        const int byte_idx = q >> 3;
        const uint8_t byte = (byte_idx < nbytes) ? __ldg(src + byte_idx) : crc;
        const uint32_t b = (byte >> (7 - (q & 7))) & 1u;
        const int pos = d_datapos[q];
        if (b) atomicOr(&su[wl][pos >> 5], 1u << (pos & 31));
    }
    __syncwarp();
    const uint32_t x = warp_transform(su[wl][lane], lane);
    if (cw_words) cw_words[(size_t)w * 32 + lane] = x;
    if (cw_bits) {
        uint8_t* out = cw_bits + (size_t)w * 1024;
        for (int it = 0; it < 32; ++it) {
            const uint32_t wv = __shfl_sync(full, x, it);
            out[it * 32 + lane] = (uint8_t)((wv >> lane) & 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
constexpr int SCL_W = ES_SCL_W;  // warps per CTA = one CTA per SM: one copy of the phi tables per SM (4 warps x 4 CTAs measured 4 % slower)
static size_t scl_smem_bytes() { return (size_t)SclLY::TAB_BYTES + (size_t)SCL_W * SclLY::WARP_BYTES; }
static size_t scl_scratch_doubles_per_warp() { return SclLY::G_DOUBLES; }

// per-device launch state (the phi tables and the function attributes belong to a device's context)
struct SclDev { int ctas_per_sm = 0; double* phi_tab = nullptr; };
static SclDev g_scl_dev[ES_MAX_DEVICES];

static int scl_configure(SclDev** out)
{
    int dev = 0;
    ES_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= ES_MAX_DEVICES) { set_error("device %d out of range", dev); return ES_EINVAL; }
    SclDev& D = g_scl_dev[dev];
    *out = &D;
    if (D.ctas_per_sm) return ES_OK;
    int nb = 0;
    {
        auto kern = scl_list_kernel<SCL_S, SCL_W, false>;
        ES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scl_smem_bytes()));
        ES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ES_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, SCL_W * 32, scl_smem_bytes()));
    }
    {
        auto kern = scl_list_kernel<SCL_S, SCL_W, true>;
        int nb2 = 0;
        ES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scl_smem_bytes()));
        ES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ES_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, kern, SCL_W * 32, scl_smem_bytes()));
        if (nb2 < nb) nb = nb2;
    }
    if (nb < 1) { set_error("scl_list_kernel does not fit on an SM"); return ES_EINVAL; }
    // phi tables (phi_impl.h layout)
    static double tab[PHI_TAB_DOUBLES];
    phi_fill_table(tab);
    double kk[PHI_NK];
    phi_fill_k(kk);
    ES_CUDA_OK(cudaMemcpyToSymbol(c_phi_k, kk, sizeof(kk)));
    ES_CUDA_OK(cudaMalloc(&D.phi_tab, sizeof(tab)));
    ES_CUDA_OK(cudaMemcpy(D.phi_tab, tab, sizeof(tab), cudaMemcpyHostToDevice));
    D.ctas_per_sm = nb;
    return ES_OK;
}

}  // namespace es

using namespace es;

extern "C" {

int es_polar_set_code(const uint8_t* frozen_host, int K)
{
    if (!frozen_host || K < 16 || K > 1024 || (K & 7)) { set_error("es_polar_set_code: bad K=%d", K); return ES_EINVAL; }
    uint32_t words[32] = {0};
    uint16_t pos[1024] = {0};
    int n = 0;
    for (int i = 0; i < 1024; ++i) {
        if (frozen_host[i]) words[i >> 5] |= 1u << (i & 31);
        else { if (n < 1024) pos[n] = (uint16_t)i; ++n; }
    }
    if (n != K) { set_error("es_polar_set_code: %d unfrozen positions != K=%d", n, K); return ES_EINVAL; }
    CodeDev& CD = g_code[current_device()];
    if (CD.ready) {
        if (CD.K == K && memcmp(CD.frozen, words, sizeof(words)) == 0) return ES_OK;      // this device already holds this code
        ES_CUDA_OK(cudaDeviceSynchronize());      // a different code: nothing in flight may still be reading the old tables
    }
    ES_CUDA_OK(cudaMemcpyToSymbol(c_frozen, words, sizeof(words)));
    ES_CUDA_OK(cudaMemcpyToSymbol(c_datapos, pos, sizeof(pos)));
    ES_CUDA_OK(cudaMemcpyToSymbol(d_datapos, pos, sizeof(pos)));
    ES_CUDA_OK(cudaMemcpyToSymbol(c_K, &K, sizeof(int)));
    {
        // maximal aligned all-frozen nodes of >= 4 bits that do not start at bit 0 (bit 0 belongs to the spine)
        uint8_t r0[256] = {0};
        for (int v = 8; v >= 1; --v) {
            const int nq = 1 << (v - 1);
            for (int q0 = nq; q0 + nq <= 256; q0 += nq) {
                bool all = true;
                for (int q = q0; q < q0 + nq && all; ++q)
                    all = (r0[q] == 0) && (((words[(4 * q) >> 5] >> ((4 * q) & 31)) & 15u) == 15u);
                if (!all) continue;
                r0[q0] = (uint8_t)v;
                for (int q = q0 + 1; q < q0 + nq; ++q) r0[q] = 255;
            }
        }
        ES_CUDA_OK(cudaMemcpyToSymbol(c_r0, r0, sizeof(r0)));
    }
    {
        uint8_t tab[256];
        for (int v = 0; v < 256; ++v) {
            uint8_t r = (uint8_t)v;
            for (int t = 0; t < 8; ++t) r = (r & 0x80) ? (uint8_t)((r << 1) ^ 0x07) : (uint8_t)(r << 1);
            tab[v] = r;
        }
        ES_CUDA_OK(cudaMemcpyToSymbol(c_crc8, tab, sizeof(tab)));
    }
    {
        // bit compress by a fixed mask (Hacker's Delight 7-4): move masks per word of 32 positions
        uint32_t cmp[32][5], cnt[32];
        for (int j = 0; j < 32; ++j) {
            uint32_t m = ~words[j];
            cnt[j] = (uint32_t)__builtin_popcount(m);
            uint32_t mk = ~m << 1;
            for (int i = 0; i < 5; ++i) {
                uint32_t mp = mk ^ (mk << 1);
                mp ^= mp << 2; mp ^= mp << 4; mp ^= mp << 8; mp ^= mp << 16;
                const uint32_t mv = mp & m;
                cmp[j][i] = mv;
                m = (m ^ mv) | (mv >> (1 << i));
                mk &= ~mp;
            }
        }
        ES_CUDA_OK(cudaMemcpyToSymbol(c_cmp, cmp, sizeof(cmp)));
        ES_CUDA_OK(cudaMemcpyToSymbol(c_cnt, cnt, sizeof(cnt)));
        // tables of the hard-decision kernel: per word the masks, the stream offset; per lane the CRC-8 state advanced over
        // the payload bytes that follow the lane's four stream bytes
        HardTab ht[32];
        uint32_t off = 0;
        for (int j = 0; j < 32; ++j) {
            for (int i = 0; i < 5; ++i) ht[j].cmp[i] = cmp[j][i];
            ht[j].unfrozen = ~words[j]; ht[j].cnt = cnt[j]; ht[j].off = off;
            off += cnt[j];
        }
        ES_CUDA_OK(cudaMemcpyToSymbol(d_hard_tab, ht, sizeof(ht)));
        uint8_t tab[256];
        for (int v = 0; v < 256; ++v) {
            uint8_t r = (uint8_t)v;
            for (int t = 0; t < 8; ++t) r = (r & 0x80) ? (uint8_t)((r << 1) ^ 0x07) : (uint8_t)(r << 1);
            tab[v] = r;
        }
        ES_CUDA_OK(cudaMemcpyToSymbol(d_crc8_g, tab, sizeof(tab)));
        static uint8_t shift[32][256];
        const int nbytes = (K - 8) / 8;
        for (int lane = 0; lane < 32; ++lane) {
            int after = nbytes - (4 * lane + 4);          // payload bytes after this lane's group
            if (after < 0) after = 0;
            for (int v = 0; v < 256; ++v) {
                uint8_t r = (uint8_t)v;
                for (int z = 0; z < after; ++z) r = tab[r];
                shift[lane][v] = r;
            }
        }
        ES_CUDA_OK(cudaMemcpyToSymbol(d_hard_shift, shift, sizeof(shift)));
    }
    { const int rc = tx_set_code(pos, K); if (rc != ES_OK) return rc; }
    CD.ready = 1;
    CD.K = K;
    memcpy(CD.frozen, words, sizeof(words));
    return ES_OK;
}


int es_scl_grid_ctas(void)
{
    SclDev* D;
    if (scl_configure(&D) != ES_OK) return -1;
    return sm_count() * D->ctas_per_sm;
}

int es_scl_ctas_per_sm(void)
{
    SclDev* D;
    if (scl_configure(&D) != ES_OK) return -1;
    return D->ctas_per_sm;
}

size_t es_scl_scratch_bytes(void)
{
    const int ctas = es_scl_grid_ctas();
    if (ctas < 0) return 0;
    return (size_t)ctas * SCL_W * scl_scratch_doubles_per_warp() * sizeof(double);
}

int es_scl_hard(const float* llr, int ncw, int neg_mode, uint8_t* hard_payload, uint8_t* hard_crc, void* stream)
{
    if (!g_code_ready) { set_error("es_scl_hard: call es_polar_set_code first"); return ES_ENOTREADY; }
    if (ncw <= 0) return ES_OK;
    if (neg_mode && (ncw & 1)) { set_error("es_scl_hard: neg_mode takes 2 codewords per row, got ncw=%d", ncw); return ES_EINVAL; }
    const int nrows = neg_mode ? ncw / 2 : ncw;
    const int wpb = 4;
    scl_hard_kernel<<<(nrows + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(llr, nrows, neg_mode, hard_payload, hard_crc);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_scl_list_margin(const float* llr, const int32_t* index, int ncw, int neg_mode, int list_size,
                       void* scratch, size_t scratch_bytes,
                       uint8_t* path_payload, uint8_t* path_crc, double* path_metric, int32_t* npaths, double* min_margin,
                       void* stream)
{
    if (!g_code_ready) { set_error("es_scl_list: call es_polar_set_code first"); return ES_ENOTREADY; }
    if (list_size < 1 || list_size > 8) { set_error("es_scl_list: list_size %d not in 1..8", list_size); return ES_EINVAL; }
    if (ncw <= 0) return ES_OK;
    SclDev* D;
    int rc = scl_configure(&D);
    if (rc != ES_OK) return rc;
    const int pair = (neg_mode && !index && (ncw % 2) == 0) ? 1 : 0;
    const int nunits = pair ? ncw / 2 : ncw;
    const int ngroups = (nunits + 3) / 4;
    int ctas = sm_count() * D->ctas_per_sm;
    const int need_ctas = (ngroups + SCL_W - 1) / SCL_W;
    if (ctas > need_ctas) ctas = need_ctas;
    const size_t need = (size_t)ctas * SCL_W * scl_scratch_doubles_per_warp() * sizeof(double);
    if (!scratch || scratch_bytes < need) { set_error("es_scl_list: scratch %zu < %zu bytes", scratch_bytes, need); return ES_EINVAL; }
    SclParams P;
    P.llr = llr; P.index = index; P.ncw = ncw; P.neg_mode = neg_mode; P.list_size = list_size;
    P.pair = pair; P.nunits = nunits;
    P.scratch = (double*)scratch; P.scratch_stride = scl_scratch_doubles_per_warp();
    P.phi_tab = D->phi_tab;
    P.path_payload = path_payload; P.path_crc = path_crc; P.path_metric = path_metric; P.npaths = npaths;
    P.min_margin = min_margin;
    if (min_margin) scl_list_kernel<SCL_S, SCL_W, true><<<ctas, SCL_W * 32, scl_smem_bytes(), (cudaStream_t)stream>>>(P);
    else scl_list_kernel<SCL_S, SCL_W, false><<<ctas, SCL_W * 32, scl_smem_bytes(), (cudaStream_t)stream>>>(P);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_scl_list(const float* llr, const int32_t* index, int ncw, int neg_mode, int list_size,
                void* scratch, size_t scratch_bytes,
                uint8_t* path_payload, uint8_t* path_crc, double* path_metric, int32_t* npaths, void* stream)
{
    return es_scl_list_margin(llr, index, ncw, neg_mode, list_size, scratch, scratch_bytes, path_payload, path_crc,
                              path_metric, npaths, nullptr, stream);
}

size_t es_scl_wide_scratch_bytes(void)
{
    return (size_t)sm_count() * WIDE_W * WIDE_SCRATCH_BYTES;
}

int es_scl_list_wide(const float* llr, int ncw, int neg_mode, int list_size, void* scratch, size_t scratch_bytes,
                     uint8_t* path_payload, uint8_t* path_crc, double* path_metric, int32_t* npaths, void* stream)
{
    if (!g_code_ready) { set_error("es_scl_list_wide: call es_polar_set_code first"); return ES_ENOTREADY; }
    if (list_size < 1 || list_size > 32) { set_error("es_scl_list_wide: list_size %d not in 1..32", list_size); return ES_EINVAL; }
    if (ncw <= 0) return ES_OK;
    if (neg_mode && (ncw & 1)) { set_error("es_scl_list_wide: neg_mode takes 2 codewords per row, got ncw=%d", ncw); return ES_EINVAL; }
    SclDev* D;
    int rc = scl_configure(&D);
    if (rc != ES_OK) return rc;
    int ctas = sm_count();
    const int need_ctas = (ncw + WIDE_W - 1) / WIDE_W;
    if (ctas > need_ctas) ctas = need_ctas;
    const size_t need = (size_t)ctas * WIDE_W * WIDE_SCRATCH_BYTES;
    if (!scratch || scratch_bytes < need) { set_error("es_scl_list_wide: scratch %zu < %zu bytes", scratch_bytes, need); return ES_EINVAL; }
    const size_t smem = (size_t)SclLY::TAB_BYTES + (size_t)WIDE_W * 32 * 32 * sizeof(uint32_t);
    ES_CUDA_OK(cudaFuncSetAttribute(scl_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // every slot pointer starts at 0 and level 0 is the LLR row itself; the bl bytes of a level are written before they are read
    WideParams P;
    P.llr = llr; P.ncw = ncw; P.neg_mode = neg_mode; P.list_size = list_size;
    P.scratch = (unsigned char*)scratch; P.phi_tab = D->phi_tab;
    P.path_payload = path_payload; P.path_crc = path_crc; P.path_metric = path_metric; P.npaths = npaths;
    scl_wide_kernel<<<ctas, WIDE_W * 32, smem, (cudaStream_t)stream>>>(P);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_scl_collect_hits(const uint8_t* hard_crc, const uint8_t* path_crc, const uint8_t* hard_payload,
                        const uint8_t* path_payload, long long ncw, int list_size, int cap,
                        int32_t* counter, int64_t* out_cw, int32_t* out_slot, uint8_t* out_payload, void* stream)
{
    if (!g_code_ready) { set_error("es_scl_collect_hits: call es_polar_set_code first"); return ES_ENOTREADY; }
    if (ncw <= 0) return ES_OK;
    const int nbytes = (g_K - 8) >> 3;
    const long long total = ncw * (list_size + 1);
    scl_collect_hits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        hard_crc, path_crc, hard_payload, path_payload, ncw, list_size, nbytes, cap, counter, out_cw, out_slot, out_payload);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_polar_encode(const uint8_t* payload, int n, uint8_t* cw_bits, uint32_t* cw_words, void* stream)
{
    if (!g_code_ready) { set_error("es_polar_encode: call es_polar_set_code first"); return ES_ENOTREADY; }
    if (n <= 0) return ES_OK;
    const int wpb = 4;
    polar_encode_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(payload, n, cw_bits, cw_words);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

}  // extern "C"
