// echoseal_b200/csrc/rx_scan.cu — RX scan stages of the detector, hand-written for sm_100a:
//   K1  es_rx_bandpass : 4 x order-8 Butterworth band-pass, fp64, chunked with zero-state warm-up; tiles move through
//                        TMA tensor maps (cp.async.bulk.tensor loads + stores) where the shapes allow
//                        (rtwm/detector.py:59-60  scipy.signal.lfilter(b, a, x.astype(float32)))
//   K2  es_rx_ncc      : cosine-normalised 63-tap preamble correlation, fp64 (rtwm/detector.py:76-79);
//                        es_rx_ncc_hist also forms K3's first pass; es_rx_scan = K1+K2 fused (measured alternative)
//   K3  es_rx_peaks    : exact median / MAD threshold, +-607 non-max suppression, first 25 peaks,
//                        top-5 fallback (rtwm/detector.py:83-99, 107-110)
//   K4  es_rx_frames   : per peak: header decode (rtwm/detector.py:452-515) and the PN-independent
//                        part of _llr: matched filter + shift search (rtwm/detector.py:322-383)
//   K5  es_rx_llr      : per (peak, counter, PN variant): despread + robust LLR scaling
//                        (rtwm/detector.py:384-414)
// Everything else of the detector (hop schedule, PN generation, AEAD validation, candidate budget)
// stays on the host by design and feeds / consumes these kernels (echoseal_b200/detector.py).
//
// Layouts (all contiguous):  x f32 [clips][n] ; y f64 [clips][4][n] ; corr f64 [clips][4][n-62] ;
// peaks i32 [clips][4][25] ; mf_aligned f32 [clips][4][25][1024].
#include "common.cuh"
#include <math_constants.h>
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

namespace es {

constexpr int PRE_L = 63, HDR_L = 128, NPAY = 1024, FRAME_LEN = 1215;
constexpr int NBANDS = 4, PEAK_LIMIT = 25, NMS_HALF = FRAME_LEN / 2;   // 607
constexpr int MAXH = 192;            // matched-filter taps per band (131/116/123/93 at 48 kHz)
constexpr int BP_CHUNK = 2048;       // outputs per thread in K1
constexpr int BP_WARM = 768;         // zero-state warm-up samples (error < 4e-13 of peak)

__constant__ double c_bp_b[NBANDS][9];
__constant__ double c_bp_a[NBANDS][9];
__constant__ double c_tpl[NBANDS][PRE_L];
__constant__ float c_mf[NBANDS][MAXH];
__constant__ int c_mf_len[NBANDS];
// per-device: what the device's constant tables hold (content hash) and which kernels have their attributes set
struct RxDev { int ready = 0; bool oddz = false; unsigned long long sig = 0; int cfg[8] = {0}; };
static RxDev g_rxdev[ES_MAX_DEVICES];
#define g_rx_ready (g_rxdev[current_device()].ready)
static unsigned long long fnv1a(unsigned long long h, const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// ---------------------------------------------------------------------------------------------
// K1: band-pass.  One LANE per (clip, chunk), all four bands: direct-form II transposed, the operation order
// of scipy's lfilter, the 4 x 8 states in registers (four independent recurrences per lane hide the DFMA
// latency); chunks other than the first start from zero state BP_WARM samples early (the filter's impulse
// response has decayed below 1e-12 by then; samples before the clip start are zeros, which keep the zero state
// exactly).  The chunk length depends on n only and is chosen so that a clip fills whole warps (144 000 samples -> 64
// chunks of 2250).  A warp owns 32 consecutive chunks and moves data through shared-memory tiles so that every
// global access is a full row (128-byte input, 256-byte output); the input tile of the next step is in flight
// (cp.async into the other half of a double buffer) while the current one is filtered, and it is read from HBM
// once for the four bands.
// ODDZ: the odd numerator taps of a Butterworth band-pass are exactly 0.0; fma(0, x, z) == z, so they are skipped.
// ---------------------------------------------------------------------------------------------
constexpr int BP_WARPS = 1;
constexpr int BP_STEP = 16;                 // samples per lane per pipeline step
constexpr int BP_MIN_CTAS = 14;             // 146 registers, 13.6 KB shared memory per warp
struct BpWarpShared {
    float xin[2][32][BP_STEP + 1];
    double yout[NBANDS][32][9];             // half a step (8 samples) of the four bands
};
#define g_bp_oddz (g_rxdev[current_device()].oddz)

template <bool ODDZ>
__global__ void __launch_bounds__(BP_WARPS * 32, BP_MIN_CTAS) bandpass_kernel(const float* __restrict__ x, int nclips, int n,
                                                                 long long x_stride, double* __restrict__ y,
                                                                 int ch, int groups)
{
    __shared__ BpWarpShared SH[BP_WARPS];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    BpWarpShared& S = SH[wl];
    const long long wid = (long long)blockIdx.x * BP_WARPS + wl;
    if (wid >= (long long)nclips * groups) return;
    const int grp = (int)(wid % groups);
    const long long clip = wid / groups;
    const float* xs = x + clip * x_stride;
    double* ys = y + clip * NBANDS * (long long)n;
    double z[NBANDS][8];
#pragma unroll
    for (int bd = 0; bd < NBANDS; ++bd)
#pragma unroll
        for (int i = 0; i < 8; ++i) z[bd][i] = 0.0;
    const int chunk0 = grp * 32;
    // step st covers samples [first - BP_WARM + BP_STEP st, +BP_STEP) of every lane's chunk
    static_assert(BP_WARM % BP_STEP == 0 && BP_STEP == 16, "fetch() maps a warp onto two 16-sample rows");
    const int nsteps = (BP_WARM + ch + BP_STEP - 1) / BP_STEP;
    auto fetch = [&](int st) {          // row r = chunk chunk0 + r, BP_STEP consecutive samples; 2 rows per instruction
        const int rel = -BP_WARM + st * BP_STEP + (lane & 15);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&S.xin[st & 1][lane >> 4][lane & 15]);
#pragma unroll 8
        for (int i = 0; i < 16; ++i) {
            const long long j = (long long)(chunk0 + 2 * i + (lane >> 4)) * ch + rel;
            const bool in = (j >= 0 && j < n);
            // 4-byte asynchronous copy; src-size 0 zero-fills (samples outside the clip)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + i * 2 * (BP_STEP + 1) * 4),
                         "l"(xs + (in ? j : 0)), "r"(in ? 4 : 0));
        }
        asm volatile("cp.async.commit_group;");
    };
    fetch(0);
#pragma unroll 1
    for (int st = 0; st < nsteps; ++st) {
        const int rel = -BP_WARM + st * BP_STEP;                   // offset from each chunk's first output
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (st + 1 < nsteps) fetch(st + 1);
        const float(*xin)[BP_STEP + 1] = S.xin[st & 1];
#pragma unroll 1
        for (int q8 = 0; q8 < BP_STEP / 8; ++q8) {
            // 8 samples, the four band recurrences interleaved: one shared load + conversion feeds 4 independent chains
#pragma unroll 2
            for (int t = 0; t < 8; ++t) {
                const double xn = (double)xin[lane][q8 * 8 + t];
#pragma unroll
                for (int bd = 0; bd < NBANDS; ++bd) {
                    const double yn = fma(c_bp_b[bd][0], xn, z[bd][0]);
#pragma unroll
                    for (int i = 0; i < 7; ++i) {
                        const double zi = (ODDZ && !(i & 1)) ? z[bd][i + 1] : fma(c_bp_b[bd][i + 1], xn, z[bd][i + 1]);
                        z[bd][i] = fma(-c_bp_a[bd][i + 1], yn, zi);
                    }
                    z[bd][7] = fma(-c_bp_a[bd][8], yn, c_bp_b[bd][8] * xn);
                    S.yout[bd][lane][t] = yn;
                }
            }
            __syncwarp();
            const int o0 = rel + q8 * 8;
            if (o0 + 7 >= 0) {
                const int o = o0 + (lane & 7);                     // four 64-byte row segments per instruction
#pragma unroll
                for (int bd = 0; bd < NBANDS; ++bd) {
                    double* yb = ys + (long long)bd * n;
#pragma unroll 4
                    for (int r = lane >> 3; r < 32; r += 4) {
                        const long long j = (long long)(chunk0 + r) * ch + o;
                        if (o >= 0 && o < ch && j < n) yb[j] = S.yout[bd][r][lane & 7];
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1, TMA form.  Same lane-per-chunk recurrence, chunk grid and input double buffer as bandpass_kernel; the RESULTS leave
// through the tensor-memory accelerator: every 8 samples the warp's tile [32 chunks][8 samples] of a band is one
// cp.async.bulk.tensor.4d store (SASS UTMASTG) from a 64-byte-swizzled shared tile, coordinates (sample, chunk, band, clip)
// in the tensor map of y.  Only the very last chunk of a clip is ragged (it ends at the clip end, n, not at the chunk
// length): tiles of that chunk beyond n go out through a second tensor map whose box has 31 rows, and lane 31 writes the
// one partially valid tile itself.  The transposing store loop of bandpass_kernel (4 bands x 8 row groups x
// ~12 instructions per 8 samples and lane, two thirds of that kernel's non-FP64 instructions) becomes 4 instructions of
// one lane.  Two tile sets alternate: a set is refilled only after cp.async.bulk.wait_group.read has seen its stores read it.
// ---------------------------------------------------------------------------------------------
struct BpTmaShared {
    double yout[2][NBANDS][32][8];          // 64-byte rows, 16-byte units XOR-swizzled with (row >> 1) & 3 (SWIZZLE_64B)
    union {
        float xtile[2][32][BP_STEP];        // XTMA: input tiles as the tensor load delivers them (64-byte rows, same swizzle)
        float xin[2][32][BP_STEP + 1];      // !XTMA: cp.async tiles, padded rows
    };
    unsigned long long bar[2];              // XTMA: one mbarrier per input stage
};

// XTMA: the INPUT tiles arrive through a tensor map of x as well (cp.async.bulk.tensor.3d, box [32 chunks][16 samples],
// completion on an mbarrier): during the warm-up the box is taken from the tail of the previous chunk row (row -1 of a
// clip is out of bounds: the hardware fills zeros, exactly the zero state the first chunk needs).
template <bool ODDZ, bool XTMA>
__global__ void __launch_bounds__(32, 10) bandpass_tma_kernel(const float* __restrict__ x, int nclips, int n, long long x_stride,
                                                              double* __restrict__ y, const __grid_constant__ CUtensorMap tmy,
                                                              const __grid_constant__ CUtensorMap tmy31,
                                                              const __grid_constant__ CUtensorMap tmx,
                                                              const __grid_constant__ CUtensorMap tmx31, int ch, int groups)
{
    extern __shared__ __align__(1024) unsigned char bp_raw[];
    BpTmaShared& S = *reinterpret_cast<BpTmaShared*>(bp_raw);
    const int lane = threadIdx.x;
    const long long wid = blockIdx.x;
    if (wid >= (long long)nclips * groups) return;
    const int grp = (int)(wid % groups);
    const int clip = (int)(wid / groups);
    const float* xs = x + (long long)clip * x_stride;
    double z[NBANDS][8];
#pragma unroll
    for (int bd = 0; bd < NBANDS; ++bd)
#pragma unroll
        for (int i = 0; i < 8; ++i) z[bd][i] = 0.0;
    const int chunk0 = grp * 32;
    // samples of the clip inside the last chunk of this warp (rows 0..30 are always whole)
    const long long lim_ll = (long long)n - (long long)(chunk0 + 31) * ch;
    const int last_lim = lim_ll >= ch ? ch : (lim_ll < 0 ? 0 : (int)lim_ll);
    const int nsteps = (BP_WARM + ch + BP_STEP - 1) / BP_STEP;
    if (XTMA) {
        if (lane == 0) {
            const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(&S.bar[0]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8u) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
    auto fetch = [&](int st) {
        if (XTMA) {
            // stage st = samples [rel0, rel0 + 16) of every chunk; before the chunk start they are the tail of the previous chunk row
            const int rel0 = -BP_WARM + st * BP_STEP;
            const int col = rel0 < 0 ? ch + rel0 : rel0;
            const int row = rel0 < 0 ? chunk0 - 1 : chunk0;
            const bool whole = rel0 < 0 || rel0 + BP_STEP <= last_lim;        // row 31 of the box lies inside the clip
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&S.xtile[st & 1][0][0]);
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&S.bar[st & 1]);
            if (lane == 0) {
                const CUtensorMap* tm = whole ? &tmx : &tmx31;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(whole ? 2048u : 1984u) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(dst), "l"(tm), "r"(col), "r"(row), "r"(clip), "r"(bar) : "memory");
            }
            if (!whole && lane == 31) {         // the ragged last chunk: its valid samples by hand, zeros beyond the clip
                const long long j0 = (long long)(chunk0 + 31) * ch + rel0;
                for (int k = 0; k < BP_STEP; ++k) {
                    const long long j = j0 + k;
                    S.xtile[st & 1][31][((((k >> 2) ^ ((31 >> 1) & 3)) << 2) | (k & 3))] = (j < n) ? __ldg(xs + j) : 0.0f;
                }
            }
            return;
        }
        // as bandpass_kernel: row r = chunk chunk0 + r, two rows per instruction
        const int rel = -BP_WARM + st * BP_STEP + (lane & 15);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&S.xin[st & 1][lane >> 4][lane & 15]);
#pragma unroll 8
        for (int i = 0; i < 16; ++i) {
            const long long j = (long long)(chunk0 + 2 * i + (lane >> 4)) * ch + rel;
            const bool in = (j >= 0 && j < n);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + i * 2 * (BP_STEP + 1) * 4),
                         "l"(xs + (in ? j : 0)), "r"(in ? 4 : 0));
        }
        asm volatile("cp.async.commit_group;");
    };
    fetch(0);
    // this lane's row in a tile: unit u (two doubles) of row `lane` sits at unit u ^ ((lane >> 1) & 3)
    const int sw = (lane >> 1) & 3;
    int tile = 0;                       // tiles stored so far: set = tile & 1
#pragma unroll 1
    for (int st = 0; st < nsteps; ++st) {
        const int rel = -BP_WARM + st * BP_STEP;
        if (XTMA) {
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&S.bar[st & 1]);
            const uint32_t parity = (uint32_t)(st >> 1) & 1u;
            uint32_t ok;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            } while (!ok);
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        if (st + 1 < nsteps) fetch(st + 1);
        const float(*xin)[BP_STEP + 1] = S.xin[st & 1];
#pragma unroll 1
        for (int q8 = 0; q8 < BP_STEP / 8; ++q8) {
            const int o0 = rel + q8 * 8;
            const bool emit = (o0 >= 0 && o0 < ch);
            const int set = tile & 1;
            float xv[8];
            if (XTMA) {
                // floats 8 q8 .. 8 q8 + 7 of this lane's row: units 2 q8 and 2 q8 + 1, each at unit ^ ((lane >> 1) & 3)
                const float4 a = *reinterpret_cast<const float4*>(&S.xtile[st & 1][lane][((2 * q8) ^ sw) << 2]);
                const float4 b = *reinterpret_cast<const float4*>(&S.xtile[st & 1][lane][((2 * q8 + 1) ^ sw) << 2]);
                xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) xv[t] = xin[lane][q8 * 8 + t];
            }
            if (emit) {
                // the stores issued from this set two tiles ago must have read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const double xn = (double)xv[t];
#pragma unroll
                for (int bd = 0; bd < NBANDS; ++bd) {
                    const double yn = fma(c_bp_b[bd][0], xn, z[bd][0]);
#pragma unroll
                    for (int i = 0; i < 7; ++i) {
                        const double zi = (ODDZ && !(i & 1)) ? z[bd][i + 1] : fma(c_bp_b[bd][i + 1], xn, z[bd][i + 1]);
                        z[bd][i] = fma(-c_bp_a[bd][i + 1], yn, zi);
                    }
                    z[bd][7] = fma(-c_bp_a[bd][8], yn, c_bp_b[bd][8] * xn);
                    if (emit) S.yout[set][bd][lane][(((t >> 1) ^ sw) << 1) | (t & 1)] = yn;
                }
            }
            if (emit) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the tile was written with ordinary stores
                __syncwarp();
                const bool whole = (o0 + 8 <= last_lim);          // row 31 of the tile lies inside the clip
                if (lane == 0) {
                    const CUtensorMap* tm = whole ? &tmy : &tmy31;
#pragma unroll
                    for (int bd = 0; bd < NBANDS; ++bd) {
                        const uint32_t src = (uint32_t)__cvta_generic_to_shared(&S.yout[set][bd][0][0]);
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                                     ::"l"(tm), "r"(o0), "r"(chunk0), "r"(bd), "r"(clip), "r"(src) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (!whole && lane == 31 && o0 < last_lim) {      // the one partially valid tile of the clip's last chunk
                    double* yr = y + (long long)clip * NBANDS * n + (long long)(chunk0 + 31) * ch + o0;
#pragma unroll
                    for (int bd = 0; bd < NBANDS; ++bd)
                        for (int t = 0; t < 8 && o0 + t < last_lim; ++t)
                            yr[(long long)bd * n + t] = S.yout[set][bd][31][(((t >> 1) ^ sw) << 1) | (t & 1)];
                }
                ++tile;
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all stores complete before the CTA exits
}

// ---------------------------------------------------------------------------------------------
// K2: normalised cross-correlation (rtwm/detector.py:75-79).  2048 outputs per CTA from a shared y tile;
// each thread owns 8 CONSECUTIVE outputs so that a loaded sample feeds 8 dot products from registers
// (8.75 shared loads per output instead of 63: the kernel is FP64-issue bound, not shared-memory bound).
// The 8 window energies share the 56 squares common to all of them and add their own 7 head/tail squares:
// only additions of non-negative terms, no sliding subtraction.
// ---------------------------------------------------------------------------------------------
constexpr int NCC_T = 8;
constexpr int NCC_TILE = 256 * NCC_T;
constexpr int NCC_IN = NCC_TILE + PRE_L - 1;
// shared layout: sample j at j + (j >> 3), so that lanes reading j = 8*lane + m are 9 doubles apart (conflict-free)
__host__ __device__ __forceinline__ constexpr int ncc_pad(int j) { return j + (j >> 3); }

constexpr int NCC_SPAN = 12;                    // consecutive tiles per CTA (input tiles double-buffered with cp.async)

// K3's first pass done by its producer (es_rx_ncc_hist): the 2048-bin monotone histogram of the row and the values of the
// six central bins (where the median of a correlation row sits), so that K3 reads corr only once.  Bins as K3's sel_bin<0>.
constexpr int NCC_HBINS = 2048;
constexpr int NCC_SPEC_LO = 1021, NCC_SPEC_HI = 1027, NCC_SPEC_CAP = 6144;
__device__ __forceinline__ int ncc_bin(double v)
{
    const double s = (v + 1.0) * 1024.0;
    return (s >= 2047.0) ? 2047 : ((s <= 0.0) ? 0 : (int)s);
}
struct NccAux { unsigned int* hist; double* spec; unsigned int* nspec; };   // [rows][2048], [rows][NCC_SPEC_CAP], [rows]

template <bool AUX>
__global__ void __launch_bounds__(256, 4) ncc_kernel(const double* __restrict__ y, int n, int nc,
                                                     double* __restrict__ corr, NccAux aux)
{
    extern __shared__ __align__(16) double ncc_sm[];
    constexpr int BUF = ncc_pad(NCC_IN) + 1;
    unsigned int* sh = reinterpret_cast<unsigned int*>(ncc_sm + 2 * BUF);     // AUX: this CTA's share of the row histogram
    if (AUX) {
        for (int b = threadIdx.x; b < NCC_HBINS; b += 256) sh[b] = 0;
        __syncthreads();
    }
    const int cb = blockIdx.y;                  // clip*4 + band
    const int band = cb & 3;
    const double* ys = y + (long long)cb * n;
    double* cs = corr + (long long)cb * nc;
    const int ntiles = (nc + NCC_TILE - 1) / NCC_TILE;
    const int tile0 = blockIdx.x * NCC_SPAN;
    const int tile1 = min(ntiles, tile0 + NCC_SPAN);
    auto fetch = [&](int tile, double* sy) {    // asynchronous copy of the tile's 2110 samples (zeros beyond the row)
        const int i0 = tile * NCC_TILE;
        for (int t = threadIdx.x; t < NCC_IN; t += 256) {
            const int j = i0 + t;
            const bool in = j < n;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sy + ncc_pad(t))),
                         "l"(ys + (in ? j : 0)), "r"(in ? 8 : 0));
        }
        asm volatile("cp.async.commit_group;");
    };
    if (tile0 < tile1) fetch(tile0, ncc_sm);
    for (int tile = tile0; tile < tile1; ++tile) {
        double* sy = ncc_sm + ((tile - tile0) & 1) * BUF;
        if (tile + 1 < tile1) {
            fetch(tile + 1, ncc_sm + ((tile + 1 - tile0) & 1) * BUF);      // in flight during this tile's 600 DFMAs
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int i0 = tile * NCC_TILE;
        const double* sv = sy + 9 * threadIdx.x;    // ncc_pad(8*t + m) = 9*t + m + (m >> 3)
#define NCC_V(m) sv[(m) + ((m) >> 3)]
        double d[NCC_T], win[NCC_T];
#pragma unroll
        for (int q = 0; q < NCC_T; ++q) d[q] = 0.0;
#pragma unroll
        for (int q = 0; q < NCC_T - 1; ++q) win[q] = NCC_V(q);
        double core = 0.0;                          // sum of v[m]^2, m = 7..62: common to the 8 windows
#pragma unroll
        for (int k = 0; k < PRE_L; ++k) {
            win[(k + NCC_T - 1) & (NCC_T - 1)] = NCC_V(k + NCC_T - 1);
            const double w = c_tpl[band][k];
#pragma unroll
            for (int q = 0; q < NCC_T; ++q) d[q] = fma(win[(k + q) & (NCC_T - 1)], w, d[q]);
            if (k >= NCC_T - 1) { const double v = win[k & (NCC_T - 1)]; core = fma(v, v, core); }
        }
        // window q = samples q..q+62 = head (q..6) + core (7..62) + tail (63..62+q)
        double head[NCC_T], tail[NCC_T];
        head[NCC_T - 1] = 0.0;
#pragma unroll
        for (int q = NCC_T - 2; q >= 0; --q) { const double v = NCC_V(q); head[q] = fma(v, v, head[q + 1]); }
        tail[0] = 0.0;
#pragma unroll
        for (int q = 1; q < NCC_T; ++q) { const double v = NCC_V(PRE_L - 1 + q); tail[q] = fma(v, v, tail[q - 1]); }
        double r[NCC_T];
#pragma unroll
        for (int q = 0; q < NCC_T; ++q) r[q] = d[q] / (sqrt((head[q] + core) + tail[q]) + 1e-12);
#undef NCC_V
        // stage the 8 results per thread through this tile's buffer for coalesced stores
        __syncthreads();
        double* so = sy + 9 * threadIdx.x;
#pragma unroll
        for (int q = 0; q < NCC_T; ++q) so[q] = r[q];
        __syncthreads();
        for (int t = threadIdx.x; t < NCC_TILE; t += 256) {
            const int i = i0 + t;
            if (i < nc) {
                const double v = sy[ncc_pad(t)];
                cs[i] = v;
                if (AUX) {
                    const int b = ncc_bin(v);
                    atomicAdd(&sh[b], 1u);
                    if (b >= NCC_SPEC_LO && b < NCC_SPEC_HI) {
                        const unsigned int p = atomicAdd(&aux.nspec[cb], 1u);
                        if (p < (unsigned)NCC_SPEC_CAP) aux.spec[(long long)cb * NCC_SPEC_CAP + p] = v;
                    }
                }
            }
        }
        __syncthreads();                            // the buffer is refilled two tiles later
    }
    if (AUX) {
        unsigned int* gh = aux.hist + (long long)cb * NCC_HBINS;
        for (int b = threadIdx.x; b < NCC_HBINS; b += 256) { const unsigned int c = sh[b]; if (c) atomicAdd(&gh[b], c); }
    }
}

// ---------------------------------------------------------------------------------------------
// K1+K2 fused (es_rx_scan): band-pass and normalised correlation in one pass; the filtered signal never leaves the SM.
// One WARP per (clip, band, 32 consecutive chunks), one lane per chunk, as in K1: the lane runs the order-8 recurrence over
// its chunk (768-sample zero-state warm-up, same chunk grid as K1, so the samples are bit-identical to K1's) and keeps
// the last 64..96 filtered samples of its chunk in a shared-memory history [slot][lane]; after every 8 new samples it
// forms the 8 correlations that have become complete (their 63-sample windows end inside the new samples) exactly like
// K2: 8 consecutive outputs from a 70-sample register window, shared energy core + head/tail squares.  A chunk's last
// 62 windows reach into the next chunk: the lane simply filters 64 samples further (its own recurrence: those samples
// differ from the neighbour's by the 4e-13 warm-up truncation at most).  Input rows arrive by 16-byte cp.async
// (zero-filled outside the clip), double-buffered; results leave through a transposing tile, 128-byte row segments.
// Per sample and band: 4 B read (x, once per band, L2-served for three of them) + 8 B written, against K1+K2's 8 + 16.
// ---------------------------------------------------------------------------------------------
constexpr int SC_L = 96;                     // history slots per lane
constexpr int SC_KEEP = 64;                  // slots carried over when the history wraps (>= 62)
constexpr int SC_XP = 20;                    // floats per input row in shared memory (16 + pad: conflict-free 16-byte reads)
constexpr int SC_EXTRA = 64;                 // samples filtered beyond the chunk for its last windows
struct ScanShared {
    double ring[SC_L][32];
    double out[32][17];
    float xin[2][32][SC_XP];
};

template <bool ODDZ, bool ALIGNED>
__global__ void __launch_bounds__(32) scan_kernel(const float* __restrict__ x, int nclips, int n, long long x_stride,
                                                  double* __restrict__ corr, long long corr_pitch, int ch, int groups)
{
    extern __shared__ __align__(16) unsigned char scan_raw[];
    ScanShared& S = *reinterpret_cast<ScanShared*>(scan_raw);
    const int lane = threadIdx.x;
    const long long wid = blockIdx.x;                    // ((clip * groups) + grp) * 4 + band: the four bands of a chunk group are neighbours (x from L2)
    const int band = (int)(wid & 3);
    const int grp = (int)((wid >> 2) % groups);
    const long long clip = (wid >> 2) / groups;
    const int nc = n - (PRE_L - 1);
    const float* xs = x + clip * x_stride;
    double* cs = corr + (clip * NBANDS + band) * corr_pitch;
    const int chunk0 = grp * 32;
    const long long c0 = (long long)(chunk0 + lane) * ch;          // first sample of this lane's chunk
    double bb[5], aa[8];
#pragma unroll
    for (int i = 0; i < 5; ++i) bb[i] = c_bp_b[band][2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) aa[i] = -c_bp_a[band][i + 1];
    double bo[4] = {0.0, 0.0, 0.0, 0.0};
    if (!ODDZ) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bo[i] = c_bp_b[band][2 * i + 1];
    }
    double z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = 0.0;

    const int nstages = (BP_WARM + ch + SC_EXTRA) / 16;            // ch % 16 == 0
    auto fetch = [&](int sg) {
        const long long g0 = c0 - BP_WARM + 16LL * sg;              // multiple of 16
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&S.xin[sg & 1][lane][0]);
        if (ALIGNED) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long g = g0 + 4 * i;
                long long left = (long long)n - g;                  // samples of the clip at and after g
                const int bytes = (g < 0 || left <= 0) ? 0 : (left >= 4 ? 16 : (int)left * 4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16 * i), "l"(xs + (bytes ? g : 0)), "r"(bytes));
            }
        } else {
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const long long g = g0 + i;
                const bool in = (g >= 0 && g < n);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4 * i), "l"(xs + (in ? g : 0)), "r"(in ? 4 : 0));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    fetch(0);
    int w = 0;                    // history write position
    int t = -BP_WARM;             // time of the next sample, relative to the chunk start
    int R = 0;                    // next correlation index, relative to the chunk start
#pragma unroll 1
    for (int sg = 0; sg < nstages; ++sg) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (sg + 1 < nstages) fetch(sg + 1);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const float4 xa = *reinterpret_cast<const float4*>(&S.xin[sg & 1][lane][8 * half]);
            const float4 xb = *reinterpret_cast<const float4*>(&S.xin[sg & 1][lane][8 * half + 4]);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            double yv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                // direct form II transposed, the operation order of K1 / scipy's lfilter
                const double xn = (double)xv[q];
                const double yn = fma(bb[0], xn, z[0]);
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    const double zi = (i & 1) ? fma(bb[(i + 1) >> 1], xn, z[i + 1]) : (ODDZ ? z[i + 1] : fma(bo[i >> 1], xn, z[i + 1]));
                    z[i] = fma(aa[i], yn, zi);
                }
                z[7] = fma(aa[7], yn, bb[4] * xn);
                yv[q] = yn;
            }
            if (t >= 0) {
                if (w == SC_L) {          // history full: carry the last SC_KEEP samples to the front
#pragma unroll 8
                    for (int k = 0; k < SC_KEEP; ++k) S.ring[k][lane] = S.ring[SC_L - SC_KEEP + k][lane];
                    w = SC_KEEP;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) S.ring[w + q][lane] = yv[q];
                w += 8;
            }
            t += 8;
            if (t >= R + 72 && R < ch) {
                // correlations R .. R+7 of this chunk: windows y[R+q .. R+q+62], history slots w-72 .. w-3
                const double* hv = &S.ring[w - 72][lane];
#define SC_V(m) hv[(m) * 32]
                double d[8], win[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) d[q] = 0.0;
#pragma unroll
                for (int q = 0; q < 7; ++q) win[q] = SC_V(q);
                double core = 0.0;                          // sum of v[m]^2, m = 7..62: common to the 8 windows
#pragma unroll
                for (int k = 0; k < PRE_L; ++k) {
                    win[(k + 7) & 7] = SC_V(k + 7);
                    const double wk = c_tpl[band][k];
#pragma unroll
                    for (int q = 0; q < 8; ++q) d[q] = fma(win[(k + q) & 7], wk, d[q]);
                    if (k >= 7) { const double v = win[k & 7]; core = fma(v, v, core); }
                }
                double head[8], tail[8];
                head[7] = 0.0;
#pragma unroll
                for (int q = 6; q >= 0; --q) { const double v = SC_V(q); head[q] = fma(v, v, head[q + 1]); }
                tail[0] = 0.0;
#pragma unroll
                for (int q = 1; q < 8; ++q) { const double v = SC_V(PRE_L - 1 + q); tail[q] = fma(v, v, tail[q - 1]); }
#undef SC_V
                const int o8 = R & 8;                       // two tiles per output row of 16
#pragma unroll
                for (int q = 0; q < 8; ++q) S.out[lane][o8 + q] = d[q] / (sqrt((head[q] + core) + tail[q]) + 1e-12);
                if (o8) {
                    __syncwarp();
                    const int col = lane & 15;
                    const long long r0 = (long long)(R - 8 + col);      // index inside the chunk
#pragma unroll 4
                    for (int r = lane >> 4; r < 32; r += 2) {
                        const long long i = (long long)(chunk0 + r) * ch + r0;
                        if (i < nc) cs[i] = S.out[r][col];
                    }
                    __syncwarp();
                }
                R += 8;
            }
        }
    }
}

// Band-passed samples of the candidate frames only (K4's input once the scan no longer writes y): thread per
// (clip, band, peak slot), the recurrence from 768 samples before the frame (zero state; exact from the clip start).
__global__ void __launch_bounds__(32) frame_bandpass_kernel(const float* __restrict__ x, int n, long long x_stride,
                                                            const int32_t* __restrict__ peaks, const int32_t* __restrict__ npeaks,
                                                            double* __restrict__ yfr /*[clips][4][25][FRAME_LEN]*/)
{
    const int cb = blockIdx.x, slot = threadIdx.x;
    if (slot >= PEAK_LIMIT || slot >= npeaks[cb]) return;
    const int band = cb & 3;
    const long long clip = cb >> 2;
    const int start = peaks[(long long)cb * PEAK_LIMIT + slot];
    if (start < 0 || start + FRAME_LEN > n) return;
    const float* xs = x + clip * x_stride;
    double* out = yfr + ((long long)cb * PEAK_LIMIT + slot) * FRAME_LEN;
    double z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = 0.0;
    const int t0 = max(0, start - BP_WARM);
#pragma unroll 1
    for (int j = t0; j < start + FRAME_LEN; ++j) {
        const double xn = (double)__ldg(xs + j);
        const double yn = fma(c_bp_b[band][0], xn, z[0]);
#pragma unroll
        for (int i = 0; i < 7; ++i) z[i] = fma(-c_bp_a[band][i + 1], yn, fma(c_bp_b[band][i + 1], xn, z[i + 1]));
        z[7] = fma(-c_bp_a[band][8], yn, c_bp_b[band][8] * xn);
        if (j >= start) out[j - start] = yn;
    }
}

// ---------------------------------------------------------------------------------------------
// K3: exact order statistics + NMS.  One CTA of 1024 threads per (clip, band).
// ---------------------------------------------------------------------------------------------
constexpr int PK_THREADS = 512;        // two CTAs (rows) per SM: one row's serial decisions overlap the other's streaming passes
constexpr int SEL_BINS = 2048;
constexpr int SEL_CAP = 4096;       // values gathered from the selected bin (shared memory)
constexpr int NMS_BLOCK = 4096;

__device__ __forceinline__ uint64_t f64_key(double v)   // order-preserving map to uint64
{
    const uint64_t u = (uint64_t)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(uint64_t k)
{
    const uint64_t u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

struct SelShared {
    unsigned int hist[SEL_BINS];
    double buf[SEL_CAP];
    unsigned int count;
    int bin, bin2;
    unsigned int below, below2, topcount;
    unsigned long long prefix;
    double result, result2;
};

// value(v) = MODE ? |v - center| : v
template <int MODE> __device__ __forceinline__ double sel_value(double v, double center)
{
    return MODE ? fabs(v - center) : v;
}
template <int MODE> __device__ __forceinline__ int sel_bin(double val)
{
    // MODE 0: corr in [-1,1] -> (val+1)*1024 ; MODE 1: |corr-med| in [0,2] -> val*1024.  Monotone in val.
    const double s = MODE ? val * 1024.0 : (val + 1.0) * 1024.0;
    int b = (s >= 2047.0) ? 2047 : ((s <= 0.0) ? 0 : (int)s);
    return b;
}

// exact MSB radix select (8 bits per pass) restricted to one linear bin; used when a bin overflows the
// shared buffer (degenerate distributions, e.g. silence: every value equal)
template <int MODE>
__device__ double radix_select_in_bin(const double* __restrict__ c, int nc, int bin, int kk, double center, SelShared& S)
{
    const int tid = threadIdx.x;
    unsigned long long prefix = 0, mask = 0;
    int kr = kk;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = tid; b < 256; b += PK_THREADS) S.hist[SEL_BINS - 256 + b] = 0;
        __syncthreads();
        for (int i = tid; i < nc; i += PK_THREADS) {
            const double v = sel_value<MODE>(c[i], center);
            if (sel_bin<MODE>(v) != bin) continue;
            const uint64_t key = f64_key(v);
            if ((key & mask) == prefix) atomicAdd(&S.hist[SEL_BINS - 256 + (int)((key >> shift) & 255ull)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int acc = 0;
            int d = 0;
            for (; d < 256; ++d) {
                const unsigned int h = S.hist[SEL_BINS - 256 + d];
                if (acc + h > (unsigned)kr) break;
                acc += h;
            }
            S.below = acc; S.prefix = prefix | ((unsigned long long)d << shift);
        }
        __syncthreads();
        kr -= (int)S.below;
        prefix = S.prefix;
        mask |= 255ull << shift;
        __syncthreads();
    }
    return key_f64(prefix);
}

// np.median of value(c[i]): the middle element, or the mean of the two middle ones (ranks k1 <= k2 = k1+1).
// Two passes over the data: a 2048-bin monotone histogram, then a gather of the bin(s) holding the two
// ranks and an exact rank count in shared memory.  `topbin` (MODE 0 only): smallest bin b with
// count(bins >= b) >= min(5, nc) — used by the top-k fallback.  All threads call it; result broadcast.
template <int MODE>
__device__ double median_of(const double* __restrict__ c, int nc, double center, SelShared& S, int* topbin)
{
    const int tid = threadIdx.x;
    const int k2 = nc >> 1, k1 = (nc - 1) >> 1;
    for (int b = tid; b < SEL_BINS; b += PK_THREADS) S.hist[b] = 0;
    if (tid == 0) S.count = 0;
    __syncthreads();
    for (int i = tid; i < nc; i += PK_THREADS) atomicAdd(&S.hist[sel_bin<MODE>(sel_value<MODE>(c[i], center))], 1u);
    __syncthreads();
    if (tid == 0) {
        unsigned int acc = 0;
        int b = 0;
        for (; b < SEL_BINS; ++b) {
            if (acc + S.hist[b] > (unsigned)k1) break;
            acc += S.hist[b];
        }
        S.bin = b; S.below = acc;
        int b2 = b;
        unsigned int acc2 = acc;
        for (; b2 < SEL_BINS; ++b2) {
            if (acc2 + S.hist[b2] > (unsigned)k2) break;
            acc2 += S.hist[b2];
        }
        S.bin2 = b2; S.below2 = acc2;
        if (topbin) {
            const unsigned int want = nc < 5 ? (unsigned)nc : 5u;
            unsigned int t = 0;
            int tb = SEL_BINS - 1;
            for (; tb > 0; --tb) { t += S.hist[tb]; if (t >= want) break; }
            if (tb == 0) t += S.hist[0];
            *topbin = tb;
            S.topcount = t;
        }
    }
    __syncthreads();
    const int b1 = S.bin, b2 = S.bin2;
    const unsigned int m = S.hist[b1] + ((b2 != b1) ? S.hist[b2] : 0u);
    const int kk1 = k1 - (int)S.below;
    if (m <= (unsigned)SEL_CAP) {
        for (int i = tid; i < nc; i += PK_THREADS) {
            const double v = sel_value<MODE>(c[i], center);
            const int bb = sel_bin<MODE>(v);
            if (bb == b1 || bb == b2) S.buf[atomicAdd(&S.count, 1u)] = v;
        }
        __syncthreads();
        // the kk-th smallest of the gathered values is the v with  #less <= kk < #less + #equal
        const int kk2 = kk1 + (k2 - k1);
        for (int i = tid; i < (int)m; i += PK_THREADS) {
            const double v = S.buf[i];
            int less = 0, eq = 0;
            for (int j = 0; j < (int)m; ++j) {
                const double w = S.buf[j];
                less += (w < v); eq += (w == v);
            }
            if (less <= kk1 && kk1 < less + eq) S.result = v;      // all writers write the same value
            if (less <= kk2 && kk2 < less + eq) S.result2 = v;
        }
        __syncthreads();
        return (S.result + S.result2) * 0.5;
    }
    const int kk2r = k2 - (int)S.below2;
    __syncthreads();
    const double v1 = radix_select_in_bin<MODE>(c, nc, b1, kk1, center, S);
    __syncthreads();
    const double v2 = (k2 == k1) ? v1 : radix_select_in_bin<MODE>(c, nc, b2, kk2r, center, S);
    __syncthreads();
    return (v1 + v2) * 0.5;
}

struct PeakShared {
    SelShared sel;
    int topbin;
    int cand[NMS_BLOCK];
    unsigned int ncand;
    int found[NMS_BLOCK];
    unsigned int nfound;
    int npeaks;
    double top_v[32];
    int top_i[32];
};

// One (clip, band) row, general form: every order statistic by histogram + gather (+ radix select), NMS by
// 4096-index blocks.  ~6 passes over the row.  peaks2_kernel below answers the common case in 2 passes and
// calls this for everything else, so it is also the definition the fast path is tested against.
__device__ void peaks_row_general(const double* __restrict__ c, int nc, int cb, int32_t* __restrict__ pk,
                                  int32_t* __restrict__ npeaks, double* __restrict__ stats, PeakShared& S)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < PEAK_LIMIT; t += PK_THREADS) pk[t] = -1;
    if (nc <= 0) {
        if (tid == 0) { npeaks[cb] = 0; stats[cb * 4 + 0] = 0; stats[cb * 4 + 1] = 0; stats[cb * 4 + 2] = 0; stats[cb * 4 + 3] = 0; }
        return;
    }
    // ---- median, MAD, threshold (rtwm/detector.py:83-86)
    const double med = median_of<0>(c, nc, 0.0, S.sel, &S.topbin);
    __syncthreads();
    const int topbin = S.topbin;
    const unsigned int topcount = S.sel.topcount;
    __syncthreads();
    const double mad = median_of<1>(c, nc, med, S.sel, nullptr) + 1e-12;
    __syncthreads();
    double thr = med + (4.5 * 1.4826) * mad;
    thr = thr < 0.95 ? thr : 0.95;
    // ---- NMS in ascending index blocks until 25 peaks are found (rtwm/detector.py:87-97, 108-110)
    if (tid == 0) S.npeaks = 0;
    __syncthreads();
    for (int blk0 = 0; blk0 < nc; blk0 += NMS_BLOCK) {
        if (tid == 0) { S.ncand = 0; S.nfound = 0; }
        __syncthreads();
        for (int i = blk0 + tid; i < min(nc, blk0 + NMS_BLOCK); i += PK_THREADS)
            if (c[i] >= thr) S.cand[atomicAdd(&S.ncand, 1u)] = i;
        __syncthreads();
        const int ncand = (int)S.ncand;
        for (int q = warp; q < ncand; q += PK_THREADS / 32) {
            const int i = S.cand[q];
            const double v = c[i];
            const int lo = max(0, i - NMS_HALF), hi = min(nc, i + NMS_HALF + 1);
            bool bigger = false;
            for (int j = lo + lane; j < hi; j += 32) bigger |= (c[j] > v);
            if (!__any_sync(0xffffffffu, bigger) && lane == 0) S.found[atomicAdd(&S.nfound, 1u)] = i;
        }
        __syncthreads();
        const int nf = (int)S.nfound;
        // append in ascending index order (rank by counting)
        for (int q = tid; q < nf; q += PK_THREADS) {
            const int i = S.found[q];
            int r = 0;
            for (int j = 0; j < nf; ++j) r += (S.found[j] < i);
            const int slot = S.npeaks + r;
            if (slot < PEAK_LIMIT) pk[slot] = i;
        }
        __syncthreads();
        if (tid == 0) S.npeaks += nf;
        __syncthreads();
        if (S.npeaks >= PEAK_LIMIT) break;
    }
    int np = S.npeaks;
    int fallback = 0;
    if (np == 0) {
        // ---- top-k fallback, k = min(5, nc): descending value (ties: larger index first)
        fallback = 1;
        const int kf = nc < 5 ? nc : 5;
        if (topcount <= (unsigned)NMS_BLOCK) {
            // one pass: everything in the top histogram bins (it holds >= kf values), ranked in shared memory
            if (tid == 0) S.ncand = 0;
            __syncthreads();
            for (int i = tid; i < nc; i += PK_THREADS)
                if (sel_bin<0>(c[i]) >= topbin) S.cand[atomicAdd(&S.ncand, 1u)] = i;
            __syncthreads();
            const int m = (int)S.ncand;
            for (int q = tid; q < m; q += PK_THREADS) {
                const int i = S.cand[q];
                const double v = c[i];
                int r = 0;
                for (int j = 0; j < m; ++j) {
                    const int ij = S.cand[j];
                    const double w = c[ij];
                    r += (w > v) || (w == v && ij > i);
                }
                if (r < kf) pk[r] = i;
            }
            __syncthreads();
        } else {
        for (int r = 0; r < kf; ++r) {
            double bv = -CUDART_INF; int bi = -1;
            for (int i = tid; i < nc; i += PK_THREADS) {
                bool taken = false;
                for (int q = 0; q < r; ++q) taken |= (pk[q] == i);
                const double v = c[i];
                if (!taken && (v > bv || (v == bv && i > bi))) { bv = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { S.top_v[warp] = bv; S.top_i[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int q = 1; q < PK_THREADS / 32; ++q)
                    if (S.top_v[q] > bv || (S.top_v[q] == bv && S.top_i[q] > bi)) { bv = S.top_v[q]; bi = S.top_i[q]; }
                pk[r] = bi;
            }
            __syncthreads();
        }
        }
        np = kf;
    }
    if (tid == 0) {
        npeaks[cb] = np < PEAK_LIMIT ? np : PEAK_LIMIT;
        stats[cb * 4 + 0] = med; stats[cb * 4 + 1] = mad; stats[cb * 4 + 2] = thr; stats[cb * 4 + 3] = (double)fallback;
    }
}

__global__ void __launch_bounds__(PK_THREADS) peaks_kernel(const double* __restrict__ corr, int nc,
                                                           int32_t* __restrict__ peaks, int32_t* __restrict__ npeaks,
                                                           double* __restrict__ stats)
{
    extern __shared__ __align__(16) unsigned char pk_raw[];
    PeakShared& S = *reinterpret_cast<PeakShared*>(pk_raw);
    const int cb = blockIdx.x;
    peaks_row_general(corr + (long long)cb * nc, nc, cb, peaks + (long long)cb * PEAK_LIMIT, npeaks, stats, S);
}

// ---------------------------------------------------------------------------------------------
// K3, two-pass form.  The general form above re-reads the 1.15 MB row about six times from DRAM (the rows of
// all resident CTAs do not fit L2).  Here:
//  pass A  histogram of corr + speculative gather of the six central bins (|corr| < 3/1024): the median of a
//          correlation row sits there, so the gather of the general form is already done;
//  then    from the SAME histogram and the exact median, a bracket r_lo < MAD <= r_hi in whole bin widths with a
//          one-bin safety margin on each side;
//  pass B  exact count of |corr - med| <= r_lo, gather of the values inside the bracket, indices of everything
//          above the conservative threshold med + 6.67 r_lo, indices of the top histogram bins (top-5 fallback);
//  then    exact MAD by rank inside the bracket, exact threshold, NMS on the few candidates in index order.
// Same results as the general form (tests/test_gpu_rx.py compares them on adversarial rows); whenever a buffer
// would overflow or the speculation misses, the row is redone by peaks_row_general.
// ---------------------------------------------------------------------------------------------
constexpr int PK2_CAP = 6144;          // gathered values per stage (about 2700 / 4200 expected for sigma = 1/sqrt(63))
constexpr int PK2_SUB = 2048;          // values of the (sub-)bins holding the two middle ranks
constexpr int PK2_NCAND = 2048;        // threshold candidates / top-bin candidates
constexpr int PK2_SPEC_LO = 1021, PK2_SPEC_HI = 1027;   // speculative median bins
static_assert(PK2_CAP == NCC_SPEC_CAP && PK2_SPEC_LO == NCC_SPEC_LO && PK2_SPEC_HI == NCC_SPEC_HI && SEL_BINS == NCC_HBINS, "K2 forms K3's first pass");
constexpr int PK2_UNROLL = 8;          // loads in flight per thread in the streaming passes
constexpr int PK2_MIN_NC = 8192;       // shorter rows go to the general form directly

struct Pk2Shared {
    unsigned int hist[SEL_BINS];
    unsigned int pref[SEL_BINS + 1];   // pref[b] = number of values in bins < b
    double buf[PK2_CAP];
    double sub[PK2_SUB];
    int cand[PK2_NCAND];
    int top[PK2_NCAND];
    int sorted[PK2_NCAND];
    unsigned int nbuf, nsub, ncand, ntop, nkeep, below_cnt, topcount;
    int topbin, b1, b2, ok, npeaks, pass[32];
    unsigned int below1;
    double r1, r2, r_lo, r_hi, thr_lo;
};
union PkUnion { PeakShared g; Pk2Shared f; };

// ranks ra <= rb (0-based) among S.sub[0..ms): the v with #less <= r < #less + #equal.  Results in S.r1 / S.r2.
__device__ __forceinline__ void pk2_select2(Pk2Shared& S, int ms, int ra, int rb)
{
    for (int i = threadIdx.x; i < ms; i += PK_THREADS) {
        const double v = S.sub[i];
        int less = 0, eq = 0;
        for (int j = 0; j < ms; ++j) {
            const double w = S.sub[j];
            less += (w < v); eq += (w == v);
        }
        if (less <= ra && ra < less + eq) S.r1 = v;        // all writers write the same value
        if (less <= rb && rb < less + eq) S.r2 = v;
    }
    __syncthreads();
}

// aux (optional): the histogram and central-bin values the producer of corr already formed (es_rx_ncc_hist): pass A is
// then a copy of 8 + <= 48 KB instead of a streaming pass over the 1.15 MB row
__global__ void __launch_bounds__(PK_THREADS) peaks2_kernel(const double* __restrict__ corr, int nc,
                                                            int32_t* __restrict__ peaks, int32_t* __restrict__ npeaks,
                                                            double* __restrict__ stats, NccAux aux)
{
    extern __shared__ __align__(16) unsigned char pk_raw[];
    PkUnion& U = *reinterpret_cast<PkUnion*>(pk_raw);
    Pk2Shared& S = U.f;
    const int cb = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const double* c = corr + (long long)cb * nc;
    int32_t* pk = peaks + (long long)cb * PEAK_LIMIT;
    if (nc < PK2_MIN_NC) { peaks_row_general(c, nc, cb, pk, npeaks, stats, U.g); return; }
    const int k2 = nc >> 1, k1 = (nc - 1) >> 1;

    // ---- pass A: histogram + speculative gather
    for (int b = tid; b < SEL_BINS; b += PK_THREADS) S.hist[b] = 0;
    if (tid == 0) { S.nbuf = 0; S.nsub = 0; S.ncand = 0; S.ntop = 0; S.nkeep = 0; S.below_cnt = 0; S.npeaks = 0; }
    __syncthreads();
    if (aux.hist) {
        const unsigned int* gh = aux.hist + (long long)cb * NCC_HBINS;
        for (int b = tid; b < SEL_BINS; b += PK_THREADS) S.hist[b] = gh[b];
        const unsigned int ns = aux.nspec[cb];
        const double* gs = aux.spec + (long long)cb * NCC_SPEC_CAP;
        const unsigned int take = ns < (unsigned)PK2_CAP ? ns : (unsigned)PK2_CAP;
        for (unsigned int i = tid; i < take; i += PK_THREADS) S.buf[i] = gs[i];
        if (tid == 0) S.nbuf = ns;
    } else {
    // PK2_UNROLL loads in flight per thread: with 1024 threads per SM the streaming passes need that to reach HBM speed
    for (int i0 = tid; i0 < nc; i0 += PK_THREADS * PK2_UNROLL) {
        double vv[PK2_UNROLL];
#pragma unroll
        for (int u = 0; u < PK2_UNROLL; ++u) { const int i = i0 + u * PK_THREADS; vv[u] = (i < nc) ? __ldg(c + i) : 0.0; }
#pragma unroll
        for (int u = 0; u < PK2_UNROLL; ++u) {
            const int i = i0 + u * PK_THREADS;
            if (i >= nc) break;
            const double v = vv[u];
            const int b = sel_bin<0>(v);
            atomicAdd(&S.hist[b], 1u);
            if (b >= PK2_SPEC_LO && b < PK2_SPEC_HI) {
                const unsigned int p = atomicAdd(&S.nbuf, 1u);
                if (p < (unsigned)PK2_CAP) S.buf[p] = v;
            }
        }
    }
    }
    __syncthreads();
    if (warp == 0) {            // exclusive prefix over the 2048 bins: 64 bins per lane
        unsigned int loc = 0;
        for (int b = lane * 64; b < lane * 64 + 64; ++b) loc += S.hist[b];
        unsigned int inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        unsigned int run = inc - loc;
        for (int b = lane * 64; b < lane * 64 + 64; ++b) { S.pref[b] = run; run += S.hist[b]; }
        if (lane == 31) S.pref[SEL_BINS] = run;
    }
    __syncthreads();
    if (tid == 0) {
        // bins of ranks k1, k2 (binary search on the prefix: largest b with pref[b] <= k)
        int lo = 0, hi = SEL_BINS - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (S.pref[mid] <= (unsigned)k1) lo = mid; else hi = mid - 1; }
        S.b1 = lo; S.below1 = S.pref[lo];
        lo = S.b1; hi = SEL_BINS - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (S.pref[mid] <= (unsigned)k2) lo = mid; else hi = mid - 1; }
        S.b2 = lo;
        // smallest bin tb with count(bins >= tb) >= min(5, nc): as in median_of<0>
        const unsigned int want = 5u;
        int tb = SEL_BINS - 1;
        for (; tb > 0; --tb) if ((unsigned)nc - S.pref[tb] >= want) break;
        S.topbin = tb; S.topcount = (unsigned)nc - S.pref[tb];
        const unsigned int inbins = S.hist[S.b1] + ((S.b2 != S.b1) ? S.hist[S.b2] : 0u);
        S.ok = (S.b1 >= PK2_SPEC_LO && S.b2 < PK2_SPEC_HI && S.nbuf <= (unsigned)PK2_CAP && inbins <= (unsigned)PK2_SUB) ? 1 : 0;
    }
    __syncthreads();
    if (!S.ok) { __syncthreads(); peaks_row_general(c, nc, cb, pk, npeaks, stats, U.g); return; }
    {
        const int b1 = S.b1, b2 = S.b2;
        const int m = (int)S.nbuf;
        for (int i = tid; i < m; i += PK_THREADS) {
            const double v = S.buf[i];
            const int bb = sel_bin<0>(v);
            if (bb == b1 || bb == b2) S.sub[atomicAdd(&S.nsub, 1u)] = v;
        }
        __syncthreads();
        pk2_select2(S, (int)S.nsub, k1 - (int)S.below1, k2 - (int)S.below1);
    }
    const double med = (S.r1 + S.r2) * 0.5;
    const int topbin = S.topbin;
    const bool want_top = S.topcount <= (unsigned)PK2_NCAND;

    // ---- MAD bracket from the corr histogram: |v - med| <= m/1024 certainly holds inside bins (bl+2 .. bh-2) and
    //      certainly fails outside bins (bl-1 .. bh+1), bl/bh = bins of med -+ m/1024
    if (tid == 0) {
        const double smed = (med + 1.0) * 1024.0;
        auto P = [&](int b) -> unsigned int { return S.pref[b < 0 ? 0 : (b > SEL_BINS ? SEL_BINS : b)]; };
        auto upper = [&](int m) -> unsigned int {
            const int bl = (int)floor(smed - (double)m), bh = (int)floor(smed + (double)m);
            return P(bh + 2) - P(bl - 1);
        };
        auto lower = [&](int m) -> unsigned int {
            const int bl = (int)floor(smed - (double)m), bh = (int)floor(smed + (double)m);
            return (bh - 1 > bl + 2) ? (P(bh - 1) - P(bl + 2)) : 0u;
        };
        int mlo = -1;
        if (upper(0) <= (unsigned)k1) {
            int lo = 0, hi = 2 * SEL_BINS;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (upper(mid) <= (unsigned)k1) lo = mid; else hi = mid - 1; }
            mlo = lo;
        }
        int lo = 0, hi = 2 * SEL_BINS + 4;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (lower(mid) >= (unsigned)k2 + 1u) hi = mid; else lo = mid + 1; }
        S.r_lo = (mlo < 0) ? -1.0 : (double)mlo / 1024.0;
        S.r_hi = (double)lo / 1024.0;
        const double t0 = med + (4.5 * 1.4826) * (((mlo < 0) ? 0.0 : S.r_lo) + 1e-12);
        S.thr_lo = (t0 < 0.95 ? t0 : 0.95) - 1e-9;
        S.nbuf = 0; S.nsub = 0;
    }
    __syncthreads();
    const double r_lo = S.r_lo, r_hi = S.r_hi, thr_lo = S.thr_lo;

    // ---- pass B
    {
        unsigned int mine = 0;
        for (int i0 = tid; i0 < nc; i0 += PK_THREADS * PK2_UNROLL) {
          double vv[PK2_UNROLL];
#pragma unroll
          for (int u = 0; u < PK2_UNROLL; ++u) { const int i = i0 + u * PK_THREADS; vv[u] = (i < nc) ? __ldg(c + i) : 0.0; }
#pragma unroll
          for (int u = 0; u < PK2_UNROLL; ++u) {
            const int i = i0 + u * PK_THREADS;
            if (i >= nc) break;
            const double v = vv[u];
            const double d = sel_value<1>(v, med);
            if (d <= r_lo) ++mine;
            else if (d <= r_hi) {
                const unsigned int p = atomicAdd(&S.nbuf, 1u);
                if (p < (unsigned)PK2_CAP) S.buf[p] = d;
            }
            if (v >= thr_lo) {
                const unsigned int p = atomicAdd(&S.ncand, 1u);
                if (p < (unsigned)PK2_NCAND) S.cand[p] = i;
            }
            if (want_top && sel_bin<0>(v) >= topbin) {
                const unsigned int p = atomicAdd(&S.ntop, 1u);
                if (p < (unsigned)PK2_NCAND) S.top[p] = i;
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0 && mine) atomicAdd(&S.below_cnt, mine);
    }
    for (int b = tid; b < 256; b += PK_THREADS) S.hist[b] = 0;
    __syncthreads();
    const int mb = (int)S.nbuf;
    const int ra0 = k1 - (int)S.below_cnt, rb0 = k2 - (int)S.below_cnt;
    if (mb > PK2_CAP || S.ncand > (unsigned)PK2_NCAND || ra0 < 0 || rb0 >= mb) {
        __syncthreads(); peaks_row_general(c, nc, cb, pk, npeaks, stats, U.g); return;
    }
    // refine inside the bracket: 256 monotone sub-bins over (base, r_hi]
    const double base = r_lo < 0.0 ? 0.0 : r_lo;
    const double scale = 256.0 / (r_hi - base);
    auto subbin = [&](double d) -> int { const double t = (d - base) * scale; return t >= 255.0 ? 255 : (t <= 0.0 ? 0 : (int)t); };
    for (int i = tid; i < mb; i += PK_THREADS) atomicAdd(&S.hist[subbin(S.buf[i])], 1u);
    __syncthreads();
    if (tid == 0) {
        unsigned int acc = 0;
        int b = 0;
        for (; b < 256; ++b) { if (acc + S.hist[b] > (unsigned)ra0) break; acc += S.hist[b]; }
        S.b1 = b; S.below1 = acc;
        int b2 = b;
        unsigned int acc2 = acc;
        for (; b2 < 256; ++b2) { if (acc2 + S.hist[b2] > (unsigned)rb0) break; acc2 += S.hist[b2]; }
        S.b2 = b2;
        const unsigned int inbins = S.hist[b] + ((b2 != b) ? S.hist[b2] : 0u);
        S.ok = (inbins <= (unsigned)PK2_SUB) ? 1 : 0;
    }
    __syncthreads();
    if (!S.ok) { __syncthreads(); peaks_row_general(c, nc, cb, pk, npeaks, stats, U.g); return; }
    {
        const int b1 = S.b1, b2 = S.b2;
        for (int i = tid; i < mb; i += PK_THREADS) {
            const double d = S.buf[i];
            const int bb = subbin(d);
            if (bb == b1 || bb == b2) S.sub[atomicAdd(&S.nsub, 1u)] = d;
        }
        __syncthreads();
        pk2_select2(S, (int)S.nsub, ra0 - (int)S.below1, rb0 - (int)S.below1);
    }
    const double mad = (S.r1 + S.r2) * 0.5 + 1e-12;
    double thr = med + (4.5 * 1.4826) * mad;
    thr = thr < 0.95 ? thr : 0.95;

    // ---- candidates >= thr, ascending index order, NMS until 25 peaks (rtwm/detector.py:87-97, 108-110)
    for (int t = tid; t < PEAK_LIMIT; t += PK_THREADS) pk[t] = -1;
    const int ncand = (int)S.ncand;
    for (int q = tid; q < ncand; q += PK_THREADS) {
        const int i = S.cand[q];
        if (c[i] >= thr) S.sorted[atomicAdd(&S.nkeep, 1u)] = i;
    }
    __syncthreads();
    const int nk = (int)S.nkeep;
    for (int q = tid; q < nk; q += PK_THREADS) {          // rank by counting -> cand[] ascending
        const int i = S.sorted[q];
        int r = 0;
        for (int j = 0; j < nk; ++j) r += (S.sorted[j] < i);
        S.cand[r] = i;
    }
    __syncthreads();
    for (int q0 = 0; q0 < nk; q0 += PK_THREADS / 32) {
        const int q = q0 + warp;
        bool pass = false;
        if (q < nk) {
            const int i = S.cand[q];
            const double v = c[i];
            const int lo = max(0, i - NMS_HALF), hi = min(nc, i + NMS_HALF + 1);
            bool bigger = false;
            for (int j = lo + lane; j < hi; j += 32) bigger |= (c[j] > v);
            pass = !__any_sync(0xffffffffu, bigger);
        }
        if (lane == 0) S.pass[warp] = pass ? 1 : 0;
        __syncthreads();
        if (tid == 0) {
            int np0 = S.npeaks;
            for (int w = 0; w < PK_THREADS / 32 && q0 + w < nk; ++w)
                if (S.pass[w]) { if (np0 < PEAK_LIMIT) pk[np0] = S.cand[q0 + w]; ++np0; }
            S.npeaks = np0;
        }
        __syncthreads();
        if (S.npeaks >= PEAK_LIMIT) break;
    }
    int np = S.npeaks;
    int fallback = 0;
    if (np == 0) {
        // ---- top-k fallback, k = min(5, nc): descending value (ties: larger index first)
        if (!want_top || S.ntop > (unsigned)PK2_NCAND) { __syncthreads(); peaks_row_general(c, nc, cb, pk, npeaks, stats, U.g); return; }
        fallback = 1;
        const int kf = nc < 5 ? nc : 5;
        const int m = (int)S.ntop;
        for (int q = tid; q < m; q += PK_THREADS) {
            const int i = S.top[q];
            const double v = c[i];
            int r = 0;
            for (int j = 0; j < m; ++j) {
                const int ij = S.top[j];
                const double w = c[ij];
                r += (w > v) || (w == v && ij > i);
            }
            if (r < kf) pk[r] = i;
        }
        np = kf;
    }
    if (tid == 0) {
        npeaks[cb] = np < PEAK_LIMIT ? np : PEAK_LIMIT;
        stats[cb * 4 + 0] = med; stats[cb * 4 + 1] = mad; stats[cb * 4 + 2] = thr; stats[cb * 4 + 3] = (double)fallback;
    }
}

// ---------------------------------------------------------------------------------------------
// K3, long-recording form (one long recording needs GLOBAL order statistics).  Same results
// as peaks_kernel, but every streaming pass over corr is spread over many CTAs and the tiny decisions
// in between run in one CTA per (clip, band).  Two-level monotone histogram (2048 x 2048 bins), gather of
// the sub-bin(s) holding the two middle ranks, exact rank count; NMS per 4096-index block with the first
// 25 peaks collected in index order afterwards.  The host picks this form for correlations >= 2^21 samples.
// ---------------------------------------------------------------------------------------------
constexpr int K3_CAP = 32768;              // gathered values per (clip, band)
constexpr int K3_CHUNK = 1 << 18;          // elements per CTA in the streaming passes

struct K3Meta {
    int b[2]; unsigned int below[2];       // level-1 bins of ranks k1, k2 and counts below them
    int sb[2]; unsigned int sbelow[2];     // level-2 sub-bins and counts below (inside the bin)
    unsigned int cnt;                      // gather counter of the current selection (reset by locate<1>)
    int topbin; unsigned int topcount;     // smallest bin with >= min(5,nc) values at or above it (corr histogram)
    unsigned int topcnt;                   // indices gathered from those bins
    int overflow;
    double med, mad, thr;
};

template <int MODE> __device__ __forceinline__ double k3_scaled(double val)
{
    return MODE ? val * 1024.0 : (val + 1.0) * 1024.0;
}
template <int MODE> __device__ __forceinline__ int k3_sub(double val, int bin)
{
    const double f = (k3_scaled<MODE>(val) - (double)bin) * 2048.0;
    return (f >= 2047.0) ? 2047 : ((f <= 0.0) ? 0 : (int)f);
}

// LEVEL 1: hist[cb][2048] over all values.  LEVEL 2: hist2[cb][2][2048] over the values of bins b[0] / b[1].
template <int MODE, int LEVEL>
// All streaming K3 kernels take the row stride and an index range [lo, hi) of the row: the whole row on one GPU,
// the rank's own part of it (inside a local array that also holds the halos) when one recording is split in time.
__global__ void __launch_bounds__(1024) k3_hist_kernel(const double* __restrict__ corr, long long stride, int lo, int hi,
                                                       const K3Meta* __restrict__ meta, unsigned int* __restrict__ hist)
{
    __shared__ unsigned int sh[2][2048];
    const int cb = blockIdx.y;
    const double* c = corr + (long long)cb * stride;
    const K3Meta M = meta[cb];
    const double center = MODE ? M.med : 0.0;
    for (int b = threadIdx.x; b < 2048; b += 1024) { sh[0][b] = 0; sh[1][b] = 0; }
    __syncthreads();
    const long long i0 = lo + (long long)blockIdx.x * K3_CHUNK;
    const long long i1 = min((long long)hi, i0 + K3_CHUNK);
    for (long long i = i0 + threadIdx.x; i < i1; i += 1024) {
        const double v = sel_value<MODE>(c[i], center);
        const int bin = sel_bin<MODE>(v);
        if (LEVEL == 1) atomicAdd(&sh[0][bin], 1u);
        else {
            if (bin == M.b[0]) atomicAdd(&sh[0][k3_sub<MODE>(v, bin)], 1u);
            else if (bin == M.b[1]) atomicAdd(&sh[1][k3_sub<MODE>(v, bin)], 1u);
        }
    }
    __syncthreads();
    unsigned int* out = hist + (long long)cb * (LEVEL == 1 ? 2048 : 4096);
    for (int b = threadIdx.x; b < 2048; b += 1024) {
        if (sh[0][b]) atomicAdd(&out[b], sh[0][b]);
        if (LEVEL == 2 && sh[1][b]) atomicAdd(&out[2048 + b], sh[1][b]);
    }
}

// one warp per (clip, band): locate the bins (LEVEL 1) or sub-bins (LEVEL 2) of ranks k1 <= k2
template <int LEVEL>
__global__ void k3_locate_kernel(const unsigned int* __restrict__ hist, int nc, K3Meta* __restrict__ meta, int want_top)
{
    const int cb = blockIdx.x;
    if (threadIdx.x != 0) return;
    K3Meta& M = meta[cb];
    const int k2 = nc >> 1, k1 = (nc - 1) >> 1;
    if (LEVEL == 1) {
        const unsigned int* h = hist + (long long)cb * 2048;
        unsigned int acc = 0; int b = 0;
        for (; b < 2048; ++b) { if (acc + h[b] > (unsigned)k1) break; acc += h[b]; }
        M.b[0] = b; M.below[0] = acc;
        for (; b < 2048; ++b) { if (acc + h[b] > (unsigned)k2) break; acc += h[b]; }
        M.b[1] = b; M.below[1] = acc;
        if (want_top) {
            const unsigned int want = nc < 5 ? (unsigned)nc : 5u;
            unsigned int t = 0; int tb = 2047;
            for (; tb > 0; --tb) { t += h[tb]; if (t >= want) break; }
            if (tb == 0) t += h[0];
            M.topbin = tb; M.topcount = t;
        }
        M.cnt = 0; M.overflow = 0;
    } else {
        for (int q = 0; q < 2; ++q) {
            // when both ranks share a bin the second half of hist2 is unused: both look in half 0
            const int half = (q == 1 && M.b[1] != M.b[0]) ? 1 : 0;
            const unsigned int* h = hist + (long long)cb * 4096 + half * 2048;
            const int kk = (q ? k2 : k1) - (int)M.below[half ? 1 : 0];
            unsigned int acc = 0; int b = 0;
            for (; b < 2048; ++b) { if (acc + h[b] > (unsigned)kk) break; acc += h[b]; }
            M.sb[q] = b; M.sbelow[q] = acc;
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(1024) k3_gather_kernel(const double* __restrict__ corr, long long stride, int lo, int hi,
                                                         K3Meta* __restrict__ meta, double* __restrict__ buf)
{
    const int cb = blockIdx.y;
    const double* c = corr + (long long)cb * stride;
    K3Meta& M = meta[cb];
    const int b0 = M.b[0], b1 = M.b[1], s0 = M.sb[0], s1 = M.sb[1];
    const double center = MODE ? M.med : 0.0;
    double* out = buf + (long long)cb * K3_CAP;
    const long long i0 = lo + (long long)blockIdx.x * K3_CHUNK;
    const long long i1 = min((long long)hi, i0 + K3_CHUNK);
    for (long long i = i0 + threadIdx.x; i < i1; i += 1024) {
        const double v = sel_value<MODE>(c[i], center);
        const int bin = sel_bin<MODE>(v);
        if (bin != b0 && bin != b1) continue;
        const int sub = k3_sub<MODE>(v, bin);
        if ((bin == b0 && sub == s0) || (bin == b1 && sub == s1)) {
            const unsigned int q = atomicAdd(&M.cnt, 1u);
            if (q < (unsigned)K3_CAP) out[q] = v; else M.overflow = 1;
        }
    }
}

// exact ranks inside the gathered sub-bin(s) -> median (MODE 0) or MAD + threshold (MODE 1)
template <int MODE>
__global__ void __launch_bounds__(1024) k3_finish_kernel(const double* __restrict__ buf, int nc, K3Meta* __restrict__ meta)
{
    __shared__ double res[2];
    const int cb = blockIdx.x;
    K3Meta& M = meta[cb];
    const double* v = buf + (long long)cb * K3_CAP;
    const int m = (int)min(M.cnt, (unsigned)K3_CAP);
    const int k2 = nc >> 1, k1 = (nc - 1) >> 1;
    // ranks inside the union of the gathered sub-bins (contiguous in rank, see median_of)
    const int base = (int)M.below[0] + (int)M.sbelow[0];
    const int kk1 = k1 - base, kk2 = k2 - base;
    for (int i = threadIdx.x; i < m; i += 1024) {
        const double x = v[i];
        int less = 0, eq = 0;
        for (int j = 0; j < m; ++j) { const double w = v[j]; less += (w < x); eq += (w == x); }
        if (less <= kk1 && kk1 < less + eq) res[0] = x;
        if (less <= kk2 && kk2 < less + eq) res[1] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double md = (res[0] + res[1]) * 0.5;
        if (MODE == 0) M.med = md;
        else {
            M.mad = md + 1e-12;
            double thr = M.med + (4.5 * 1.4826) * M.mad;
            M.thr = thr < 0.95 ? thr : 0.95;
        }
    }
}

// NMS per 4096-index block: up to 25 peaks of the block in index order + their count
// (nc = length of the row as stored: the +-607 windows clamp there; candidates come from [lo, hi) only)
__global__ void __launch_bounds__(1024) k3_nms_kernel(const double* __restrict__ corr, int nc, int lo, int hi,
                                                      const K3Meta* __restrict__ meta,
                                                      int nblk, int32_t* __restrict__ blk_peaks, int32_t* __restrict__ blk_count)
{
    __shared__ int cand[NMS_BLOCK];
    __shared__ int found[NMS_BLOCK];
    __shared__ unsigned int ncand, nfound;
    const int cb = blockIdx.y, blk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* c = corr + (long long)cb * nc;
    const double thr = meta[cb].thr;
    if (tid == 0) { ncand = 0; nfound = 0; }
    __syncthreads();
    const int blk0 = lo + blk * NMS_BLOCK;
    for (int i = blk0 + tid; i < min(hi, blk0 + NMS_BLOCK); i += 1024)
        if (c[i] >= thr) cand[atomicAdd(&ncand, 1u)] = i;
    __syncthreads();
    const int nca = (int)ncand;
    for (int q = warp; q < nca; q += 32) {
        const int i = cand[q];
        const double v = c[i];
        const int lo = max(0, i - NMS_HALF), hi = min(nc, i + NMS_HALF + 1);
        bool bigger = false;
        for (int j = lo + lane; j < hi; j += 32) bigger |= (c[j] > v);
        if (!__any_sync(0xffffffffu, bigger) && lane == 0) found[atomicAdd(&nfound, 1u)] = i;
    }
    __syncthreads();
    const int nf = (int)nfound;
    int32_t* out = blk_peaks + ((long long)cb * nblk + blk) * PEAK_LIMIT;
    for (int q = tid; q < nf; q += 1024) {
        const int i = found[q];
        int r = 0;
        for (int j = 0; j < nf; ++j) r += (found[j] < i);
        if (r < PEAK_LIMIT) out[r] = i;
    }
    if (tid == 0) blk_count[(long long)cb * nblk + blk] = nf;
}

// first 25 peaks in index order; top-5 fallback from the top histogram bins (gathered by k3_top_kernel)
// (nc = number of values of the WHOLE row: k of the top-k fallback)
__global__ void __launch_bounds__(1024) k3_collect_kernel(const double* __restrict__ corr, long long stride, int nc, K3Meta* __restrict__ meta, int nblk,
                                                          const int32_t* __restrict__ blk_peaks, const int32_t* __restrict__ blk_count,
                                                          const int32_t* __restrict__ top_idx,
                                                          int32_t* __restrict__ peaks, int32_t* __restrict__ npeaks, double* __restrict__ stats)
{
    const int cb = blockIdx.x, tid = threadIdx.x;
    const double* c = corr + (long long)cb * stride;
    K3Meta& M = meta[cb];
    int32_t* pk = peaks + (long long)cb * PEAK_LIMIT;
    __shared__ int np_s;
    if (tid == 0) {
        int np = 0;
        for (int t = 0; t < PEAK_LIMIT; ++t) pk[t] = -1;
        for (int blk = 0; blk < nblk && np < PEAK_LIMIT; ++blk) {
            const int nf = blk_count[(long long)cb * nblk + blk];
            const int32_t* src = blk_peaks + ((long long)cb * nblk + blk) * PEAK_LIMIT;
            for (int q = 0; q < nf && q < PEAK_LIMIT && np < PEAK_LIMIT; ++q) pk[np++] = src[q];
            if (nf > PEAK_LIMIT) np = PEAK_LIMIT;   // more than 25 peaks inside one block: the first 25 are all there
        }
        np_s = np;
    }
    __syncthreads();
    int np = np_s, fallback = 0;
    if (np == 0) {
        fallback = 1;
        const int kf = nc < 5 ? nc : 5;
        const int m = (int)min(M.topcnt, (unsigned)K3_CAP);    // filled by k3_top_kernel
        const int32_t* ti = top_idx + (long long)cb * K3_CAP;
        for (int q = tid; q < m; q += 1024) {
            const int i = ti[q];
            const double v = c[i];
            int r = 0;
            for (int j = 0; j < m; ++j) { const int ij = ti[j]; const double w = c[ij]; r += (w > v) || (w == v && ij > i); }
            if (r < kf) pk[r] = i;
        }
        np = kf;
    }
    __syncthreads();
    if (tid == 0) {
        npeaks[cb] = np;
        stats[cb * 4 + 0] = M.med; stats[cb * 4 + 1] = M.mad; stats[cb * 4 + 2] = M.thr; stats[cb * 4 + 3] = (double)fallback;
    }
}

// indices of all values in bins >= topbin (the top-5 live there)
__global__ void __launch_bounds__(1024) k3_top_kernel(const double* __restrict__ corr, long long stride, int lo, int hi,
                                                      K3Meta* __restrict__ meta, int32_t* __restrict__ top_idx)
{
    const int cb = blockIdx.y;
    const double* c = corr + (long long)cb * stride;
    K3Meta& M = meta[cb];
    const int topbin = M.topbin;
    int32_t* out = top_idx + (long long)cb * K3_CAP;
    const long long i0 = lo + (long long)blockIdx.x * K3_CHUNK;
    const long long i1 = min((long long)hi, i0 + K3_CHUNK);
    for (long long i = i0 + threadIdx.x; i < i1; i += 1024) {
        if (sel_bin<0>(c[i]) >= topbin) {
            const unsigned int q = atomicAdd(&M.topcnt, 1u);
            if (q < (unsigned)K3_CAP) out[q] = (int32_t)i; else M.overflow = 1;
        }
    }
}

__global__ void k3_flag_kernel(const K3Meta* meta, int ncb, int32_t* overflow)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncb && meta[i].overflow) atomicOr(overflow, 1);
}

// ---------------------------------------------------------------------------------------------
// K4: per-peak front end: header decode + matched filter / shift search of the payload
// ---------------------------------------------------------------------------------------------
constexpr int FR_THREADS = 256;
constexpr int MF_MAX = NPAY + 2 * MAXH;      // upper bound of conv length we keep

struct FrameShared {
    double hs[MAXH + 8];             // taps (zero-padded by 4 on both sides) and frame widened once: float->double
    double frame[FRAME_LEN];         // conversions run on the quarter-rate unit, two per DFMA would bound the kernel
    float mf[MF_MAX];
    double pre[MF_MAX + 1];
    float score[2 * MAXH + 64];
    float d[HDR_L];
    float sums[16];
    int best;
};

// out[t - tbase] = np.convolve(sig, h, 'full')[t] = sum_k sig[k] * h[t-k] for t in [tbase, tbase + nout), fp64
// accumulation in ascending k (sig, h hold float32 values; h is zero outside [0, nh), 4 zeros readable on
// each side).  Four consecutive outputs per thread: one sample load and one tap load feed four DFMAs; the
// zero taps only add exact zeros, so every output equals its own k-ascending sum.
__device__ __forceinline__ void conv_block(const double* __restrict__ sig, int nsig, const double* __restrict__ h, int nh,
                                           int tbase, int nout, float* __restrict__ out)
{
    for (int o = 4 * threadIdx.x; o < nout; o += 4 * FR_THREADS) {
        const int t0 = tbase + o;
        const int k0 = max(0, t0 - (nh - 1)), k1 = min(nsig - 1, t0 + 3);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        double h0 = h[t0 - k0], h1 = h[t0 + 1 - k0], h2 = h[t0 + 2 - k0], h3 = h[t0 + 3 - k0];
#pragma unroll 4
        for (int k = k0; k <= k1; ++k) {
            const double sk = sig[k];
            a0 = fma(sk, h0, a0); a1 = fma(sk, h1, a1); a2 = fma(sk, h2, a2); a3 = fma(sk, h3, a3);
            h3 = h2; h2 = h1; h1 = h0; h0 = h[t0 - k - 1];
        }
        out[o] = (float)a0;
        if (o + 1 < nout) out[o + 1] = (float)a1;
        if (o + 2 < nout) out[o + 2] = (float)a2;
        if (o + 3 < nout) out[o + 3] = (float)a3;
    }
}

__global__ void __launch_bounds__(FR_THREADS) frames_kernel(const double* __restrict__ y, int n, int framed,
                                                            const int32_t* __restrict__ peaks,
                                                            const int32_t* __restrict__ npeaks,
                                                            const uint8_t* __restrict__ hdr_pn /*[clips][16]*/,
                                                            float* __restrict__ mf_aligned, int32_t* __restrict__ llr_best_s,
                                                            float* __restrict__ hdr_out, int32_t* __restrict__ hdr_best_s)
{
    __shared__ FrameShared S;
    const int slot = blockIdx.x, cb = blockIdx.y;
    const int band = cb & 3, clip = cb >> 2;
    const int tid = threadIdx.x;
    const long long pidx = (long long)cb * PEAK_LIMIT + slot;
    const int start = (slot < npeaks[cb]) ? peaks[pidx] : -1;
    if (start < 0 || start + FRAME_LEN > n) {       // rtwm/detector.py:112-113
        if (tid == 0) { hdr_out[pidx * 4 + 0] = -1.0f; hdr_out[pidx * 4 + 1] = 0; hdr_out[pidx * 4 + 2] = 0; hdr_out[pidx * 4 + 3] = 0;
                        llr_best_s[pidx] = 0; hdr_best_s[pidx] = 0; }
        return;
    }
    // framed: y holds the band-passed samples of the candidate frames only, [clip][band][slot][FRAME_LEN] (frame_bandpass_kernel)
    const double* ys = framed ? y + pidx * FRAME_LEN : y + (long long)cb * n + start;
    for (int t = tid; t < FRAME_LEN; t += FR_THREADS) S.frame[t] = (double)(float)ys[t];   // frame.astype(float32)
    __syncthreads();
    const int nh = c_mf_len[band];
    const int mem = nh - 1;
    for (int t = tid; t < nh + 8; t += FR_THREADS)                       // divergent indices below: keep taps in smem
        S.hs[t] = (t >= 4 && t < nh + 4) ? (double)c_mf[band][t - 4] : 0.0;
    __syncthreads();
    const double* h = S.hs + 4;
    // ======== header (rtwm/detector.py:452-515) ========
    {
        const int prefix = min(mem, PRE_L);
        const double* sig = S.frame + PRE_L - prefix;
        const int nsig = prefix + HDR_L;
        const int nmf = nsig + nh - 1;
        conv_block(sig, nsig, h, nh, 0, nmf, S.mf);
        __syncthreads();
        const int offset = mem + prefix;
        int maxs = min(HDR_L / 2 + prefix, 4 * nh);
        if (maxs < mem) maxs = mem;
        const int wstart = max(0, offset - maxs), wstop = min(nmf, offset + HDR_L + maxs);
        const int base = offset - wstart;
        const int wlen = wstop - wstart;
        const int guard = max(8, min(32, nh / 8));
        const float* win = S.mf + wstart;
        for (int si = tid; si <= 2 * maxs; si += FR_THREADS) {
            const int s = si - maxs;
            const int i0 = base + s;
            float sc = -2.0f;                           // skipped shifts never win (score >= 0 > -1)
            if (i0 >= 0 && i0 + HDR_L <= wlen) {
                double acc = 0.0;
                for (int k = guard; k < HDR_L; ++k) {
                    const float pn = ((hdr_pn[clip * 16 + (k >> 3)] >> (7 - (k & 7))) & 1) ? 1.0f : -1.0f;
                    acc += (double)(win[i0 + k] * pn);
                }
                sc = fabsf((float)acc);
            }
            S.score[si] = sc;
        }
        __syncthreads();
        if (tid == 0) {
            int bs = 0; float bsc = -1.0f;
            for (int si = 0; si <= 2 * maxs; ++si) if (S.score[si] > bsc) { bsc = S.score[si]; bs = si - maxs; }
            S.best = bs;
        }
        __syncthreads();
        const int i0 = base + S.best;
        if (tid < HDR_L) {
            const float pn = ((hdr_pn[clip * 16 + (tid >> 3)] >> (7 - (tid & 7))) & 1) ? 1.0f : -1.0f;
            S.d[tid] = win[i0 + tid] * pn;
        }
        __syncthreads();
        if (tid < 16) {
            float s8 = 0.0f;
            for (int k = 0; k < 8; ++k) s8 += S.d[tid * 8 + k];
            S.sums[tid] = s8;
        }
        __syncthreads();
        if (tid == 0) {
            int val = 0, npos = 0;
            double mabs = 0.0, msq = 0.0, mean = 0.0;
            for (int q = 0; q < 16; ++q) {
                val = (val << 1) | (S.sums[q] < 0.0f ? 1 : 0);     // inverted bit sense, as the reference (quirk 2)
                npos += (S.sums[q] > 0.0f);
                mabs += fabs((double)S.sums[q]);
            }
            mabs /= 16.0;
            for (int k = 0; k < HDR_L; ++k) { msq += (double)S.d[k] * (double)S.d[k]; mean += (double)S.d[k]; }
            msq /= HDR_L; mean /= HDR_L;
            double var = 0.0;
            for (int k = 0; k < HDR_L; ++k) { const double e = (double)S.d[k] - mean; var += e * e; }
            var /= HDR_L;
            const float margin = (float)mabs / ((float)sqrt(msq) + 1e-12f);
            const float score = (float)(mabs / (sqrt(var) + 1e-12));
            const bool ok = (npos >= 10) && (margin > 0.5f);
            hdr_out[pidx * 4 + 0] = ok ? 1.0f : 0.0f;
            hdr_out[pidx * 4 + 1] = (float)val;
            hdr_out[pidx * 4 + 2] = score;
            hdr_out[pidx * 4 + 3] = margin;
            hdr_best_s[pidx] = S.best;
        }
        __syncthreads();
    }
    // ======== payload matched filter + shift search (rtwm/detector.py:322-383) ========
    {
        const int pstart = PRE_L + HDR_L;
        const int prefix = min(mem, pstart);
        const double* sig = S.frame + pstart - prefix;
        const int nsig = prefix + NPAY;
        const int nmf = nsig + nh - 1;
        const int offset = prefix + mem;
        const int raw = min(min(NPAY / 2, 4 * nh), HDR_L);
        const int maxs = max(mem, raw);
        const int wstart = max(0, offset - maxs), wstop = min(nmf, offset + NPAY + maxs);
        const int wlen = wstop - wstart;
        const int base = offset - wstart;
        int guard = min(NPAY / 4, max(nh / 2, 24));
        conv_block(sig, nsig, h, nh, wstart, wlen, S.mf);
        __syncthreads();
        // prefix sums of |mf| (fp64) -> score(s) = mean |mf_win[i0+guard : i0+n]|  (PN cancels under |.|, quirk 3)
        if (tid < 32) {
            // warp 0: blocked scan
            const int per = (wlen + 31) / 32;
            const int a0 = tid * per, a1 = min(wlen, a0 + per);
            double s = 0.0;
            for (int t = a0; t < a1; ++t) s += (double)fabsf(S.mf[t]);
            double incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += v;
            }
            double run = incl - s;
            if (tid == 0) S.pre[0] = 0.0;
            for (int t = a0; t < a1; ++t) { run += (double)fabsf(S.mf[t]); S.pre[t + 1] = run; }
        }
        __syncthreads();
        for (int si = tid; si <= 2 * maxs; si += FR_THREADS) {
            const int i0 = base + si - maxs;
            float sc = -2.0f;
            if (i0 >= 0 && i0 + NPAY <= wlen)
                sc = (float)((S.pre[i0 + NPAY] - S.pre[i0 + guard]) / (double)(NPAY - guard));
            S.score[si] = sc;
        }
        __syncthreads();
        if (tid == 0) {
            int bs = 0; float bsc = -1.0f;
            for (int si = 0; si <= 2 * maxs; ++si) if (S.score[si] > bsc) { bsc = S.score[si]; bs = si - maxs; }
            S.best = bs;
            llr_best_s[pidx] = bs;
        }
        __syncthreads();
        const int i0 = base + S.best;
        float* out = mf_aligned + pidx * NPAY;
        for (int t = tid; t < NPAY; t += FR_THREADS) out[t] = S.mf[i0 + t];
    }
}

// ---------------------------------------------------------------------------------------------
// K5: despread + LLR scaling, one WARP per (item, variant): the 1024 despread chips live in registers
// (32 per lane); the exact medians (np.median of the tail and of |tail - median|) come from a bit-wise
// radix select over order-preserving 32-bit keys with warp-wide REDUX counts — no sort, no barriers.
// ---------------------------------------------------------------------------------------------
constexpr int LLR_WARPS = 4;

__device__ __forceinline__ uint32_t f32_key(float v)
{
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// kth smallest (0-based) of the 32x32 keys held by the warp (invalid entries carry key 0xffffffff)
// MSB-first: after bit b the answer lies in the bucket [prefix, prefix + 2^b), whose population is tracked; as soon
// as it holds one key that key is the answer (the smallest key >= prefix), typically after ~20 of the 32 bits.
__device__ __forceinline__ uint32_t warp_select(const uint32_t (&key)[32], int kth)
{
    uint32_t prefix = 0;
    int below = 0, inb = 1024;              // keys < prefix ; keys inside the current bucket
#pragma unroll 1
    for (int b = 31; b >= 0; --b) {
        const uint32_t cand = prefix | (1u << b);
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) cnt += (key[j] < cand) ? 1 : 0;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt <= kth) { inb -= cnt - below; below = cnt; prefix = cand; }
        else inb = cnt - below;
        if (inb == 1) {
            uint32_t m = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < 32; ++j) m = (key[j] >= prefix) ? min(m, key[j]) : m;
            return __reduce_min_sync(0xffffffffu, m);
        }
    }
    return prefix;
}

// np.median semantics on the valid keys: middle element, or the float32 mean of the two middle ones
__device__ __forceinline__ float warp_median(const uint32_t (&key)[32], int nt)
{
    if (nt & 1) return key_f32(warp_select(key, nt >> 1));
    const int k1 = (nt >> 1) - 1;
    const uint32_t q1 = warp_select(key, k1);
    int le = 0;
    uint32_t nxt = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        le += (key[j] <= q1) ? 1 : 0;
        if (key[j] > q1) nxt = min(nxt, key[j]);
    }
    le = __reduce_add_sync(0xffffffffu, le);
    nxt = __reduce_min_sync(0xffffffffu, nxt);
    const uint32_t q2 = (le > k1 + 1) ? q1 : nxt;
    return (key_f32(q1) + key_f32(q2)) * 0.5f;
}

__global__ void __launch_bounds__(LLR_WARPS * 32) llr_kernel(const float* __restrict__ mf_aligned,
                                                             const int32_t* __restrict__ item_peak,
                                                             const uint8_t* __restrict__ pn_packed /*[items][152]*/,
                                                             int nwork, float* __restrict__ llr)
{
    const int lane = threadIdx.x & 31;
    const int work = blockIdx.x * LLR_WARPS + (threadIdx.x >> 5);
    if (work >= nwork) return;
    const int item = work >> 1, variant = work & 1;
    const int pidx = item_peak[item];
    const int band = (pidx / PEAK_LIMIT) & 3;
    const int nh = c_mf_len[band];
    const int guard = min(NPAY / 4, max(nh / 2, 24));
    const int nt = NPAY - guard;
    const float* mf = mf_aligned + (long long)pidx * NPAY;
    const uint8_t* pn = pn_packed + (long long)item * 152;
    const int pn_off = variant ? 0 : (PRE_L + HDR_L);      // variant 0: bits [191,1215) ; variant 1: bits [0,1024)
    float d[32];
    uint32_t key[32];
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int t = j * 32 + lane;
        const int q = pn_off + t;
        const float s = ((__ldg(pn + (q >> 3)) >> (7 - (q & 7))) & 1) ? 1.0f : -1.0f;
        d[j] = __ldg(mf + t) * s;
        const bool tail = t >= guard;
        key[j] = tail ? f32_key(d[j]) : 0xffffffffu;
        if (tail) acc += (double)d[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float mu = (float)(acc / nt);
    const float medv = warp_median(key, nt);
    double a2 = 0.0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const bool tail = (j * 32 + lane) >= guard;
        key[j] = tail ? f32_key(fabsf(d[j] - medv)) : 0xffffffffu;
        if (tail) { const double e = (double)d[j] - (double)mu; a2 += e * e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    const float madv = warp_median(key, nt);
    const float sd = (float)sqrt(a2 / nt);
    const double sigma_mad = 1.4826 * ((double)madv + 1e-12);
    const double sigma_std = (double)sd + 1e-12;
    const double sigma = fmax(fmax(sigma_mad, sigma_std), 0.1);
    double scale = 2.0 / (sigma * sigma);
    scale = fmin(fmax(scale, 0.5), 30.0);
    const float fscale = (float)scale;
    float* out = llr + (long long)work * NPAY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float v = (d[j] - mu) * fscale;
        out[j * 32 + lane] = fminf(fmaxf(v, -12.0f), 12.0f);
    }
}

// ---------------------------------------------------------------------------------------------
// K9: polyphase resampler = scipy.signal.resample_poly(x, up, down) (rtwm/utils.py:58-66), one thread per
// output sample.  out[m] = sum_j hp[phase + j*up] * x[i_hi - j],  t = (m + n_pre_remove)*down, i_hi = t / up,
// phase = t % up, with hp the zero-pre-padded Kaiser(5) low-pass of 2*10*max(up,down)+1 taps scaled by `up`
// (designed on the host, passed in polyphase order taps[phase][j]).  fp64 accumulation, float32 output.
// ---------------------------------------------------------------------------------------------
template <typename TIN>
__global__ void __launch_bounds__(256) resample_kernel(const TIN* __restrict__ x, long long n_in, long long x_stride,
                                                       int up, int down, const double* __restrict__ taps, int J,
                                                       long long n_pre_remove, long long n_out,
                                                       float* __restrict__ y)
{
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_out) return;
    const TIN* xs = x + (long long)blockIdx.y * x_stride;
    const long long t = (m + n_pre_remove) * down;
    const long long i_hi = t / up;
    const int phase = (int)(t - i_hi * up);
    const double* h = taps + (long long)phase * J;
    double acc = 0.0;
    for (int j = 0; j < J; ++j) {
        const long long i = i_hi - j;
        if (i < 0) break;
        if (i < n_in) acc = fma(__ldg(h + j), (double)xs[i], acc);
    }
    y[(long long)blockIdx.y * n_out + m] = (float)acc;
}

}  // namespace es

using namespace es;

extern "C" {

int es_rx_set_filters(const double* bp_b /*[4][9]*/, const double* bp_a /*[4][9]*/, const double* tpl /*[4][63]*/,
                      const float* mf /*[4][192]*/, const int* mf_len /*[4]*/)
{
    for (int b = 0; b < NBANDS; ++b)
        if (mf_len[b] < 1 || mf_len[b] > MAXH) { set_error("es_rx_set_filters: mf_len[%d]=%d out of range", b, mf_len[b]); return ES_EINVAL; }
    RxDev& RD = g_rxdev[current_device()];
    unsigned long long sig = 1469598103934665603ull;
    sig = fnv1a(sig, bp_b, sizeof(double) * NBANDS * 9); sig = fnv1a(sig, bp_a, sizeof(double) * NBANDS * 9);
    sig = fnv1a(sig, tpl, sizeof(double) * NBANDS * PRE_L); sig = fnv1a(sig, mf, sizeof(float) * NBANDS * MAXH);
    sig = fnv1a(sig, mf_len, sizeof(int) * NBANDS);
    if (RD.ready) {
        if (RD.sig == sig) return ES_OK;                 // this device already holds exactly these tables: nothing to upload
        ES_CUDA_OK(cudaDeviceSynchronize());             // different filters / taps: nothing in flight may still read the old ones
    }
    RD.sig = sig;
    ES_CUDA_OK(cudaMemcpyToSymbol(c_bp_b, bp_b, sizeof(double) * NBANDS * 9));
    ES_CUDA_OK(cudaMemcpyToSymbol(c_bp_a, bp_a, sizeof(double) * NBANDS * 9));
    ES_CUDA_OK(cudaMemcpyToSymbol(c_tpl, tpl, sizeof(double) * NBANDS * PRE_L));
    ES_CUDA_OK(cudaMemcpyToSymbol(c_mf, mf, sizeof(float) * NBANDS * MAXH));
    ES_CUDA_OK(cudaMemcpyToSymbol(c_mf_len, mf_len, sizeof(int) * NBANDS));
    g_bp_oddz = true;
    for (int b = 0; b < NBANDS; ++b)
        for (int i = 1; i < 8; i += 2) g_bp_oddz = g_bp_oddz && (bp_b[b * 9 + i] == 0.0);
    g_rx_ready = 1;
    return ES_OK;
}

// chunk grid of K1 (and of the fused scan): a function of n ONLY -- the chunk boundaries decide where the (4e-13)
// warm-up truncation falls, so a clip's filtered samples, and everything downstream, are bit-identical whatever batch
// it is verified in.  About BP_CHUNK samples per chunk, a clip = a whole number of 32-chunk warps, the chunk length a
// multiple of 16 samples so that the rows of y start on 128-byte lines.
static void bp_chunk_grid(int n, int* groups_out, int* ch_out)
{
    int groups = (int)(((long long)n + 16LL * BP_CHUNK) / (32LL * BP_CHUNK));
    if (groups < 1) groups = 1;
    int ch = (int)(((long long)n + 32LL * groups - 1) / (32LL * groups));
    ch = (ch + 15) & ~15;
    *groups_out = groups; *ch_out = ch;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int g_bp_force_plain = 0, g_bp_force_cpasync = 0;
// tests: 1 = the non-TMA form on the same chunk grid; 2 = TMA stores but cp.async input tiles; 0 = default
void es_rx_bandpass_force_plain(int on) { g_bp_force_plain = (on == 1) ? 1 : 0; g_bp_force_cpasync = (on == 2) ? 1 : 0; }

int es_rx_bandpass(const float* x, int nclips, int n, long long x_stride, double* y, void* stream)
{
    if (!g_rx_ready) { set_error("es_rx_bandpass: call es_rx_set_filters first"); return ES_ENOTREADY; }
    if (nclips <= 0 || n <= 0) return ES_OK;
    int groups, ch;
    bp_chunk_grid(n, &groups, &ch);
    const long long warps = (long long)nclips * groups;
    const unsigned grid = (unsigned)((warps + BP_WARPS - 1) / BP_WARPS);
    // TMA form: rows of y 16-byte aligned (n even), and only the last chunk of a clip ragged
    const bool tma_ok = !g_bp_force_plain && (n % 2) == 0 && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0) &&
                        (32LL * groups * ch - n) < ch && warps <= 0x7fffffffLL;
    EncodeTiledFn enc = tma_ok ? tensor_map_encoder() : nullptr;
    if (enc) {
        // y as a 4-d tensor (sample in chunk, chunk, band, clip); box = 8 samples x 32 (31) chunks of one band of one clip.
        // The chunk dimension counts only chunks that START inside the clip; a box row beyond it is dropped by the hardware.
        CUtensorMap tm, tm31;
        const cuuint64_t nchunks = (cuuint64_t)(((long long)n + ch - 1) / ch);
        const cuuint64_t dims[4] = {(cuuint64_t)ch, nchunks, (cuuint64_t)NBANDS, (cuuint64_t)nclips};
        const cuuint64_t strides[3] = {(cuuint64_t)ch * 8u, (cuuint64_t)n * 8u, (cuuint64_t)n * 8u * NBANDS};
        const cuuint32_t box[4] = {8, 32, 1, 1}, box31[4] = {8, 31, 1, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = enc(&tm31, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)y, dims, strides, box31, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        // x as a 3-d tensor (sample in chunk, chunk, clip) when its rows are 16-byte aligned: the input tiles come by TMA too
        CUtensorMap tmx, tmx31;
        // (the warm-up box is the tail of ONE previous chunk row: chunks at least as long as the warm-up)
        bool xtma = r == CUDA_SUCCESS && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((x_stride & 3) == 0) && ch >= BP_WARM;
        if (xtma) {
            const cuuint64_t xdims[3] = {(cuuint64_t)ch, nchunks, (cuuint64_t)nclips};
            const cuuint64_t xstrides[2] = {(cuuint64_t)ch * 4u, (cuuint64_t)x_stride * 4u};
            const cuuint32_t xbox[3] = {BP_STEP, 32, 1}, xbox31[3] = {BP_STEP, 31, 1};
            const cuuint32_t xe[3] = {1, 1, 1};
            CUresult rx = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, xdims, xstrides, xbox, xe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rx == CUDA_SUCCESS)
                rx = enc(&tmx31, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, xdims, xstrides, xbox31, xe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            xtma = (rx == CUDA_SUCCESS) && !g_bp_force_cpasync;
        }
        if (!xtma) { tmx = tm; tmx31 = tm31; }
        if (r == CUDA_SUCCESS) {
            int& configured = g_rxdev[current_device()].cfg[5];
            if (!configured) {
                const void* fns[4] = {(const void*)bandpass_tma_kernel<false, false>, (const void*)bandpass_tma_kernel<false, true>,
                                      (const void*)bandpass_tma_kernel<true, false>, (const void*)bandpass_tma_kernel<true, true>};
                for (int i = 0; i < 4; ++i) {
                    ES_CUDA_OK(cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BpTmaShared) + 1024));
                    ES_CUDA_OK(cudaFuncSetAttribute(fns[i], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                }
                configured = 1;
            }
            const size_t smem = sizeof(BpTmaShared);
            const unsigned gw = (unsigned)warps;
            cudaStream_t st = (cudaStream_t)stream;
            if (g_bp_oddz) {
                if (xtma) bandpass_tma_kernel<true, true><<<gw, 32, smem, st>>>(x, nclips, n, x_stride, y, tm, tm31, tmx, tmx31, ch, groups);
                else bandpass_tma_kernel<true, false><<<gw, 32, smem, st>>>(x, nclips, n, x_stride, y, tm, tm31, tmx, tmx31, ch, groups);
            } else {
                if (xtma) bandpass_tma_kernel<false, true><<<gw, 32, smem, st>>>(x, nclips, n, x_stride, y, tm, tm31, tmx, tmx31, ch, groups);
                else bandpass_tma_kernel<false, false><<<gw, 32, smem, st>>>(x, nclips, n, x_stride, y, tm, tm31, tmx, tmx31, ch, groups);
            }
            ES_CUDA_OK(cudaGetLastError());
            return ES_OK;
        }
    }
    if (g_bp_oddz) bandpass_kernel<true><<<grid, BP_WARPS * 32, 0, (cudaStream_t)stream>>>(x, nclips, n, x_stride, y, ch, groups);
    else bandpass_kernel<false><<<grid, BP_WARPS * 32, 0, (cudaStream_t)stream>>>(x, nclips, n, x_stride, y, ch, groups);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_rx_scan(const float* x, int nclips, int n, long long x_stride, double* corr, void* stream)
{
    if (!g_rx_ready) { set_error("es_rx_scan: call es_rx_set_filters first"); return ES_ENOTREADY; }
    const int nc = n - (PRE_L - 1);
    if (nclips <= 0 || nc <= 0) return ES_OK;
    int groups, ch;
    bp_chunk_grid(n, &groups, &ch);          // the chunk grid of es_rx_bandpass
    const long long blocks = (long long)nclips * groups * NBANDS;
    if (blocks > 0x7fffffffLL) { set_error("es_rx_scan: %lld warps exceed the grid", blocks); return ES_EINVAL; }
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((x_stride & 3) == 0);
    const int variant = (g_bp_oddz ? 2 : 0) | (aligned ? 1 : 0);
    int& configured = g_rxdev[current_device()].cfg[4];
    if (!configured) {
        const void* fns[4] = {(const void*)scan_kernel<false, false>, (const void*)scan_kernel<false, true>,
                              (const void*)scan_kernel<true, false>, (const void*)scan_kernel<true, true>};
        for (int i = 0; i < 4; ++i) {
            ES_CUDA_OK(cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanShared)));
            ES_CUDA_OK(cudaFuncSetAttribute(fns[i], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
        configured = 1;
    }
    const unsigned grid = (unsigned)blocks;
    const size_t smem = sizeof(ScanShared);
    cudaStream_t st = (cudaStream_t)stream;
    switch (variant) {
    case 0: scan_kernel<false, false><<<grid, 32, smem, st>>>(x, nclips, n, x_stride, corr, nc, ch, groups); break;
    case 1: scan_kernel<false, true><<<grid, 32, smem, st>>>(x, nclips, n, x_stride, corr, nc, ch, groups); break;
    case 2: scan_kernel<true, false><<<grid, 32, smem, st>>>(x, nclips, n, x_stride, corr, nc, ch, groups); break;
    default: scan_kernel<true, true><<<grid, 32, smem, st>>>(x, nclips, n, x_stride, corr, nc, ch, groups); break;
    }
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

// hist u32[rows][2048], spec f64[rows][6144], nspec u32[rows] (rows = clips * 4): K3's first pass formed by K2 while the
// correlation values are at hand; all three may be null (plain es_rx_ncc).  hist and nspec are zeroed here.
int es_rx_ncc_hist(const double* y, int nclips, int n, double* corr, uint32_t* hist, double* spec, uint32_t* nspec, void* stream)
{
    if (!g_rx_ready) { set_error("es_rx_ncc: call es_rx_set_filters first"); return ES_ENOTREADY; }
    const int nc = n - (PRE_L - 1);
    if (nclips <= 0 || nc <= 0) return ES_OK;
    const bool aux_on = hist && spec && nspec;
    const int ntiles = (nc + NCC_TILE - 1) / NCC_TILE;
    const size_t smem = 2 * (size_t)(ncc_pad(NCC_IN) + 1) * sizeof(double) + (aux_on ? NCC_HBINS * sizeof(unsigned int) : 0);
    int& configured = g_rxdev[current_device()].cfg[0];
    if (!configured) {
        const int big = (int)(2 * (size_t)(ncc_pad(NCC_IN) + 1) * sizeof(double) + NCC_HBINS * sizeof(unsigned int));
        ES_CUDA_OK(cudaFuncSetAttribute(ncc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        ES_CUDA_OK(cudaFuncSetAttribute(ncc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        configured = 1;
    }
    dim3 grid((ntiles + NCC_SPAN - 1) / NCC_SPAN, nclips * NBANDS);
    NccAux aux{hist, spec, nspec};
    if (aux_on) {
        ES_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)nclips * NBANDS * NCC_HBINS * sizeof(uint32_t), (cudaStream_t)stream));
        ES_CUDA_OK(cudaMemsetAsync(nspec, 0, (size_t)nclips * NBANDS * sizeof(uint32_t), (cudaStream_t)stream));
        ncc_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(y, n, nc, corr, aux);
    } else {
        ncc_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(y, n, nc, corr, aux);
    }
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_rx_ncc(const double* y, int nclips, int n, double* corr, void* stream)
{
    return es_rx_ncc_hist(y, nclips, n, corr, nullptr, nullptr, nullptr, stream);
}

static int g_peaks_general = 0;
void es_rx_peaks_force_general(int on) { g_peaks_general = on ? 1 : 0; }

// hist / spec / nspec: the outputs of es_rx_ncc_hist for the same corr (all null: K3 makes its own first pass)
int es_rx_peaks_hist(const double* corr, int nclips, int nc, const uint32_t* hist, const double* spec, const uint32_t* nspec,
                     int32_t* peaks, int32_t* npeaks, double* stats, void* stream)
{
    if (nclips <= 0) return ES_OK;
    int& configured = g_rxdev[current_device()].cfg[1];
    if (!configured) {
        ES_CUDA_OK(cudaFuncSetAttribute(peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PeakShared)));
        ES_CUDA_OK(cudaFuncSetAttribute(peaks2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PkUnion)));
        configured = 1;
    }
    NccAux aux{const_cast<uint32_t*>(hist), const_cast<double*>(spec), const_cast<uint32_t*>(nspec)};
    if (!(hist && spec && nspec)) aux = NccAux{nullptr, nullptr, nullptr};
    if (g_peaks_general)
        peaks_kernel<<<nclips * NBANDS, PK_THREADS, sizeof(PeakShared), (cudaStream_t)stream>>>(corr, nc, peaks, npeaks, stats);
    else
        peaks2_kernel<<<nclips * NBANDS, PK_THREADS, sizeof(PkUnion), (cudaStream_t)stream>>>(corr, nc, peaks, npeaks, stats, aux);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_rx_peaks(const double* corr, int nclips, int nc, int32_t* peaks, int32_t* npeaks, double* stats, void* stream)
{
    return es_rx_peaks_hist(corr, nclips, nc, nullptr, nullptr, nullptr, peaks, npeaks, stats, stream);
}

size_t es_rx_peaks_long_scratch_bytes(int nclips, int nc)
{
    const size_t ncb = (size_t)nclips * NBANDS;
    const size_t nblk = ((size_t)nc + NMS_BLOCK - 1) / NMS_BLOCK;
    return ncb * (sizeof(K3Meta) + 2048 * 4 + 4096 * 4 + (size_t)K3_CAP * 8 + (size_t)K3_CAP * 4 + nblk * (PEAK_LIMIT + 1) * 4) + 256;
}

struct K3Scratch {
    K3Meta* meta; unsigned int* hist1; unsigned int* hist2; double* buf; int32_t* top_idx; int32_t* blk_peaks; int32_t* blk_count;
    size_t off[7];
};
static K3Scratch k3_carve(void* scratch, int ncb, int nblk)
{
    K3Scratch k;
    unsigned char* p0 = (unsigned char*)scratch;
    unsigned char* p = p0;
    k.meta = (K3Meta*)p; k.off[0] = 0; p += ((size_t)ncb * sizeof(K3Meta) + 255) / 256 * 256;
    k.hist1 = (unsigned int*)p; k.off[1] = (size_t)(p - p0); p += (size_t)ncb * 2048 * 4;
    k.hist2 = (unsigned int*)p; k.off[2] = (size_t)(p - p0); p += (size_t)ncb * 4096 * 4;
    k.buf = (double*)p; k.off[3] = (size_t)(p - p0); p += (size_t)ncb * K3_CAP * 8;
    k.top_idx = (int32_t*)p; k.off[4] = (size_t)(p - p0); p += (size_t)ncb * K3_CAP * 4;
    k.blk_peaks = (int32_t*)p; k.off[5] = (size_t)(p - p0); p += (size_t)ncb * nblk * PEAK_LIMIT * 4;
    k.blk_count = (int32_t*)p; k.off[6] = (size_t)(p - p0);
    return k;
}

// Byte offsets inside the scratch of es_rx_peaks_long(_phase) and the layout of its per-row record, for callers that
// exchange the histograms / gathered values between GPUs (echoseal_b200/long_sharded.py):
// out = off(meta), off(hist1 u32[rows][2048]), off(hist2 u32[rows][4096]), off(buf f64[rows][cap]), off(top_idx),
//       sizeof(meta record), offsetof cnt, topcnt, overflow, topbin, med, mad, thr, cap
int es_rx_peaks_long_layout(int nclips, int n_range, long long* out)
{
    const int ncb = nclips * NBANDS;
    const int nblk = (n_range + NMS_BLOCK - 1) / NMS_BLOCK;
    const K3Scratch k = k3_carve(nullptr, ncb, nblk);
    out[0] = (long long)k.off[0]; out[1] = (long long)k.off[1]; out[2] = (long long)k.off[2]; out[3] = (long long)k.off[3];
    out[4] = (long long)k.off[4];
    out[5] = (long long)sizeof(K3Meta);
    out[6] = (long long)offsetof(K3Meta, cnt); out[7] = (long long)offsetof(K3Meta, topcnt);
    out[8] = (long long)offsetof(K3Meta, overflow); out[9] = (long long)offsetof(K3Meta, topbin);
    out[10] = (long long)offsetof(K3Meta, med); out[11] = (long long)offsetof(K3Meta, mad); out[12] = (long long)offsetof(K3Meta, thr);
    out[13] = (long long)K3_CAP;
    return ES_OK;
}

// One phase of the long-recording K3 on the index range [lo, hi) of rows stored with `nloc` values each; `nc_total`
// is the number of values of the WHOLE row (all ranks), which the rank arithmetic uses.  Between phases the caller
// may sum hist1 / hist2 over ranks (after phases 0, 1, 3, 4) and merge the gathered values + counts (after 2, 5):
//   0 clear, corr histogram            1 locate bins, sub-histogram          2 locate sub-bins, gather
//   3 median, top-bin indices, |corr-med| histogram   4 locate, sub-histogram   5 locate, gather
//   6 MAD + threshold, NMS per block, first 25 peaks of the range (+ local top-k candidates), overflow flag
int es_rx_peaks_long_phase(int phase, const double* corr, int nclips, int nloc, int lo, int hi, int nc_total,
                           void* scratch, size_t scratch_bytes,
                           int32_t* peaks, int32_t* npeaks, double* stats, int32_t* overflow_dev, void* stream)
{
    if (nclips <= 0 || hi <= lo) return ES_OK;
    if (lo < 0 || hi > nloc || nc_total < hi - lo) { set_error("es_rx_peaks_long_phase: bad range [%d,%d) of %d (total %d)", lo, hi, nloc, nc_total); return ES_EINVAL; }
    if (!scratch || scratch_bytes < es_rx_peaks_long_scratch_bytes(nclips, hi - lo)) { set_error("es_rx_peaks_long: scratch too small"); return ES_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const int ncb = nclips * NBANDS;
    const int nr = hi - lo;
    const int nblk = (nr + NMS_BLOCK - 1) / NMS_BLOCK;
    const K3Scratch k = k3_carve(scratch, ncb, nblk);
    K3Meta* meta = k.meta;
    const long long stride = nloc;
    dim3 grid((unsigned)(((long long)nr + K3_CHUNK - 1) / K3_CHUNK), ncb);
    switch (phase) {
    case 0:     // ---- median of corr (rtwm/detector.py:83)
        ES_CUDA_OK(cudaMemsetAsync(meta, 0, (size_t)ncb * sizeof(K3Meta), st));
        ES_CUDA_OK(cudaMemsetAsync(overflow_dev, 0, sizeof(int32_t), st));
        ES_CUDA_OK(cudaMemsetAsync(k.hist1, 0, (size_t)ncb * 2048 * 4, st));
        k3_hist_kernel<0, 1><<<grid, 1024, 0, st>>>(corr, stride, lo, hi, meta, k.hist1);
        break;
    case 1:
        k3_locate_kernel<1><<<ncb, 32, 0, st>>>(k.hist1, nc_total, meta, 1);
        ES_CUDA_OK(cudaMemsetAsync(k.hist2, 0, (size_t)ncb * 4096 * 4, st));
        k3_hist_kernel<0, 2><<<grid, 1024, 0, st>>>(corr, stride, lo, hi, meta, k.hist2);
        break;
    case 2:
        k3_locate_kernel<2><<<ncb, 32, 0, st>>>(k.hist2, nc_total, meta, 0);
        k3_gather_kernel<0><<<grid, 1024, 0, st>>>(corr, stride, lo, hi, meta, k.buf);
        break;
    case 3:
        k3_finish_kernel<0><<<ncb, 1024, 0, st>>>(k.buf, nc_total, meta);
        // candidates of the top-5 fallback (unconditional: no host round trip)
        k3_top_kernel<<<grid, 1024, 0, st>>>(corr, stride, lo, hi, meta, k.top_idx);
        // ---- MAD of |corr - med| and the threshold (rtwm/detector.py:84-86)
        ES_CUDA_OK(cudaMemsetAsync(k.hist1, 0, (size_t)ncb * 2048 * 4, st));
        k3_hist_kernel<1, 1><<<grid, 1024, 0, st>>>(corr, stride, lo, hi, meta, k.hist1);
        break;
    case 4:
        k3_locate_kernel<1><<<ncb, 32, 0, st>>>(k.hist1, nc_total, meta, 0);
        ES_CUDA_OK(cudaMemsetAsync(k.hist2, 0, (size_t)ncb * 4096 * 4, st));
        k3_hist_kernel<1, 2><<<grid, 1024, 0, st>>>(corr, stride, lo, hi, meta, k.hist2);
        break;
    case 5:
        k3_locate_kernel<2><<<ncb, 32, 0, st>>>(k.hist2, nc_total, meta, 0);
        k3_gather_kernel<1><<<grid, 1024, 0, st>>>(corr, stride, lo, hi, meta, k.buf);
        break;
    case 6:
        k3_finish_kernel<1><<<ncb, 1024, 0, st>>>(k.buf, nc_total, meta);
        // ---- NMS per index block, then the first 25 peaks / the fallback (rtwm/detector.py:87-99, 108-110)
        k3_nms_kernel<<<dim3(nblk, ncb), 1024, 0, st>>>(corr, nloc, lo, hi, meta, nblk, k.blk_peaks, k.blk_count);
        k3_collect_kernel<<<ncb, 1024, 0, st>>>(corr, stride, nc_total, meta, nblk, k.blk_peaks, k.blk_count, k.top_idx, peaks, npeaks, stats);
        k3_flag_kernel<<<(ncb + 255) / 256, 256, 0, st>>>(meta, ncb, overflow_dev);
        break;
    default:
        set_error("es_rx_peaks_long_phase: phase %d not in 0..6", phase);
        return ES_EINVAL;
    }
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

// long-recording form of es_rx_peaks (same outputs).  *overflow_dev (int32, device) is set non-zero when a gather
// buffer overflowed (degenerate data): the caller then falls back to es_rx_peaks.
int es_rx_peaks_long(const double* corr, int nclips, int nc, void* scratch, size_t scratch_bytes,
                     int32_t* peaks, int32_t* npeaks, double* stats, int32_t* overflow_dev, void* stream)
{
    if (nclips <= 0 || nc <= 0) return ES_OK;
    for (int phase = 0; phase <= 6; ++phase) {
        const int rc = es_rx_peaks_long_phase(phase, corr, nclips, nc, 0, nc, nc, scratch, scratch_bytes, peaks, npeaks, stats,
                                              overflow_dev, stream);
        if (rc != ES_OK) return rc;
    }
    return ES_OK;
}

int es_rx_frames(const double* y, int nclips, int n, const int32_t* peaks, const int32_t* npeaks, const uint8_t* hdr_pn,
                 float* mf_aligned, int32_t* llr_best_s, float* hdr_out, int32_t* hdr_best_s, void* stream)
{
    if (!g_rx_ready) { set_error("es_rx_frames: call es_rx_set_filters first"); return ES_ENOTREADY; }
    if (nclips <= 0) return ES_OK;
    dim3 grid(PEAK_LIMIT, nclips * NBANDS);
    frames_kernel<<<grid, FR_THREADS, 0, (cudaStream_t)stream>>>(y, n, 0, peaks, npeaks, hdr_pn, mf_aligned, llr_best_s, hdr_out, hdr_best_s);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_rx_frames_x(const float* x, int nclips, int n, long long x_stride, const int32_t* peaks, const int32_t* npeaks,
                   const uint8_t* hdr_pn, double* yframes, float* mf_aligned, int32_t* llr_best_s, float* hdr_out,
                   int32_t* hdr_best_s, void* stream)
{
    if (!g_rx_ready) { set_error("es_rx_frames_x: call es_rx_set_filters first"); return ES_ENOTREADY; }
    if (nclips <= 0) return ES_OK;
    frame_bandpass_kernel<<<nclips * NBANDS, 32, 0, (cudaStream_t)stream>>>(x, n, x_stride, peaks, npeaks, yframes);
    ES_CUDA_OK(cudaGetLastError());
    dim3 grid(PEAK_LIMIT, nclips * NBANDS);
    frames_kernel<<<grid, FR_THREADS, 0, (cudaStream_t)stream>>>(yframes, n, 1, peaks, npeaks, hdr_pn, mf_aligned, llr_best_s, hdr_out, hdr_best_s);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_rx_llr(const float* mf_aligned, const int32_t* item_peak, const uint8_t* pn_packed, int nitems, float* llr, void* stream)
{
    if (!g_rx_ready) { set_error("es_rx_llr: call es_rx_set_filters first"); return ES_ENOTREADY; }
    if (nitems <= 0) return ES_OK;
    const int nwork = nitems * 2;
    llr_kernel<<<(nwork + LLR_WARPS - 1) / LLR_WARPS, LLR_WARPS * 32, 0, (cudaStream_t)stream>>>(mf_aligned, item_peak, pn_packed, nwork, llr);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

int es_rx_resample(const void* x, int x_is_f64, int nclips, long long n_in, long long x_stride, int up, int down,
                   const double* taps /*[up][J] device*/, int J, long long n_pre_remove, long long n_out,
                   float* y /*[clips][n_out]*/, void* stream)
{
    if (nclips <= 0 || n_out <= 0) return ES_OK;
    if (up < 1 || down < 1 || J < 1) { set_error("es_rx_resample: bad up/down/J"); return ES_EINVAL; }
    dim3 grid((unsigned)((n_out + 255) / 256), nclips);
    if (x_is_f64)
        resample_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)x, n_in, x_stride, up, down, taps, J, n_pre_remove, n_out, y);
    else
        resample_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, n_in, x_stride, up, down, taps, J, n_pre_remove, n_out, y);
    ES_CUDA_OK(cudaGetLastError());
    return ES_OK;
}

}  // extern "C"
