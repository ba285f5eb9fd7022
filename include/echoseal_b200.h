/*
 * include/echoseal_b200.h — C ABI of libechoseal_b200.so, the drop-in boundary of the B200 hot path.
 *
 * The reference (PetarSt98/EchoSeal) has no FFI: its boundary is the Python API of rtwm.detector /
 * rtwm.embedder / rtwm.polar_fast.  Each entry point below names the reference code it replaces; the
 * Python drop-in classes in echoseal_b200/ (same names and signatures as the reference) bind these with
 * ctypes — see INTEGRATION.md for the stub a maintainer adds to rtwm/ itself.
 *
 * Conventions: plain pointers and sizes, no exceptions, no torch types.  `*_dev` / unqualified data
 * pointers of the es_scl_* / es_polar_* / es_rx_* / es_tx_* families are DEVICE pointers (the caller owns
 * every buffer); `stream` is a cudaStream_t passed as void*; kernels are launched asynchronously on it.
 * es_host_* functions take HOST pointers and never touch the GPU.  Return 0 on success, negative on
 * error (es_last_error() gives the text).  Bit-packed inputs are MSB-first per byte (np.packbits order).
 */
#ifndef ECHOSEAL_B200_H
#define ECHOSEAL_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int es_version(void);
const char* es_last_error(void);
int es_device_sm_count(void);

/* ---------------- Polar(1024,K)+CRC-8 codec ------------------------------------------------------ */
/* frozen mask (1 = frozen) of rtwm/fastpolar.py:219-229; K = info + CRC bits, K % 8 == 0. Host pointer. */
int es_polar_set_code(const uint8_t* frozen_host /*[1024]*/, int K);
/* persistent-grid geometry / scratch needed by es_scl_list */
int es_scl_grid_ctas(void);
int es_scl_ctas_per_sm(void);
size_t es_scl_scratch_bytes(void);
/* hard-decision fast path of PolarCode.decode (rtwm/fastpolar.py:261-276).
 * neg_mode 0: codeword w = row w; 1: codeword w = (w odd ? -1 : +1) * row (w/2)  (the detector's sign-flip
 * retry, rtwm/detector.py:182-190). */
int es_scl_hard(const float* llr /*[rows][1024]*/, int ncw, int neg_mode,
                uint8_t* hard_payload /*[ncw][(K-8)/8]*/, uint8_t* hard_crc /*[ncw]*/, void* stream);
/* CA-SCL list stage of PolarCode.decode (rtwm/fastpolar.py:278-330) + final ordering (:335): paths in
 * ascending-metric order with CRC flags; the caller applies the validator / selection rule (:335-359).
 * index (optional) lists the codeword ids to decode. list_size in 1..8. */
int es_scl_list(const float* llr, const int32_t* index, int ncw, int neg_mode, int list_size,
                void* scratch, size_t scratch_bytes,
                uint8_t* path_payload /*[ncw_total][L][(K-8)/8]*/, uint8_t* path_crc /*[ncw_total][L]*/,
                double* path_metric /*[ncw_total][L]*/, int32_t* npaths /*[ncw_total]*/, void* stream);
/* The same decode, also reporting per codeword the smallest RELATIVE gap between the last candidate kept and the first
 * one dropped over all pruning steps — (m[L] - m[L-1]) / m[L] in the sorted candidate list of rtwm/fastpolar.py:288-299
 * (+inf when no step pruned).  A margin within a few ulp (< 1e-11) marks a codeword whose survivor choice the
 * reference itself leaves to libm rounding (tie contract).  min_margin: [ncw_total] doubles. */
int es_scl_list_margin(const float* llr, const int32_t* index, int ncw, int neg_mode, int list_size,
                       void* scratch, size_t scratch_bytes,
                       uint8_t* path_payload, uint8_t* path_crc, double* path_metric, int32_t* npaths,
                       double* min_margin /*[ncw_total]*/, void* stream);
/* Wide-list decoder, list sizes 1..32 (rtwm/detector.py:27 accepts any list_size; the reference's quick test uses 32): the
 * arithmetic and rules of es_scl_list in the plain bit-by-bit form, one warp per codeword, one lane per path; about a
 * tenth of the SCL-8 rate.  Outputs as es_scl_list.  scratch: es_scl_wide_scratch_bytes() bytes of device memory. */
size_t es_scl_wide_scratch_bytes(void);
int es_scl_list_wide(const float* llr, int ncw, int neg_mode, int list_size, void* scratch, size_t scratch_bytes,
                     uint8_t* path_payload, uint8_t* path_crc, double* path_metric, int32_t* npaths, void* stream);
/* compaction of the CRC-passing candidates (hard decision = slot 0, list rank r = slot r+1): the inputs of the
 * validator callback of PolarCode.decode (rtwm/fastpolar.py:269-276, 335-349). Unordered; *counter may exceed cap. */
int es_scl_collect_hits(const uint8_t* hard_crc, const uint8_t* path_crc, const uint8_t* hard_payload,
                        const uint8_t* path_payload, long long ncw, int list_size, int cap,
                        int32_t* counter, int64_t* out_cw /*[cap]*/, int32_t* out_slot /*[cap]*/,
                        uint8_t* out_payload /*[cap][(K-8)/8]*/, void* stream);
/* PolarCode.encode / polar_fast.encode (rtwm/fastpolar.py:237-252, rtwm/polar_fast.py:26-53) */
int es_polar_encode(const uint8_t* payload /*[n][(K-8)/8]*/, int n, uint8_t* cw_bits /*[n][1024] or NULL*/,
                    uint32_t* cw_words /*[n][32] or NULL*/, void* stream);

/* ---------------- RX scan (rtwm/detector.py:56-152, 296-515) -------------------------------------- */
/* host constants: b/a of butter(4,[lo,hi],'band') per band (rtwm/utils.py:52-55), preamble templates
 * (rtwm/detector.py:67-69), matched-filter taps (rtwm/detector.py:260-294). Host pointers. */
int es_rx_set_filters(const double* bp_b /*[4][9]*/, const double* bp_a /*[4][9]*/, const double* tpl /*[4][63]*/,
                      const float* mf /*[4][192]*/, const int* mf_len /*[4]*/);
/* K1: y = lfilter(b, a, x) for the 4 bands (rtwm/detector.py:59-60) */
int es_rx_bandpass(const float* x /*[clips][x_stride]*/, int nclips, int n, long long x_stride,
                   double* y /*[clips][4][n]*/, void* stream);
/* test hook: 1 = K1 without its TMA tensor-store form (same chunk grid: both forms give the same bits) */
void es_rx_bandpass_force_plain(int on);
/* K2: cosine-normalised preamble correlation (rtwm/detector.py:76-79) */
int es_rx_ncc(const double* y, int nclips, int n, double* corr /*[clips][4][n-62]*/, void* stream);
/* es_rx_ncc that also forms K3's first pass while the correlation values are at hand: per (clip, band) row the 2048-bin
 * monotone histogram of corr (hist u32[rows][2048]) and the values of its six central bins (spec f64[rows][6144], count in
 * nspec u32[rows]); es_rx_peaks_hist then reads corr once instead of twice.  rows = clips * 4. */
int es_rx_ncc_hist(const double* y, int nclips, int n, double* corr, uint32_t* hist, double* spec, uint32_t* nspec, void* stream);
/* K1+K2 fused (rtwm/detector.py:59-60,75-79): band-pass and normalised correlation in one pass, the filtered signal
 * stays on the SM.  Same chunk grid and arithmetic as es_rx_bandpass + es_rx_ncc; windows that cross a chunk boundary see
 * the filter's 4e-13 warm-up truncation on the far side.  corr f64 [clips][4][n-62]. */
int es_rx_scan(const float* x /*[clips][x_stride]*/, int nclips, int n, long long x_stride, double* corr, void* stream);
/* K3: median/MAD threshold, NMS, first 25 peaks, top-5 fallback (rtwm/detector.py:83-99,107-110).
 * stats = med, mad, thr, used_fallback */
int es_rx_peaks(const double* corr, int nclips, int nc, int32_t* peaks /*[clips][4][25]*/,
                int32_t* npeaks /*[clips][4]*/, double* stats /*[clips][4][4]*/, void* stream);
int es_rx_peaks_hist(const double* corr, int nclips, int nc, const uint32_t* hist, const double* spec, const uint32_t* nspec,
                     int32_t* peaks, int32_t* npeaks, double* stats, void* stream);
/* test hook: on != 0 makes es_rx_peaks run every row through the general multi-pass form instead of the two-pass
 * form (the two give identical results; tests compare them) */
void es_rx_peaks_force_general(int on);
/* K3 for one LONG recording: same outputs, every pass spread over many CTAs; *overflow_dev != 0 means a
 * gather buffer overflowed (degenerate data) and the caller must fall back to es_rx_peaks */
size_t es_rx_peaks_long_scratch_bytes(int nclips, int nc);
int es_rx_peaks_long(const double* corr, int nclips, int nc, void* scratch, size_t scratch_bytes,
                     int32_t* peaks, int32_t* npeaks, double* stats, int32_t* overflow_dev, void* stream);
/* The same, one phase (0..6) at a time and on the index range [lo, hi) of rows stored with nloc values each;
 * nc_total = number of values of the WHOLE row over all ranks.  For ONE recording split in time over several GPUs
 * (SURVEY 8e): between phases the caller sums the histograms over the ranks (after phases 0, 1, 3, 4) and merges the
 * gathered values and their counts (after 2, 5); es_rx_peaks_long_layout gives the byte offsets of those buffers and
 * of the per-row record inside the scratch.  peaks are indices into the local rows. */
int es_rx_peaks_long_phase(int phase, const double* corr, int nclips, int nloc, int lo, int hi, int nc_total,
                           void* scratch, size_t scratch_bytes,
                           int32_t* peaks, int32_t* npeaks, double* stats, int32_t* overflow_dev, void* stream);
int es_rx_peaks_long_layout(int nclips, int n_range, long long* out /*[14]*/);
/* K4: per peak header decode (rtwm/detector.py:452-515) + matched filter / shift search of _llr
 * (rtwm/detector.py:322-383). hdr_out = ok (-1: no frame), val, score, margin */
int es_rx_frames(const double* y, int nclips, int n, const int32_t* peaks, const int32_t* npeaks,
                 const uint8_t* hdr_pn /*[clips][16]*/, float* mf_aligned /*[clips][4][25][1024]*/,
                 int32_t* llr_best_s, float* hdr_out /*[clips][4][25][4]*/, int32_t* hdr_best_s, void* stream);
/* K4 from the audio itself: the candidate frames are band-passed on their own (768-sample zero-state warm-up, exact from
 * the clip start) into yframes f64 [clips][4][25][1215] (scratch), then decoded as es_rx_frames does. */
int es_rx_frames_x(const float* x /*[clips][x_stride]*/, int nclips, int n, long long x_stride, const int32_t* peaks,
                   const int32_t* npeaks, const uint8_t* hdr_pn /*[clips][16]*/, double* yframes,
                   float* mf_aligned /*[clips][4][25][1024]*/, int32_t* llr_best_s, float* hdr_out /*[clips][4][25][4]*/,
                   int32_t* hdr_best_s, void* stream);
/* K5: despread + LLR scaling for (peak, counter) items, both PN variants (rtwm/detector.py:306-314,384-414) */
int es_rx_llr(const float* mf_aligned, const int32_t* item_peak /*[items]*/, const uint8_t* pn_packed /*[items][152]*/,
              int nitems, float* llr /*[2*items][1024]*/, void* stream);

/* K9: resample_to / scipy.signal.resample_poly(x, up, down) (rtwm/utils.py:58-66): taps = host-designed Kaiser(5)
 * low-pass in polyphase order [up][J] (device pointer), fp64 accumulation, float32 out */
int es_rx_resample(const void* x, int x_is_f64, int nclips, long long n_in, long long x_stride, int up, int down,
                   const double* taps, int J, long long n_pre_remove, long long n_out, float* y, void* stream);

/* ---------------- TX (rtwm/embedder.py:44-151) ----------------------------------------------------- */
int es_tx_set_filters(const double* bp_b /*[4][9]*/, const double* bp_a /*[4][9]*/, const uint8_t* preamble_bits /*[63]*/);
/* K7: _make_frame_chips (rtwm/embedder.py:78-151) for a batch of frames */
int es_tx_frames(const uint8_t* payload /*[F][55]*/, const uint8_t* pn /*[F][152]*/, const uint8_t* hdr_pn /*[F][16]*/,
                 const int32_t* band /*[F]*/, const int32_t* ctr_lo16 /*[F]*/, int nframes,
                 float* chips /*[F][1215]*/, void* stream);
/* K8: the mix of process() (rtwm/embedder.py:50-75), one block per stream */
int es_tx_mix(const float* x /*[S][blk]*/, const float* chips /*[S][blk]*/, int nstreams, int blk,
              double alpha, double floor_scale, float* out /*[S][blk]*/, float* scale_out /*[S] or NULL*/, void* stream);

/* ---------------- host feeder (rtwm/crypto.py, rtwm/utils.py:27-36,83-132) — HOST pointers ---------- */
void* es_host_keys_new(const uint8_t* keys /*[nkeys][32]*/, int nkeys, int nthreads);
void es_host_keys_free(void* bank);
int es_host_threads(void* bank);
int es_host_hdr_pn(void* bank, const int32_t* key_idx, int n, uint8_t* out /*[n][16]*/);
int es_host_hop(void* bank, int key, uint32_t lo, uint32_t hi, uint8_t* out);               /* choose_band */
int es_host_pn(void* bank, int key, const uint64_t* ctrs, int n, uint8_t* out /*[n][152]*/); /* pn_bits */
/* candidate counters + 400-try budget + PN bits (rtwm/detector.py:105-151); item_peak == NULL: count only */
int64_t es_host_rx_enumerate(void* bank, const int32_t* key_idx, int nb, int n_samples,
                             const int32_t* peaks, const int32_t* npeaks, const float* hdr,
                             int32_t* band_count /*[nb][4]*/, int64_t* item_offset /*[nb+1]*/,
                             int32_t* item_peak, uint32_t* item_ctr, int32_t* item_clip, uint8_t* pn);
/* AEAD validator, magic / counter checks, nonce latch, band and attempt order (rtwm/detector.py:44-53,161-233) */
int es_host_rx_validate(void* bank, const int32_t* key_idx, int nb,
                        const int32_t* band_count, const int64_t* item_offset, const uint32_t* item_ctr,
                        const int64_t* hit_cw, const int32_t* hit_slot, const uint8_t* hit_payload /*[nhits][55]*/,
                        int64_t nhits, uint8_t* nonce_state /*[nb][9]*/, uint8_t* verdict /*[nb]*/,
                        uint8_t* plaintext /*[nb][27] or NULL*/);
/* _build_payload + pn_bits + choose_band for a batch of frames (rtwm/embedder.py:82-119,153-168) */
int es_host_tx_prepare(void* bank, const int32_t* key_idx, const uint32_t* ctr, const uint8_t* session_nonce /*[F][8]*/,
                       const uint8_t* rnd /*[F][23]*/, int F, uint8_t* payload /*[F][55]*/, uint8_t* pn /*[F][152]*/,
                       uint8_t* hdr_pn /*[F][16]*/, int32_t* band, int32_t* ctr_lo16);

#ifdef __cplusplus
}
#endif
#endif
