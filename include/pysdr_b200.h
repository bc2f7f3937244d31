/* pysdr_b200.h — C ABI of the B200-native receive-DSP path (libpysdr_b200.so).
 *
 * The reference (aa2il/pySDR) has no FFI: its seam is the Python object protocol of the external
 * module `sig_proc` (reference receiver.py:45 `import sig_proc as dsp`).  Each entry point below names
 * the reference call it stands behind.  All functions return 0 on success, <0 on error
 * (message via pysdr_last_error()); nothing throws across the boundary.  Device buffers are owned by
 * the caller (PyTorch); the library owns only handles, coefficient tables and scratch.  One handle is
 * single-threaded; different handles are independent.  `stream` is a cudaStream_t passed as void*.
 */
#ifndef PYSDR_B200_H
#define PYSDR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYSDR_MAX_RX 128        /* receivers per bank; the reference stops at MAX_RX = 6 (params.py:33), the many-channel
                                   bank (BASELINE config 5) runs its audio-rate stages in groups of up to this many */
#define PYSDR_AGC_NB 8

/* demodulator kinds: reference Tables.py:34 MODES */
enum {
    PYSDR_MODE_AM = 0,          /* envelope detector                     */
    PYSDR_MODE_USB = 1,         /* also SSB                              */
    PYSDR_MODE_LSB = 2,
    PYSDR_MODE_CW = 3,
    PYSDR_MODE_IQ = 4,          /* also RTTY (IQ feed, receiver.py:286-290) */
    PYSDR_MODE_NFM = 5,
    PYSDR_MODE_RAW = 6,         /* Re{resampler output}, no demod filter: second stage of WFM (gui.py:1703,1759-1762) */
    PYSDR_MODE_AMSYNC = 7       /* AM-Synch: carrier PLL, in-phase arm (Tables.py:34; demod.am_pll, receiver.py:649) */
};

enum {
    PYSDR_OK = 0,
    PYSDR_ERR_ARG = -1,
    PYSDR_ERR_CUDA = -2,
    PYSDR_ERR_CAPACITY = -3,
    PYSDR_ERR_ALIGN = -4,       /* process() must start on an IN_CHUNK_SIZE boundary of the stream */
    PYSDR_ERR_STATE = -5
};

const char *pysdr_last_error(void);
int pysdr_version(void);

/* ---- a3: NCO helpers (dsp.signal_generator; reference receiver.py:822,552-553; gui.py:1928) ---- */
uint64_t pysdr_freq_to_phase_inc(double f_hz, double fs_hz);
double pysdr_phase_inc_to_freq(uint64_t inc, double fs_hz);
/* x[i] *= exp(-j*2*pi*(acc0 + i*inc)/2^64), in place or out of place; complex64 device pointers. */
int pysdr_quad_mixer(const void *d_x, void *d_y, int64_t n, uint64_t acc0, uint64_t inc, void *stream);

/* CS16 sources (reference receiver.py:614-617): complex64 out[i] = scale*(in[2i] + j in[2i+1]), scale = 1/2048 there.
 * d_in: device int16[2n], d_out: device complex64[n]. */
int pysdr_cs16_to_cf32(const void *d_in, void *d_out, int64_t n, double scale, void *stream);
/* rx.auto_mute(x) building block (reference receiver.py:238-245): *d_out = mean(|x|^2), float32 device. */
int pysdr_mean_power(const void *d_x, int64_t n, float *d_out, void *stream);

/* dsp.convolver(h).convolve_fast(x) building block (reference receiver.py:862,216): valid FIR,
 * out[i] = sum_j taps[j]*src[i+(L-1)-j]; src = L-1 history samples then n new; float32 or complex64. */
int pysdr_fir_valid(const void *d_src, int src_is_complex, const float *taps_host, int L, int64_t n,
                    void *d_out, void *stream);

/* ---- a4..a7: receiver bank = all dsp.Receiver objects fed by one IQ stream -------------------- */
typedef struct pysdr_bank pysdr_bank;

typedef struct {
    double srate;               /* P.SRATE                                   params.py:222-231 */
    int32_t up, down;           /* P.UP, P.DOWN = up_dn(SRATE,FS_OUT)        params.py:405     */
    int64_t in_chunk;           /* P.IN_CHUNK_SIZE (AGC / DC-removal block)  params.py:444     */
    int32_t n_rx;               /* P.NUM_RX                                                    */
    int32_t filt_len;           /* P.FILT_LEN, resampler prototype length    params.py:345     */
    int32_t af_len;             /* demodulator FIR length                                     */
    int64_t max_in;             /* largest n_in one process() call may carry                  */
} pysdr_bank_config;

int pysdr_bank_create(const pysdr_bank_config *cfg, pysdr_bank **out);
int pysdr_bank_destroy(pysdr_bank *b);
/* stream reset: histories, AGC, absolute indices (new Receiver objects) */
int pysdr_bank_reset(pysdr_bank *b);

/* rx.lo.change_freq(f)           reference gui.py:1938, receiver.py:112 */
int pysdr_bank_set_lo(pysdr_bank *b, int rx, uint64_t phase_inc);
/* rx.dec.h = rx.dec.filter_bank[idx]   reference gui.py:1713, receiver.py:127 ; h: host float32[n] */
int pysdr_bank_set_dec_taps(pysdr_bank *b, int rx, const float *h, int n);
/* P.MODE / P.AF_BW / P.AF_FILTER_NUM / P.BFO as read by demod_data at call time
 * (reference receiver.py:115,130-131; gui.py:1698,1756-1757).  taps: host float32, n real taps, or n
 * interleaved (re,im) pairs when is_complex. */
int pysdr_bank_set_demod(pysdr_bank *b, int rx, int mode, const float *taps, int n, int is_complex,
                         uint64_t bfo_phase_inc);
/* WFM2 (Tables.py:34; BASELINE config 4): a 3-row resampler bank fed with the real FM multiplex — row 0 LO 0 (L+R),
 * row 1 LO 38 kHz (L-R), row 2 LO 19 kHz (pilot), all rows in IQ mode — turns rows 0/1 into L/R after the AF
 * filters: u = z2/|z2|, D = 2 Re{z1 conj(u)^2} (0 when |z2| <= pilot_min), L = Re z0 + D, R = Re z0 - D; one AGC
 * (block peak = max of both) serves the pair; process_back returns L in row 0 and R in row 1 as float32. */
int pysdr_bank_set_stereo(pysdr_bank *b, int on, double pilot_min);
/* rx.demod.am_pll.reset()        reference receiver.py:649 ; loop state out2 = {phi [rad], w [rad/sample]} */
int pysdr_bank_pll_reset(pysdr_bank *b, int rx);
int pysdr_bank_pll_get(pysdr_bank *b, int rx, double out2[2], void *stream);
/* rx.agc.reset()                 reference receiver.py:648 */
int pysdr_bank_agc_reset(pysdr_bank *b, int rx);
int pysdr_bank_agc_config(pysdr_bank *b, int rx, double ref, double beta);
/* rx.agc.{agc,gain,maxbuf,ref,err}   reference watchdog.py:298-302 ; synchronises the stream */
int pysdr_bank_agc_get(pysdr_bank *b, int rx, double out5[5], void *stream);

/* # of audio samples the next process(n_in) will emit per receiver (index arithmetic only). */
int64_t pysdr_bank_n_out(const pysdr_bank *b, int64_t n_in);
int64_t pysdr_bank_position(const pysdr_bank *b);          /* absolute input index n0 */
int64_t pysdr_bank_n_blocks(const pysdr_bank *b, int64_t n_in);

/* rx.demod_data(x) for every receiver of the bank on one shared read of x
 * (reference receiver.py:235 via :724-725).
 *   d_iq      complex64[n_in] device; n_in may span many IN_CHUNK_SIZE blocks
 *   halo_in_place  !=0: d_iq[-(lp-1)..-1] are valid preceding samples and replace the internal filter
 *                  memory (time-sharded captures); 0: use the memory carried from the previous call
 *   d_iq_bb   complex64[n_rx][out_stride]  -> rx.iq        (may be NULL)
 *   d_am      float32  [n_rx][2*out_stride] -> rx.am  (real modes fill the first n_out floats of a row;
 *             IQ mode fills n_out interleaved complex)
 *   d_am_dc   like d_am: `am - mean(am)` per block for AM/USB (reference receiver.py:250-252), NULL to skip
 *   n_out     host, receives the # of audio samples per receiver
 */
int pysdr_bank_process(pysdr_bank *b, const void *d_iq, int64_t n_in, int halo_in_place,
                       void *d_iq_bb, float *d_am, float *d_am_dc, int64_t out_stride,
                       int64_t *n_out, void *stream);

/* The reference-facing per-chunk call with HOST buffers (what `for irx: demodulate_data(P, self.x, irx)` amounts to,
 * reference receiver.py:724-725): h_iq complex64[n_in] in host memory (page-locked memory is read by DMA directly, pageable
 * memory goes through a pinned staging copy); on return *h_am / *h_iq_bb / *h_am_dc point at rows of a page-locked result
 * block owned by the bank, valid until the next call: row r starts *row_floats floats after row r-1 and holds n_out
 * float32 (real modes) or n_out complex64 (IQ / RTTY rows, and every *h_iq_bb row).  One call = upload, kernels, download,
 * one synchronisation of `stream`.  want_iq / want_dc = 0 skip the baseband / DC-removed downloads. */
int pysdr_bank_process_host(pysdr_bank *b, const void *h_iq, int64_t n_in, int want_iq, int want_dc, float **h_am,
                            void **h_iq_bb, float **h_am_dc, int64_t *row_floats, int64_t *n_out, void *stream);

/* Device copy of the chunk the last pysdr_bank_process_host call uploaded (e.g. for the auto-mute power detector). */
int pysdr_bank_host_chunk_ptr(pysdr_bank *b, void **d_in);

/* Split form for time-sharded multi-GPU runs: front = everything up to the per-block AGC peaks
 * (no cross-block dependency), back = AGC recursion + gain/DC application.  Between the two the
 * caller may all-gather d_peaks across ranks (NCCL) and pass the peaks of ALL earlier blocks.
 *   d_peaks: float32[n_rx][n_blocks(n_in)] written by front (device).
 *   back:  d_prev_peaks float32[n_rx][n_prev] = peaks of the n_prev blocks preceding this call's
 *          first REAL block (NULL/0: continue from the carried AGC state);
 *          skip_blocks = leading blocks of the front call that only warmed the filter memories (their audio and
 *          peaks are not valid and are excluded from the recursion). */
int pysdr_bank_process_front(pysdr_bank *b, const void *d_iq, int64_t n_in, int halo_in_place,
                             void *d_iq_bb, int64_t out_stride, float *d_peaks, int64_t *n_out,
                             void *stream);
int pysdr_bank_process_back(pysdr_bank *b, const float *d_prev_peaks, int64_t n_prev, int64_t skip_blocks,
                            float *d_am, float *d_am_dc, int64_t out_stride, void *stream);
/* O(1) AGC carry between time shards (replaces the all-gather of every block peak).  The per-block update
 * gain <- min(w_b, beta w_b + (1-beta) gain) is closed under composition, and w_b depends on the last 8 peaks only, so a
 * shard of n >= 8 blocks is summarised by PYSDR_AGC_SUMMARY_LEN doubles per receiver:
 *     [0..2]  (A, C, D): gain_out = min(A, C + D gain_in) over the shard's blocks 7 .. n-1 (own peaks only)
 *     [3..9]  the shard's first 7 block peaks (their w_b needs the previous shard's last peaks)
 *     [10..17] the shard's last 8 block peaks (the peak buffer it hands on), oldest first
 *     [18]    n
 * agc_summary: after process_front; writes d_summary[n_rx][LEN] (device) for this call's blocks skip_blocks .. n-1.
 * agc_enter:   sets the AGC state the NEXT process_back(NULL, 0, skip, ...) continues from, by running the summaries of
 *              the n_before earlier shards (d_summaries[n_before][n_rx][LEN], device, in stream order) from the reset
 *              state: 7 plain updates + one composed step per shard.  n_before = 0 gives the reset state. */
#define PYSDR_AGC_SUMMARY_LEN 19
#define PYSDR_XCHG_DEPTH 16      /* ring depth of the peer-memory exchange buffer (steps a fast rank may run ahead) */
#define PYSDR_XCHG_MAX_WORLD 16
int pysdr_bank_agc_summary(pysdr_bank *b, int64_t skip_blocks, double *d_summary, void *stream);
int pysdr_bank_agc_enter(pysdr_bank *b, const double *d_summaries, int n_before, void *stream);
/* process_back with the entry state taken from the summaries of the n_before earlier shards (agc_enter fused into the same
 * launch as the scan and the gain application). */
int pysdr_bank_process_back_carry(pysdr_bank *b, const double *d_summaries, int n_before, int64_t skip_blocks,
                                  float *d_am, float *d_am_dc, int64_t out_stride, void *stream);
/* The same carry over NVLink PEER MEMORY instead of a collective library call: every rank owns one buffer of
 * pysdr_xchg_bytes(world, n_rx) bytes that is mapped into all peers (torch.distributed._symmetric_memory), zero-initialised.
 * peer_bases[q] = device address of rank q's buffer as mapped into THIS process (q = rank: the local buffer).
 *   agc_summary_push : after process_front; computes this rank's summaries and stores them into slot seq % PYSDR_XCHG_DEPTH of
 *                      every LATER rank's buffer, then raises this rank's flag for step seq there.  seq = 1, 2, 3, ... on
 *                      all ranks, every value exactly once per buffer.  Before reusing a slot the kernel waits for the later
 *                      ranks' acknowledgement of step seq - PYSDR_XCHG_DEPTH, so a fast rank runs at most DEPTH steps ahead.
 *   process_back_xchg: process_back whose scanner CTAs wait (in the kernel) for the flags of the ranks < rank, enter from
 *                      their summaries in the local buffer and acknowledge the step in the earlier ranks' buffers.
 * A rank waits for earlier ranks' data and later ranks' acknowledgements of older steps only (no cycle); a wait that lasts
 * 10 s traps (dead peer) instead of hanging the GPU.  No host synchronisation and no collective library on this path. */
int64_t pysdr_xchg_bytes(int world, int n_rx);
int pysdr_bank_agc_summary_push(pysdr_bank *b, int64_t skip_blocks, const uint64_t *peer_bases, int world, int rank, uint64_t seq,
                                void *stream);
int pysdr_bank_process_back_xchg(pysdr_bank *b, const uint64_t *peer_bases, int world, int rank, uint64_t seq, int64_t skip_blocks,
                                 float *d_am, float *d_am_dc, int64_t out_stride, void *stream);
/* A whole time shard in the same three launches as a single-GPU step: process_front, then ONE back kernel that computes the
 * block peaks, pushes this rank's summaries (agc_summary_push), waits for the earlier ranks' (process_back_xchg), scans and
 * applies the gains.  skip_blocks = warm-up blocks at the head of d_iq whose outputs rebuild the filter memories only. */
int pysdr_bank_process_shard_xchg(pysdr_bank *b, const void *d_iq, int64_t n_in, int halo_in_place, void *d_iq_bb,
                                  const uint64_t *peer_bases, int world, int rank, uint64_t seq, int64_t skip_blocks,
                                  float *d_am, float *d_am_dc, int64_t out_stride, int64_t *n_out, void *stream);
/* Testing: on != 0 runs the tail as stand-alone kernels (block peaks, AGC scan, gain application, seek reset) instead of the
 * fused back kernel (one co-resident grid, two grid barriers). */
int pysdr_bank_force_unfused(pysdr_bank *b, int on);

/* Move the stream position without processing (time shards): n0 must be a multiple of in_chunk.
 * LO/BFO accumulators follow; filter memories are cleared; n0 == 0 also resets the AGC (stream restart).
 * Asynchronous on `stream`. */
int pysdr_bank_seek(pysdr_bank *b, int64_t n0_abs, void *stream);

/* Checkpoint = carry state (filter memories, AGC, indices) as a flat byte blob. */
int64_t pysdr_bank_state_size(const pysdr_bank *b);
int pysdr_bank_get_state(pysdr_bank *b, void *host_blob, int64_t size, void *stream);
int pysdr_bank_set_state(pysdr_bank *b, const void *host_blob, int64_t size, void *stream);

/* Which K1 variant the next process() will use: 0 generic, 1 tap-stationary fast path. */
int pysdr_bank_k1_variant(const pysdr_bank *b);
/* Tensor-core K1 (k1_mma.cu: the mix + polyphase decimation as a split-TF32 GEMM on tcgen05, samples streamed by TMA into
 * tensor memory).  mode 0: never; 1 (default): calls whose interior has at least 8192 super-periods (whole captures, long
 * segments) — short per-chunk calls keep the tap-stationary FP32 kernel; 2: whenever the geometry and alignment allow.
 * The two kernels agree to ~1e-6 of peak, not bit for bit: pin mode 0 where chunked and whole-capture K1 outputs must be
 * identical.  k1_last: the kernel the last call ran (0 generic, 1 tap-stationary, 2 tensor-core interior + edge tiles).
 *
 * Banks of 16 or more receivers (the many-channel receivers of BASELINE config 5, any set of offsets) have a second
 * tensor-core kernel, k1_chan.cu: channels are the GEMM's columns, the output instants of one polyphase class its rows, the
 * channels' folded taps stream from L2 as the B operand.  The same mode switch governs it (mode 1: calls with at least 2048
 * interior super-periods); k1_last reports 3 when it ran. */
int pysdr_bank_set_k1_mma(pysdr_bank *b, int mode);
int pysdr_bank_k1_mma_available(const pysdr_bank *b);
int pysdr_bank_k1_last(const pysdr_bank *b);
int pysdr_bank_force_generic(pysdr_bank *b, int on);
/* Test hook of k1_chan.cu (host arithmetic only, no device): the plan, the per-call geometry and the tap images exactly as
 * the kernel gets them, so that the CPU suite can emulate the contraction in numpy.  g_host: folded taps
 * complex64[n_rx][up][lp_pad]; x_addr: the device address the capture would have (alignment only).  out[0..15]: used, q_a,
 * out_lo, out_hi, n_steps, ngroups, nch, N, ncls, S, n_tiles; out[16 + 8 cls ..]: phase, parity, sample offset, image, rows,
 * first sample of row 0.  Returns the image's float count (copied into img when img_cap allows) or a negative error code. */
int64_t pysdr_k1chan_debug_plan(int up, int down, int lp, int n_rx, const float *g_host, int lp_pad, int64_t n0, int64_t n_in,
                                int64_t m0, int64_t n_out, uint64_t x_addr, int64_t min_rows, int64_t *out, float *img,
                                int64_t img_cap);
/* K1-only bank (WFM video stage: LO + FIR at the RF rate, UP = DOWN = 1): process() stops after K1; its output is the
 * new-sample part of the complex memory (pysdr_bank_c_memory) and, when given, d_iq_bb. */
int pysdr_bank_set_k1_only(pysdr_bank *b, int on);
/* on != 0: the caller promises that every input sample has a zero imaginary part (the WFM resampler rows are fed with the
 * real FM discriminator output stored as complex64): the tap-stationary K1 then skips the Im-x half of its FMAs. */
int pysdr_bank_set_real_input(pysdr_bank *b, int on);
/* 3-point FM discriminator at any rate (reference sigs/nfm.m:123-127) with two carried samples:
 * out[n] = (Im(conj(y[n-1]) * (y[n] - y[n-2])), 0) as complex64; d_prev2: complex64[2] in/out. */
int pysdr_fm_disc(const void *d_y, int64_t n, void *d_prev2, void *d_out, void *stream);
/* WFM / WFM2 video stage at the RF rate ("demodulate first, then resample", reference gui.py:1703,1759-1762): LO mix +
 * video FIR (rx.demod.wfm_video.h, gui.py:1704) + 3-point FM discriminator (sigs/nfm.m:123-127) fused in one overlap-save
 * FFT-convolution kernel.
 *   pysdr_fir_spectrum : d_taps_c64[L] complex64 (device) -> d_H, the filter's spectrum in the transform's position order
 *                        with 1/N folded in (4096 complex64 for L <= 2047, else 8192).  For the video stage the taps are
 *                        h[j] * e^{+j 2 pi f j / fs} (LO folded in, host float64).
 *   pysdr_wfm_video_disc: d_x complex64[n_in]; d_hist complex64[L+1] = the raw samples preceding d_x (zeros at the stream
 *                        origin), updated to the last L+1 samples of the call; d_prev2 complex64[2][2]: slot prev2_slot holds
 *                        the last two video-filter outputs of the previous call (zeros at the origin), the new pair is written
 *                        to the other slot (the caller toggles prev2_slot after every call with n_in > 0); acc0 = LO phase
 *                        accumulator at d_x[0], inc per sample (pysdr_freq_to_phase_inc); d_fm complex64[n_in] receives (fm, 0). */
int pysdr_fir_spectrum(const void *d_taps_c64, int L, void *d_H, void *stream);
int pysdr_wfm_video_disc(const void *d_x, int64_t n_in, void *d_hist, void *d_prev2, int prev2_slot, const void *d_H, int L,
                         uint64_t acc0, uint64_t inc, void *d_fm, void *stream);
/* AF filter variant: default = overlap-save FFT convolution in shared memory; on != 0 forces the direct-form FIR. */
int pysdr_bank_force_direct_fir(pysdr_bank *b, int on);
/* On-stream stage timing for bench.py's roofline: out4 = {sum K1 ms, sum rest-of-front ms, sum back ms,
 * # process calls} since the last get; get synchronises `stream`. */
int pysdr_bank_set_timing(pysdr_bank *b, int on);
int pysdr_bank_get_timing(pysdr_bank *b, double out4[4], void *stream);
/* # of kernels launched by this handle so far (bench.py's gpu_launches). */
int64_t pysdr_bank_launch_count(const pysdr_bank *b);
/* The bank's complex memory: row r = [hist_len carried baseband samples | the n_out samples of the last call], complex64,
 * rows row_stride apart.  The new-sample part of a row IS rx.iq of the last call (valid until the next one), so a host
 * that passes d_iq_bb = NULL to process / process_front reads rx.iq from here without a second copy being written —
 * except for receivers in PYSDR_MODE_AMSYNC, whose new samples are de-rotated in place by the carrier loop. */
int pysdr_bank_c_memory(pysdr_bank *b, void **d_ptr, int64_t *row_stride, int32_t *hist_len);
/* AGC trace of the last process / process_back call (what reference watchdog.py:298-302 prints, per block instead of
 * per tick): host_peaks / host_gains receive [n_rx][*n_blocks] float32 — the per-IN_CHUNK_SIZE-block peak of the pre-AGC
 * audio and the gain applied to that block.  capacity = floats per row the caller provides.  Synchronises `stream`. */
int pysdr_bank_agc_trace(pysdr_bank *b, float *host_peaks, float *host_gains, int64_t capacity, int64_t *n_blocks,
                         void *stream);
/* Many-channel operation (wola.cu): the bank's complex memory can live in caller-owned device memory (n_rx rows,
 * row_stride >= the bank's own stride, zero-initialised, not freed by the bank), and K1 can be left to the caller, who then
 * writes the new baseband samples of every row at [hist_len, hist_len + n_out) on the same stream before process(). */
int pysdr_bank_adopt_c_memory(pysdr_bank *b, void *d_ptr, int64_t row_stride);
int pysdr_bank_set_k1_external(pysdr_bank *b, int on);

/* Many channels on a uniform raster (BASELINE config 5; beyond the reference's MAX_RX): baseband IQ of n_ch channel
 * receivers whose offsets are f0 + c*df with df/fs = a/3125 in lowest terms, from ONE windowing pass and ONE 3125-point
 * inverse DFT per output instant (wola.cu) — same indexing contract and arithmetic definition as the per-receiver K1:
 *   out[c][i] = e^{-j th_c(n_m)} sum_j G0[p_m][j] e^{+j 2 pi (a c) j/3125} x[n_m - j],  m = m0 + i, n_m = (m*down)/up, p_m = (m*down)%up
 * d_x[0] is absolute sample n0, n_in samples; the n_before samples preceding it are at d_hist (NULL: in place, directly in
 * front of d_x); anything earlier reads as zero.
 * d_g0: complex64[up][lp] folded taps of channel 0 (lp <= 625); d_pos[c] = base-5 digit reversal of (a*c) mod 3125;
 * d_inc[c]: 64-bit LO phase increment of channel c (phase 0 at absolute sample 0); d_out: complex64[n_ch][out_stride]. */
int pysdr_wola_channelize(const void *d_x, const void *d_hist, int64_t n0, int64_t n_before, int64_t n_in, int64_t m0, int64_t n_out,
                          int32_t up, int32_t down, int32_t lp, const void *d_g0, int32_t n_ch, const int32_t *d_pos,
                          const uint64_t *d_inc, void *d_out, int64_t out_stride, void *stream);

/* ---- a13: scipy.signal.lfilter(b,a,x,zi) with carried state (reference sigs/iir.py:90-105) ----
 * float64 direct-form-II-transposed, evaluated as a block-parallel linear scan.
 *   b,a: host float64[nb],[na]; d_x,d_y: device float32[n] (n_ch rows of `stride`);
 *   d_zi: device float64[n_ch][order] in/out (order = max(na,nb)-1). */
/* 0 = auto (block scan unless chaining would be ill-conditioned: max|Phi^B| > 1e3), 1 = scan, 2 = sequential */
int pysdr_lfilter_set_mode(int mode);
int pysdr_lfilter(const double *b, int nb, const double *a, int na, const float *d_x, float *d_y,
                  int64_t n, int n_ch, int64_t stride, double *d_zi, void *stream);

/* Elementwise helpers of the squelch detector (reference sigs/squelch.m:125-128 |z|, :141 sq1./sq2). */
int pysdr_abs_f32(const float *d_x, float *d_y, int64_t n, void *stream);
int pysdr_ratio_f32(const float *d_a, const float *d_b, float *d_r, float floor_v, int64_t n, void *stream);

/* ---- a11: dsp.spectrum(fs,chunk,NFFT,overlap) (reference Plotting.py:376-377,462) --------------- */
typedef struct pysdr_psd pysdr_psd;
int pysdr_psd_create(int32_t chunk_size, int32_t nfft, int32_t hop, const float *window_host,
                     pysdr_psd **out);
int pysdr_psd_destroy(pysdr_psd *p);
/* Frames k = 0..n_frames-1 start at k*hop, length chunk_size, zero-padded to nfft; every `navg`
 * consecutive frames are averaged into one output line: d_out float32[n_lines][nfft], fftshifted,
 * 10*log10 when dB.  complex64 input when is_complex else float32.  Returns # lines via n_lines. */
int pysdr_psd_lines(pysdr_psd *p, const void *d_x, int64_t n, int is_complex, int32_t navg, int dB,
                    float *d_out, int64_t *n_lines, void *stream);
int64_t pysdr_psd_launch_count(const pysdr_psd *p);
/* Quarter-symbol FFT filterbank of the RTTY executive (reference rtty.py:784-786,833-843): frame f starts at
 * (f / sub)*hop + sub_off[f % sub] (hop = samples per symbol N, sub_off = RTTY_Params.NSTART, rtty.py:389-394);
 * PYSDR_PSD_RAW: |X|^2 without the 1/(navg*sum w^2) normalisation and without the dB floor (rtty.py:841);
 * PYSDR_PSD_FLIP: line = flipud(fftshift(.)) (rtty.py:843).  sub = 1, flags = 0 restores pysdr_psd_lines' default. */
#define PYSDR_PSD_RAW 1
#define PYSDR_PSD_FLIP 2
int pysdr_psd_configure(pysdr_psd *p, int32_t sub, const int32_t *sub_off, int32_t flags);

/* Transform lengths that are not a power of two (or exceed 16384) — the reference's RF panel uses chunk 32818 /
 * NFFT 65636 (Plotting.py:370-375) — go through Bluestein's chirp-z identity on a 2^17-point four-step FFT (czt.cu).
 * Tables come from the host in float64-derived complex64: wc[n] = w[n]*conj(b[n]) (chunk entries), bspec = FFT_M of
 * the wrapped chirp b[m] = exp(j*pi*m^2/nfft) divided by M, laid out [512][256] in the transforms' position order
 * (pysdr_fft_pos_to_freq gives the bin of a position).  Same frame / averaging / dB / fftshift contract as
 * pysdr_psd_lines.  Requires nfft + chunk - 1 <= 131072. */
typedef struct pysdr_czt pysdr_czt;
int pysdr_fft_pos_to_freq(int nfft, int pos);
int pysdr_czt_create(int32_t chunk_size, int32_t nfft, int32_t hop, const float *window_host,
                     const float *wc_host, const float *bspec_host, pysdr_czt **out);
int pysdr_czt_destroy(pysdr_czt *p);
int pysdr_czt_lines(pysdr_czt *p, const void *d_x, int64_t n, int is_complex, int32_t navg, int dB,
                    float *d_out, int64_t *n_lines, void *stream);
int64_t pysdr_czt_launch_count(const pysdr_czt *p);

/* ---- a12: three_box_plot waterfall compute (reference Plotting.py:536-548,583-587,618-626,689-695)
 * d_wf float32[nfft][ncols] state; shift-in one line, optional roll by nbins, background =
 * median(mean(wf[:, -cnt:],1)) -> *d_bkgnd, image = max(wf-bkgnd, max(wf-bkgnd)-pan_dr) -> d_img. */
int pysdr_waterfall_push(float *d_wf, int32_t nfft, int32_t ncols, int32_t cnt, const float *d_line,
                         int32_t npsd, int32_t roll_bins, float pan_dr, float *d_img, float *d_bkgnd,
                         float *d_scratch, void *stream);

/* Peak picker of the panadapter (reference Plotting.py:594: scipy.signal.find_peaks(PSD2, distance=PEAK_DIST/df,
 * height=bkgnd+10)) on the device: strict local maxima (plateaus at their middle), height >= *d_bkgnd + height_above_bkgnd
 * (d_bkgnd = the median pysdr_waterfall_push left on the device; NULL: height >= min_height), then the distance rule from the
 * highest peak down.  d_peaks receives at most 4096 ascending indices, d_count[0] their number.  One small launch, no host
 * round trip between the waterfall update and the peak list. */
int pysdr_find_peaks(const float *d_x, int32_t n, const float *d_bkgnd, float height_above_bkgnd, float min_height, float distance,
                     int32_t *d_peaks, int32_t *d_count, void *stream);
/* Colour mapping of the waterfall image (reference Plotting.py:139-142: 256-entry 'jet' lookup table on the image
 * item): d_rgba[i] = lut[round(255*(img[i]-lo)/(hi-lo))], [lo,hi] = [max-pan_dr, max] as left by
 * pysdr_waterfall_push in d_scratch/d_bkgnd.  d_lut256: 256 x RGBA8, d_rgba: n x RGBA8. */
int pysdr_waterfall_rgba(const float *d_img, int64_t n, const float *d_bkgnd, const float *d_scratch,
                         int32_t nfft, int32_t ncols, float pan_dr, const void *d_lut256, void *d_rgba,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif
