/* pycusdr_b200 -- C ABI of the B200-native demodulator hot path.
 *
 * Drop-in boundary for pyCuSDR's per-chunk matched-filter Doppler search, symbol-timing recovery
 * and symbol decisions.  The reference has no C ABI of its own: its demodulator class reaches the
 * GPU through PyCUDA prepared kernels (pyCuSDR/demodulator/demodulator_base.py:353-390,505-506)
 * and a ctypes binding to cuFFT (pyCuSDR/lib/cufft.py:143,231,266,283,365).  Each entry point below
 * names the reference call sequence it replaces; the Python mirror of the reference class
 * (pycusdr_b200/demodulator) binds them with ctypes, and INTEGRATION.md shows the stub a pyCuSDR
 * maintainer would add.
 *
 * Conventions: plain C types only; every function returns 0 on success or a negative pcs_status;
 * pcs_last_error() returns a thread-local message for the last failure.  One handle owns one CUDA
 * stream, all device memory and one pinned host chunk buffer; calls on a handle must not overlap
 * (the reference makes strictly alternating uploadAndFindCarrier/demodulate calls from one process,
 * pyCuSDR/demodulator_process.py:293-297).  There is no CPU fallback: without a CUDA device
 * pcs_create fails.
 */
#ifndef PYCUSDR_B200_H
#define PYCUSDR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCS_ABI_VERSION 1

typedef enum {
    PCS_OK = 0,
    PCS_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
    PCS_ERR_CUDA = -2,         /* CUDA runtime error (message in pcs_last_error) */
    PCS_ERR_NO_DEVICE = -3,    /* no usable CUDA device */
    PCS_ERR_STATE = -4         /* call order violated (e.g. demod before a chunk was uploaded) */
} pcs_status;

/* Which formulation computes the search metric:
 *   AUTO / OVERLAP_SAVE  the (bins x masks x samples) correlation surface by overlap-save with B-point transforms that
 *                        never leave the SM: 256-point register-resident blocks for filters up to 128 taps, 2^9..2^13
 *                        shared-memory blocks for longer ones (up to 4096 taps); same numbers as Nfft-point inverse
 *                        transforms to fp32 rounding
 *   FULL                 reserved (Nfft-point inverse transforms); pcs_create rejects it in this build
 *   PARSEVAL             labelled variant: energies by Parseval's theorem, no inverse transform, no peak / offsets
 *                        (peak_* = -1); identical Doppler estimate, shift and demodulated bits                     */
typedef enum { PCS_PATH_AUTO = 0, PCS_PATH_OVERLAP_SAVE = 1, PCS_PATH_FULL = 2, PCS_PATH_PARSEVAL = 3 } pcs_path;

typedef struct {
    int32_t abi_version;             /* PCS_ABI_VERSION */
    int32_t device;                  /* conf['GPU'][..]['CUDA']['device']          dem_base:178 */
    int32_t nfft;                    /* 2**blockSize                                dem_base:89,113 */
    int32_t num_dopplers;            /* doppCarrierSteps                            dem_base:130 */
    int32_t element_offset;          /* 1 when a noise row is prepended, else 0     dem_base:150-159 */
    int32_t num_masks;               /* protocol.get_filter()[0]                    dem_base:196 */
    int32_t window_width;            /* bitWindowWidth (WINDOW_WIDTH)               dem_base:116,412 */
    int32_t sum_all_masks;           /* protocol.SUM_ALL_MASKS_PYTHON               dem_base:123-127,416 */
    int32_t code_search_mask_offset; /* CODE_SEARCH_MASK_OFFSET                     dem_base:120,417 */
    int32_t samples_per_sym;         /* samplesPerSym                               dem_base:103 */
    int32_t path;                    /* pcs_path */
    int32_t log2_block;              /* 0 = choose; else force the overlap-save block size 2**log2_block */
    int32_t snr_window;              /* half width of the computeSNR windows (5)    dem_base:620 */
    int32_t reserved[3];             /* tuning knobs, all 0 by default.  [0] bit 0: 1 = never replay the per-chunk sequence as a
                                        CUDA graph, bit 1: 96-register build of the shifted-filter kernel (20 warps per SM), bit 2: factorised bank without shared partial sums; [1] bits 0-7: groups per CTA of the 256-point search kernel (0 = 8; 4, 16), bits
                                        8+: (bin, block) items per CTA of the shifted-filter search kernel (0 = 64); [2] form of the
                                        256-point search: 0 = shifted filters (block spectra shared by all bins), 1 / 2 = rotate the
                                        chunk per bin with the block spectrum in shared memory / registers (comparison variants), 3 = shifted
                                        filters with the long-filter bank never factorised (pcs_factorise_bank; comparison variant) */
} pcs_config;

/* Per-chunk scalar results (filled by pcs_search / pcs_demod / pcs_process). */
typedef struct {
    float best_idx;      /* findDopplerEst res[0]: weighted index of the two best bins  kern:562,590 */
    float metric_db;     /* findDopplerEst res[1]                                       kern:565,592 */
    int32_t low_idx;     /* int(best_idx)                                               dem_base:610 */
    int32_t high_idx;    /* ceil(best_idx)                                              dem_base:611 */
    int32_t shift;       /* dopplerIdxlast = round(interpolated spectrum shift)         dem_base:618 */
    int32_t status;      /* 0 ok; 1 = NaN estimate (caller returns zeros)               dem_base:625-630 */
    float timing[3];     /* findCodeRateAndPhase: index, atan2 phase, |.|^2             kern:306-310 */
    int32_t n_sym;       /* number of symbol decisions = int(Nfft / spSym)              dem_base:999 */
    double sp_sym;       /* Nfft / timing[0]                                            dem_base:735 */
    double code_offset;  /* -phase/pi*spSym/2 (+ spSym-1 if negative)                   dem_base:745-747 */
    float peak_val;      /* max |y|^2 over (bin, mask, timing offset) -- north-star peak */
    int32_t peak_bin, peak_mask, peak_offset;
    int32_t sig_start, sig_len, noise_start, noise_len; /* circular spectrum windows for computeSNR */
    int32_t demod_shift; /* the shift the demod stage used */
    int32_t xchg_timeout;/* pcs_shard_*: bit 0 = a peer's rows or the chunk never arrived, bit 1 = a flag overran */
} pcs_result;

typedef struct pcs_handle pcs_handle;

/* Replaces Demodulator.__init__'s device set-up (dem_base:177-221): context, mask upload
 * (__uploadMaskToGPU :246-263), buffers (:433-498), FFT plans (:275-338,501), Doppler shift upload (:221).
 * shifts: int32[num_dopplers + element_offset] = doppCyperSymNorm; masks: complex64[num_masks][nfft]
 * (interleaved re,im) = conj(FFT(template)) exactly as protocol.get_filter returns them. */
int pcs_create(const pcs_config* cfg, const int32_t* shifts, const float* masks, pcs_handle** out);

/* Replaces Demodulator.__del__ (dem_base:517-533). */
int pcs_destroy(pcs_handle* h);

/* Replaces get_signalBufferHostPointer (dem_base:1055-1060): pinned complex64[nfft] owned by the
 * handle; the caller writes samples in place. */
void* pcs_host_buffer(pcs_handle* h);

/* Replaces uploadToGPU (dem_base:548-558): asynchronous copy of the pinned chunk to HBM and the
 * forward FFT.  Returns without synchronising. */
int pcs_upload(pcs_handle* h);

/* Chunks straight from the caller's own sample memory (extension of dem_base:1055-1060 / demodulator_process.py:287, where
 * the caller copies every block into the one pinned buffer -- 8 * nfft bytes of host memcpy per chunk, more than the H2D
 * copy costs).  pcs_host_register page-locks a range once (cudaHostRegister; e.g. the ring a receiver thread writes, of
 * which consecutive chunks are overlapping windows: sigFIFO.py:147-181); pcs_set_host_source names complex64[nfft] inside
 * such a range as the source of the NEXT pcs_upload / pcs_chunk_to_bits (one shot; NULL cancels).  The range must stay
 * unchanged until the next synchronising call on the handle.  Pageable memory is refused (PCS_ERR_INVALID).
 * pcs_upload_thresholded always uses the handle's own buffer (it clips in place). */
int pcs_host_register(void* ptr, uint64_t bytes);
int pcs_host_unregister(void* ptr);
int pcs_set_host_source(pcs_handle* h, const void* chunk);

/* Replaces __thresholdInput followed by uploadToGPU (STX backend: STX.py:13-20, dem_base:670-707, 548-558): the pinned
 * chunk is copied to HBM, clipped there in two passes to scale * mean(|x|) (scale = peakThresholdScale; the means are
 * float32 pairwise sums in np.mean's order), and copied back into the pinned buffer, which the reference clips in
 * place and whose tail the caller carries into the next chunk (demodulator_process.py:337).  clipped_idx receives the
 * ascending indices of the samples clipped by the second pass (clippedPeakIPure, dem_base:683-684), at most cap of
 * them; *n_clipped is their full count.  thresholds (float[2], may be NULL) = the two clip levels.  Synchronises.
 * nfft <= 2^22. */
int pcs_upload_thresholded(pcs_handle* h, float scale, int64_t* clipped_idx, int32_t cap, int32_t* n_clipped,
                           float* thresholds);

/* clippedPeakI (dem_base:686-705): idx (ascending, unique: what pcs_upload_thresholded returns) with every gap shorter
 * than min_gap (peakMinGap = 100) samples filled in.  Host only.  *n_out is the full count even when it exceeds cap. */
int pcs_fill_gaps(const int64_t* idx, int32_t n, int32_t min_gap, int64_t* out, int32_t cap, int32_t* n_out);

/* Same, but the chunk is already in HBM (device pointer to complex64[nfft]); no copy is made and the
 * buffer must stay valid until the next synchronising call. */
int pcs_upload_device(pcs_handle* h, const void* d_chunk);

/* Doppler-RATE hypothesis.  The reference prepares complexHeterodyne (cuda_kernels.cu:755-778,
 * dem_base:388: out[x] = in[x] * exp(j theta), theta = fmod(((a x) + b) x, 2 pi) + c, fp32) and never calls it; this is the
 * same statement applied to the uploaded chunk in the time domain.  After the call every pcs_search / pcs_demod /
 * pcs_process works on the de-chirped chunk, until the next upload or pcs_heterodyne (each call starts again from the chunk
 * as uploaded; a = b = c = 0 restores it).  A rate search is a loop of pcs_heterodyne + pcs_search over the hypotheses
 * (a = -pi * rate / fs^2) followed by pcs_heterodyne(best) + pcs_process: `Demodulator.findUHFRates`. */
int pcs_heterodyne(pcs_handle* h, float a, float b, float c);
int pcs_get_chunk(pcs_handle* h, float* x_out /* complex64[nfft]: the chunk the search currently works on */);

/* Replaces __findUHF's device part (dem_base:571-605): energy surface reduction, findDopplerEst,
 * shift interpolation (:610-618).  Synchronises.  E_out (float32[(D+off)*M], reference layout) and
 * res may be NULL. */
int pcs_search(pcs_handle* h, pcs_result* res, float* E_out);

/* Replaces __demodulate's device part (dem_base:776-803): surface at the selected shift, timing
 * recovery, symbol decisions.  shift < 0 uses the shift found by the last search (kept on the
 * device).  Synchronises.  sym/centre: int32[max_sym], mag: float32[max_sym]; the first res->n_sym
 * entries are valid.  Any output pointer may be NULL. */
int pcs_demod(pcs_handle* h, int32_t shift, pcs_result* res, int32_t* sym, int32_t* centre, float* mag);

/* pcs_search followed by pcs_demod(shift = found) with a single synchronisation and no host round
 * trip in between.  */
int pcs_process(pcs_handle* h, pcs_result* res, float* E_out, int32_t* sym, int32_t* centre, float* mag);

/* Enqueue one whole chunk (search + demod) on a chunk already in HBM without synchronising; results
 * are fetched later with pcs_fetch().  Used to keep several chunks in flight. */
int pcs_enqueue_device(pcs_handle* h, const void* d_chunk);
int pcs_fetch(pcs_handle* h, pcs_result* res, float* E_out, int32_t* sym, int32_t* centre, float* mag);

/* Upper bound of n_sym (= nfft / (samples_per_sym / 2), dem_base:468). */
int32_t pcs_max_symbols(const pcs_handle* h);

/* The two spectrum windows computeSNR averages (dem_base:653-661), gathered on the device:
 * complex64[res.sig_len] starting at bin res.sig_start (circular) and the same for the noise window. */
int pcs_snr_windows(pcs_handle* h, float* sig_win, float* noise_win);

/* Inspection / parity hooks (synchronise; not on the hot path). */
int pcs_get_spectrum(pcs_handle* h, float* X_out /* complex64[nfft] */);
int pcs_get_peaks(pcs_handle* h, float* peak_val /* [D*M] */, int32_t* peak_offset /* [D*M] */);
int pcs_get_demod_surface(pcs_handle* h, int32_t shift, float* y_out /* complex64[M*nfft] */);
int pcs_get_demod_magnitudes(pcs_handle* h, float* ymag_out /* float32[M*nfft] */, float* p_out /* float32[nfft] */);

/* Plan introspection: which path was chosen and its geometry. */
typedef struct {
    int32_t path;           /* pcs_path actually used */
    int32_t log2_block;     /* overlap-save block size */
    int32_t valid_per_block;
    int32_t num_blocks;
    int32_t support_pos;    /* filter time support: taps at n = 0..support_pos */
    int32_t support_neg;    /* and at n = -support_neg..-1 */
    int32_t groups_per_cta;
    int32_t search_ctas;
    int32_t search_smem_bytes;
    int32_t sm_count;
    int64_t device_bytes;   /* HBM allocated by the handle */
} pcs_plan_info;
int pcs_get_plan(const pcs_handle* h, pcs_plan_info* info);

/* Segment factorisation of the matched-filter bank (plan-time helper of pcs_create for filters too long for the 256-point
 * kernel; host only, no GPU -- exported so that the tables can be checked without one).  The reference's FSK-2 bank
 * (pyCuSDR/protocol/FSK2_base.py:17-46: 2^k templates, each k one-symbol tone segments with a continuous phase; CC11xx:
 * 8 x 3 x 128 taps) has only R = 2 distinct segments up to a complex constant.  With the taps g_m[n] = IFFT(Mk[m]) on
 * n = -support_neg .. support_pos cut into J segments of S taps counted from the support_pos end,
 *     g_m[n] = sum_j c[m][j] * b_{sel[m][j]}[n + j S],     b_r supported on n = support_pos - S + 1 .. support_pos,
 * every filter output is y_m[i] = sum_j c[m][j] * u_{sel[m][j]}[i + j S] with u_r = b_r * x, so the search needs R inverse
 * transforms per (bin, block) instead of M (a6-a8, kern:339-373, 421-480: same numbers to fp32 rounding).  The structure is
 * detected from `masks` (the spectra given to pcs_create), never assumed; *num_basis = 0 means none was found, or it would
 * not pay, or it failed the acceptance test (the factorised bank must reproduce Mk[m][(k N/B - shift) % N] to 1e-5 of
 * the peak).  Outputs, sized by the caller for the maxima: sel_out int32[M * PCS_FB_MAX_SEG] (used [M][J]), coef_out
 * complex64[D * M * PCS_FB_MAX_SEG] (used [D][M][J]; includes the bin's exp(-2 pi i shift j S / N)), basis_spec_out
 * complex64[D * PCS_FB_MAX_BASIS * B] (used [D][R][B], B = 2**log2_block: the B-point spectra of the bin's shifted basis
 * filters, scaled like the kernel's filter spectra); each may be NULL. */
#define PCS_FB_MAX_SEG 4
#define PCS_FB_MAX_BASIS 4
int pcs_factorise_bank(const float* masks /* complex64[M*nfft] */, int32_t nfft, int32_t num_masks, int32_t support_pos,
                       int32_t support_neg, const int32_t* shifts, int32_t num_shifts, int32_t log2_block, int32_t* seg_len,
                       int32_t* num_seg, int32_t* num_basis, int32_t* sel_out, float* coef_out, float* basis_spec_out);
/* Which form of the factorised search a bank can take (host only, no GPU; the second plan-time step of pcs_create after
 * pcs_factorise_bank).  *form = 1: general (selectors read at run time; sel / coef untouched).  2: a COMPLETE BINARY BANK --
 * num_basis = 2 and the num_masks = 2^num_seg selector rows are all different (what an FSK-2 bank is): coef is rewritten in
 * code order, code = sum_j sel[m][j] 2^j, and sel[0 .. M) becomes the table code -> mask, so that every selection in the
 * kernel is static.  3: as 2, and (only looked for when allow_shared_sums != 0) the coefficient of segment j is the same
 * for all codes with the same low j + 1 bits -- for FSK-2 the phase a symbol starts with is set by the symbols before it --
 * so filters with the same prefix share their partial sums (2 + 4 + .. + 2^J complex multiply-adds per output instead of
 * J 2^J).  sel: int32[M * J] in, int32[>= M] out; coef: complex64[D * M * J] in / out. */
int pcs_bank_code_order(int32_t num_masks, int32_t num_seg, int32_t num_basis, int32_t num_shifts, int32_t* sel, float* coef,
                        int32_t allow_shared_sums, int32_t* form);
/* What the handle's search uses: out[0] = 0 unfactorised, 1 factorised (general form: selectors read at run time), 2 = a
 * complete binary bank (R = 2, the M = 2^J selector rows all different: static selection, combinations from registers), 3 = and
 * the coefficient of segment j depends on the selectors of segments 0..j only (partial sums shared between filters with
 * the same prefix; an FSK-2 bank); out[1..3] = S, J, R. */
int pcs_get_bank_factor(const pcs_handle* h, int32_t out[4]);

/* Kernel launch counter (all launches issued through this handle since creation). */
int64_t pcs_launch_count(const pcs_handle* h);

/* CUDA stream the handle enqueues on (cudaStream_t as an integer), for event timing by the caller. */
uint64_t pcs_stream(const pcs_handle* h);

/* Doppler-bin sharding across GPUs (one process and one handle per GPU, every handle created with the
 * FULL shift table).  Rank r restricts its search to rows [lo, hi) with pcs_set_bin_range, enqueues
 * pcs_enqueue_search_local (search kernel + partial reduction, no estimate), all-gathers the row slices of
 * the three tables returned by pcs_shard_buffers (float32[D*M] energies, float32[D*M] peak values,
 * int32[D*M] peak offsets -- device pointers) with NCCL on the handle's stream, and then every rank runs
 * pcs_enqueue_estimate_and_demod, which reproduces the single-GPU estimate bit for bit because it scans the
 * same full table.  Results are collected with pcs_fetch. */
int pcs_set_bin_range(pcs_handle* h, int32_t lo, int32_t hi);
int pcs_shard_buffers(pcs_handle* h, void** d_energy, void** d_peak_val, void** d_peak_off);
int pcs_enqueue_search_local(pcs_handle* h);
int pcs_enqueue_estimate_and_demod(pcs_handle* h, int32_t with_demod);

/* Bin-sharded STREAMING search without a collective on the data path (replaces, for N GPUs, the chunk loop
 * of pyCuSDR/demodulator_process.py:284-338 around uploadAndFindCarrier / demodulate, and sigFIFO.py:147-181 as the one
 * source of samples).  One process and one handle per GPU, every handle created with the FULL shift table.
 *   - rank 0 is the ingest rank: it alone receives samples (PCS_SRC_HOST: a host pointer, or NULL after writing into the
 *     pinned slot pcs_shard_host_slot returned; PCS_SRC_DEVICE: a device pointer).  The chunk is copied once into rank 0's
 *     HBM and from there into every peer's chunk ring over NVLink by the copy engines; a data flag tells the peer's stream.
 *   - every rank searches its slice of the Doppler bins; the finishing CTAs of its search kernel store the slice's rows of
 *     the three [D][M] tables straight into the exchange region of the chunk's OWNER (seq % world) and raise a row flag.
 *   - the owner alone runs estimate + demodulation + timing + symbol decisions on the gathered tables (the very tables one
 *     GPU produces: results are bit-identical to pcs_process) on one of four tail streams -- from an owned chunk's second
 *     pass through a stage on as ONE CUDA graph launch -- and keeps the results in one of eight result stages until
 *     pcs_shard_fetch collects them.
 * pcs_shard_init allocates the exchange region (tables, flags, chunk ring of `ring` slots: even, <= 8 * world, 0 = choose: 8)
 * and returns its 64-byte CUDA IPC handle; the caller all-gathers the handles with any transport and passes them, in rank
 * order, to pcs_shard_attach.  pcs_shard_submit(seq) is then called on EVERY rank for seq = 0, 1, 2, ... (src is ignored
 * on ranks other than 0); it never blocks on another rank.  The owner must fetch chunk seq before it submits chunk
 * seq + 8 * world.  A handle initialised for sharding is dedicated to it.  world = 1 is allowed (a single GPU streaming
 * through the same engine). */
enum { PCS_SRC_DEVICE = 1, PCS_SRC_HOST = 2 };
int pcs_shard_init(pcs_handle* h, int32_t rank, int32_t world, int32_t ring, void* ipc_handle_out /* 64 bytes */);
int pcs_shard_attach(pcs_handle* h, const void* ipc_handles /* world x 64 bytes, rank order */);
int pcs_shard_info(const pcs_handle* h, int32_t* ring, int32_t* lanes, int32_t* bin_lo, int32_t* bin_hi, int32_t* stages);
int pcs_shard_host_slot(pcs_handle* h, int64_t seq, void** out /* pinned complex64[nfft] for chunk seq (rank 0) */);
int pcs_shard_submit(pcs_handle* h, int64_t seq, int32_t src_kind, const void* src);
/* Owner only; blocks until the tail of chunk seq has finished.  res->xchg_timeout != 0: a peer's rows or the chunk never
 * arrived (bit 0) or a flag overran (bit 1) -- the results are then not valid.  sig_mean / noise_mean / snr_ok as in
 * pcs_snr_means.  Any output pointer may be NULL. */
int pcs_shard_fetch(pcs_handle* h, int64_t seq, pcs_result* res, float* E_out, int32_t* sym, int32_t* centre, float* mag,
                    float* sig_mean, float* noise_mean, int32_t* snr_ok);
int pcs_shard_sync(pcs_handle* h);
/* Diagnostic timeline (environment PCS_SHARD_TRACE=1 at pcs_shard_init): for the last <= 64 chunks, milliseconds since init of
 * {block spectra start, search end, tail end}; out = float[3 * 64]. */
int pcs_shard_trace(pcs_handle* h, int64_t* first, int32_t* count, float* out);
int pcs_shard_streams(const pcs_handle* h, uint64_t* out4 /* lane 0, lane 1, copy, tail (cudaStream_t as integers) */);

/* Make the handle enqueue on a caller-owned stream (cudaStream_t as an integer), e.g. the framework
 * stream NCCL collectives are ordered on.  The handle's own stream is destroyed. */
int pcs_set_stream(pcs_handle* h, uint64_t stream);

/* Per-stage device timing with CUDA events on the handle's stream (bench / roofline reporting).
 * Stages: 0 chunk spectrum, 1 search kernel, 2 estimate, 3 demod surface, 4 timing + symbols,
 * 5 partial reduction (search forms without the fused finish), 6 block spectra of the chunk.
 * pcs_get_profile fills double[7] accumulated milliseconds and int64[7] counts. */
#define PCS_NUM_STAGES 7
int pcs_set_profiling(pcs_handle* h, int enable);
int pcs_get_profile(pcs_handle* h, double* stage_ms, int64_t* stage_count);

/* Measured fp32 FMA throughput of a device in TFLOP/s (the roofline denominator of the FFT-bound
 * kernels; 16 independent FMA chains per thread, best of 4 timed launches). */
int pcs_measure_fp32_peak(int device, double* tflops);

/* Host-side bit post-processing of a chunk in C++ (no device work): replaces the NumPy code the reference runs after
 * cudaFindCentres -- extractBits / extractBitsNRZs (dem_base:1012-1051), checkSymbolOverlap (dem_base:863-988), the
 * clipped-peak tagging of the trust (dem_base:817-837) and the uint8 casts of the return statement (dem_base:859).
 * The stitcher keeps the cross-chunk state (poswinP / posSymEnd, dem_base:977-979).  Exactly one of bit_lut
 * (uint8[num_symbols], protocol.get_symbolLUT2()[0]) and symbol_lut (int32[num_symbols][2][lut_k], the NRZ-S table)
 * is given.  Outputs hold at most n_sym entries; *n_out are valid. */
typedef struct pcs_stitcher pcs_stitcher;
typedef struct {
    int32_t nfft;             /* 2**blockSize                                   dem_base:89 */
    int32_t overlap;          /* 2**overlap samples                             dem_base:90 */
    int32_t overlap_offset;   /* symbol_check_overlap_offset                    dem_base:97 */
    int32_t error_threshold;  /* symbol_check_error_threshold                   dem_base:98 */
    int32_t match_threshold;  /* overlap_offset - num_errors_allowed            dem_base:99 */
    int32_t num_symbols;      /* rows of the look-up table (= number of masks) */
    int32_t lut_k;            /* successors per row of symbol_lut, 0 with bit_lut */
    int32_t reserved;
} pcs_stitch_config;
int pcs_stitch_create(const pcs_stitch_config* cfg, const uint8_t* bit_lut, const int32_t* symbol_lut, pcs_stitcher** out);
int pcs_stitch_chunk(pcs_stitcher* s, const int32_t* sym, const int32_t* centre, const float* mag, int32_t n_sym,
                     const int64_t* clipped, int32_t n_clipped, double sp_sym, uint8_t* bits_out, uint8_t* centres_out,
                     uint8_t* trust_out, int32_t* n_out);
int pcs_stitch_reset(pcs_stitcher* s);
int pcs_stitch_destroy(pcs_stitcher* s);
/* The carry between consecutive chunks (poswinP, posSymEnd of dem_base:977-979) as bytes: n_poswin bits followed by
 * n_posend bits.  It depends on its own chunk's symbols only, so a host that post-processes consecutive chunks in
 * different processes (bin sharding: the owner of a chunk rotates over the ranks) passes it from the owner of chunk k to
 * the owner of chunk k + 1 and gets exactly the bit stream a single process produces. */
int pcs_stitch_get_state(const pcs_stitcher* s, uint8_t* buf, int32_t cap, int32_t* n_poswin, int32_t* n_posend);
int pcs_stitch_set_state(pcs_stitcher* s, const uint8_t* buf, int32_t n_poswin, int32_t n_posend);

/* computeSNR's two window means (dem_base:657-663): mean |X| over the signal and the noise window gathered by the last
 * search, when the windows do not touch the ends of the spectrum (*ok = 1); otherwise *ok = 0 and the caller applies the
 * reference's slicing rules to pcs_snr_windows / pcs_get_spectrum itself.  shifts = the table given to pcs_create. */
int pcs_snr_means(pcs_handle* h, const int32_t* shifts, float* sig_mean, float* noise_mean, int32_t* ok);

/* mean(|z|) of complex64[n] as float32 (np.mean(np.abs(z)) of dem_base:657-661; host only, no GPU): the one
 * implementation every schedule uses for the two computeSNR window means. */
int pcs_mean_abs_c64(const float* z, int32_t n, float* out);

/* Whole chunk in one call for a host that wants bits (UHF backend): pcs_upload + pcs_process + pcs_snr_means +
 * pcs_stitch_chunk; the symbol tables go from the result staging area to the stitcher inside the library.  E_out, sym,
 * centre, mag (inspection copies) and the three SNR outputs may be NULL.  When the device part succeeded and only the
 * stitcher failed (the reference raises from demodulate() in that case, dem_base:868-869), the stitcher's status is
 * returned with *n_out = -1 and res / E_out / sym / centre / mag are valid. */
int pcs_chunk_to_bits(pcs_handle* h, pcs_stitcher* st, const int32_t* shifts, const int64_t* clipped, int32_t n_clipped,
                      pcs_result* res, float* E_out, int32_t* sym, int32_t* centre, float* mag, float* sig_mean,
                      float* noise_mean, int32_t* snr_ok, uint8_t* bits_out, uint8_t* centres_out, uint8_t* trust_out,
                      int32_t* n_out);

/* Decoder-side frame sync search on the bit stream this library hands over (decoder.py:96-104):
 * score = np.convolve(bits, mask) with mask = protocol.get_mask() (+-1 header, flipped); candidates are the positions
 * with score >= threshold (= numOnesHeader - headerTol).  idx_out / score_out receive at most `cap` candidates in
 * increasing order (idx = position in the full convolution; packet start = idx - m + 1); *n_found counts all of them. */
int pcs_sync_search(const uint8_t* bits, int64_t n, const int8_t* mask, int32_t m, int32_t threshold, int32_t* idx_out,
                    int32_t* score_out, int32_t cap, int32_t* n_found);

/* Soft-combiner bit-stream alignment (softCombiner.py:697-722 with lib/customXCorr.py:5-30), host only, exact integers:
 * out[k] = sum_j a_pad[(j + k) mod n] * b_pad[j] for k in [0, n): the circular cross-correlation of two 0/1 streams zero
 * padded to n (the reference: a = the slave's bits, n = 2**ceil(log2(len(a))), b = master[:len(a)]; its
 * np.abs(customXCorr(...)) equals these integers up to FFT rounding).  n >= na, nb.  pcs_topk_i32 is the reference's
 * "15 largest by repeated arg-max" (:708-715), first index on ties. */
int pcs_bit_xcorr(const uint8_t* a, int64_t na, const uint8_t* b, int64_t nb, int64_t n, int32_t* out /* int32[n] */);
int pcs_topk_i32(const int32_t* v, int64_t n, int32_t k, int64_t* idx_out, int32_t* val_out);

/* Native sample ingest: what the reference's SigFIFO ring buffer (sigFIFO.py:13-181) and the chunk loop of
 * demodulator_process.py:284-338 do on the host, as a pipeline.  Samples are pushed in arbitrary block sizes; every
 * nfft - overlap new samples become a chunk whose first `overlap` samples are carried over from the previous chunk ON
 * THE DEVICE; the H2D copy of a chunk overlaps the kernels of the chunks before it; chunks go round-robin to the
 * `n_handles` handles (all created with the same configuration), results are popped in chunk order.  Per-chunk results
 * are identical to pcs_upload + pcs_process on the same samples.  pcs_ingest_push may block while the oldest chunk of a
 * handle is still running (back-pressure); pcs_ingest_pop(block = 0) never blocks.  sig_win / noise_win receive the
 * computeSNR windows (complex64[res->sig_len] each, see pcs_snr_windows). */
typedef struct pcs_ingest pcs_ingest;
int pcs_ingest_create(pcs_handle* const* handles, int32_t n_handles, int32_t nfft, int32_t overlap, int32_t num_bins,
                      int32_t num_masks, int32_t device, pcs_ingest** out);
int pcs_ingest_push(pcs_ingest* s, const void* samples /* complex64[n] */, int64_t n, int32_t* chunks_submitted);
int pcs_ingest_pop(pcs_ingest* s, int32_t block, pcs_result* res, float* E_out, int32_t* sym, int32_t* centre, float* mag,
                   float* sig_win, float* noise_win, int32_t* ready);
int pcs_ingest_pending(const pcs_ingest* s, int64_t* submitted, int64_t* popped);
int pcs_ingest_destroy(pcs_ingest* s);

const char* pcs_last_error(void);
int pcs_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PYCUSDR_B200_H */
