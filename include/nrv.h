/*
 * nrv.h -- C-ABI of libnrv.so: the B200 (sm_100a) implementation of NanoReviser's
 * revision-inference hot path.
 *
 * The reference (pkubioinformatics/nanoreviser) has no FFI of its own: the boundary is the
 * body of provide_fasta() (NanoReviser.py:105-183) and the Python functions it composes.
 * Each entry point below names the reference interface it replaces.  Plain pointers and
 * sizes only; no C++ or torch types.  All functions return 0 on success, a negative
 * NRV_E_* code otherwise; nrv_last_error() gives the message.  There is no CPU fallback:
 * without a CUDA device nrv_create() fails.
 *
 * Ragged batches are CSR: per-read offsets into flat arrays.
 *   signal   int16  [sig_off[R]]   raw signal AFTER abs_event_start, i.e. or_raw_signal[a0:]
 *                                  (NanoReviser.py:120) of every read, concatenated
 *   starts   int32  [base_off[R]]  base starts relative to the read's signal[0]
 *                                  (get_read_data's `start - abs_event_start`,
 *                                   nanorev_fast5_handeler.py:146)
 *   bases    uint8  [base_off[R]]  ASCII event bases (nanorev_fast5_handeler.py:98-114)
 *   ev_mean, ev_std float [base_off[R]]   event mean / stdv per base (:100-113)
 *   last_dur int32  [R]            int(event_length[-1]) (NanoReviser.py:123), 3 or 5
 * Per-base durations are not passed: length = diff(start) (+ last_dur), as :121-126.
 */
#ifndef NRV_H
#define NRV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRV_OK               0
#define NRV_E_INVALID       -1   /* bad argument */
#define NRV_E_CUDA          -2   /* CUDA runtime error (sticky for the handle) */
#define NRV_E_CAPACITY      -3   /* caller-provided output buffer too small */
#define NRV_E_NODEVICE      -4   /* no CUDA device / not sm_100 */

/* per-read status (nrv_result.status) -- mirrors provide_fasta's per-stage try/except
 * (NanoReviser.py:114-132,146-154): any non-zero status means "write the original bases". */
#define NRV_READ_OK          0
#define NRV_READ_TOO_SHORT   1   /* N <= W: no window; passes through unchanged */
#define NRV_READ_SCALE_ZERO  2   /* MAD == 0: reference would divide by zero (preprocessing.py:119) */
#define NRV_READ_BAD_EVENTS  3   /* starts not increasing / beyond the signal
                                    (nanorev_fast5_handeler.py:142-143, preprocessing.py:168-169) */

#define NRV_SIGNAL_LEN 50        /* lstmmodel.py:28 SIGNEL_LEN, preprocessing.py:85 query_len */
#define NRV_VEC_LEN     6        /* lstmmodel.py:29 */

/* One direction of a Keras-2.2.4 LSTM: gate order i,f,c,o along the 4u axis. */
typedef struct {
    const float* kernel;      /* [in][4u] */
    const float* recurrent;   /* [u][4u]  */
    const float* bias;        /* [4u]     */
} nrv_lstm_dir;

/* Weights of one predict model, in the layout Keras saved them
 * (lstmmodel.py:32-81 model1, :84-133 model2; nanorevcnn.py:17-38).  Host pointers;
 * nrv_create copies (and re-packs) them to the device. */
typedef struct {
    int32_t window;           /* W = feature.kernel.shape[0] / 6 (11 for the shipped files) */
    int32_t n_class;          /* 6 (model1) or 5 (model2) */
    const float* conv1_k;     /* [3][1][8] */
    const float* conv1_b;     /* [8] */
    const float* bn1;         /* [4][8]  gamma, beta, moving_mean, moving_variance */
    const float* conv2_k;     /* [3][8][8] */
    const float* conv2_b;     /* [8] */
    const float* bn2;         /* [4][8] */
    const float* sig_dense_k; /* [400][64]  TimeDistributed(Dense(64)) on Flatten([50][8]) */
    const float* sig_dense_b; /* [64] */
    nrv_lstm_dir lstm[4][2];  /* read_rnn1(6->16), read_rnn11(32->64), total_rnn1(192->128),
                                 total_rnn2(256->64); [layer][0=forward,1=backward] */
    const float* bn_rnn[3];   /* [4][32], [4][128], [4][256] after lstm 0,1,2 */
    const float* dense1_k;    /* [128][128] */
    const float* dense1_b;
    const float* dense2_k;    /* [128][32] */
    const float* dense2_b;
    const float* main_k;      /* [32][6] */
    const float* main_b;
    const float* feat_k;      /* [W*6][16] */
    const float* feat_b;
    const float* final_k;     /* [16][n_class] */
    const float* final_b;
} nrv_model_weights;

typedef struct {
    int64_t n_reads;
    const int16_t* signal;
    const int64_t* sig_off;   /* [n_reads+1], ALWAYS host memory */
    const int32_t* starts;
    const int64_t* base_off;  /* [n_reads+1], ALWAYS host memory */
    const uint8_t* bases;
    const float*   ev_mean;
    const float*   ev_std;
    const int32_t* last_dur;  /* [n_reads] */
    const uint8_t* qual;      /* optional (may be NULL), [total_bases]: the basecaller's Phred score (0..93, i.e. the Fastq
                                 quality character minus 33) of every base; only read when nrv_result.revised_qual is set:
                                 bases that pass through unrevised keep it (NULL: Phred 40).  The reference copies the
                                 basecaller's qualities in its per-read fallback (NanoReviser.py:172-181) */
} nrv_batch;

/* Outputs.  Every pointer except revised/out_off/status is optional (NULL = not wanted).
 * n_windows = sum over reads of max(N_r - W, 0); windows are ordered read by read. */
typedef struct {
    uint8_t* revised;         /* concatenated revised sequences (ASCII) */
    int64_t  revised_cap;     /* capacity in bytes; 2*total_bases + n_reads always suffices */
    int64_t* out_off;         /* [n_reads+1] offsets into revised */
    int32_t* status;          /* [n_reads] NRV_READ_* */
    uint8_t* y1;              /* [n_windows] argmax of model1 (label space 0..5) */
    uint8_t* y2;              /* [n_windows] argmax of model2 (class 0..4 == label-1) */
    float*   p1;              /* [n_windows][6] softmax of model1 */
    float*   p2;              /* [n_windows][5] softmax of model2 */
    uint8_t* revised_qual;    /* optional (may be NULL), [revised_cap]: Phred+33 quality character of every byte of `revised`
                                 (-F fastq, NanoReviser.py:158-181 / prep_read_fastq output_handeler.py:48-62).  The reference
                                 defines no qualities for the two-model path; definition D6' (DESIGN.md): symbols emitted because
                                 of the models carry min over both models of floor(-10 log10(1 - p_argmax)) capped at 60, the
                                 leading label symbol that of model1, bases that pass through keep nrv_batch.qual */
} nrv_result;

typedef struct nrv_handle nrv_handle;

/* Replaces get_model1()/get_model2() + load_weights (NanoReviser.py:129-130, lstmmodel.py:32,84).
 * One handle per GPU, driven by one host thread. */
int nrv_create(int device, const nrv_model_weights* model1, const nrv_model_weights* model2,
               nrv_handle** out);
void nrv_destroy(nrv_handle* h);
const char* nrv_last_error(const nrv_handle* h);   /* h may be NULL: last create error */
const char* nrv_version(void);

/* Number of CUDA kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t nrv_launch_count(const nrv_handle* h);
/* Per-stage device time.  nrv_set_stage_timing(h, 1) resets the totals and makes every later call
 * record CUDA-event pairs on the handle's stream around each stage's launches (no synchronisation on
 * the hot path); nrv_get_stage_ms() synchronises once and returns the totals in ms since the reset,
 * nrv_get_stage_launches() the kernel launches of each stage over the same period.  Stages are named
 * by nrv_stage_name(0 .. nrv_stage_count()-1): read_stats, base_features, cnn, lstm0, proj1, rec1,
 * proj2, rec2, proj3, rec3, heads_gemm, heads, decode (projN / recN = input-projection GEMM and
 * recurrence of Bi-LSTM layer N; on the fp32 SIMT path the fused layer kernels report under recN).
 * `n` is the capacity of `out` and must be >= nrv_stage_count(). */
int nrv_stage_count(void);
const char* nrv_stage_name(int i);
int nrv_set_stage_timing(nrv_handle* h, int enable);
int nrv_get_stage_ms(nrv_handle* h, float* out, int n);
int nrv_get_stage_launches(const nrv_handle* h, int64_t* out, int n);

/* The CUDA stream all work of this handle is enqueued on (cudaStream_t as void*). */
void* nrv_stream(const nrv_handle* h);
int nrv_synchronize(nrv_handle* h);

/* A2 + A3: replaces signal_segmentation() (preprocessing.py:85-170) and the caller-side
 * scaling / feature columns (NanoReviser.py:124-125, nanorevtrainutils.py:159-169).
 * Host buffers.  Any output may be NULL.
 *   shift, scale  double [R]        seg_mean, seg_std  double [total_bases] (raw, un-normalised)
 *   x             float  [total_bases][6]   feature columns as the model sees them (fp32)
 *   sig_win       float  [total_bases][50]  normalised, symmetric zero-padded windows (fp32)  */
int nrv_segment(nrv_handle* h, const nrv_batch* b, double* shift, double* scale,
                double* seg_mean, double* seg_std, float* x, float* sig_win, int32_t* status);

/* A5-A8: replaces model1.predict([S[..., None], X]) and model2.predict(...) (Keras) on explicit
 * windows.  S [n][W][50], X [n][W][6] host fp32; p1 [n][6], p2 [n][5] host fp32 (either may be NULL). */
int nrv_predict_windows(nrv_handle* h, int64_t n, const float* S, const float* X, float* p1, float* p2);

/* A9 + A10: replaces get_base_1(event_bases[5:5+M], y1, y2+2) (output_handeler.py:104-122) applied
 * per read with the pass-through edges of composition D4.  y1/y2 are per-window labels in window
 * order (n_windows entries), status as produced by nrv_segment / nrv_revise_batch (NULL = all ok). */
int nrv_decode(nrv_handle* h, int64_t n_reads, const int64_t* base_off, const uint8_t* bases,
               const uint8_t* y1, const uint8_t* y2, const int32_t* status,
               uint8_t* revised, int64_t revised_cap, int64_t* out_off);

/* The whole path (A2..A10) for one ragged batch; replaces the body of provide_fasta between
 * get_read_data and prep_read_fasta.  Host buffers in, host buffers out; H2D and D2H copies
 * are inside; returns after the results are in the caller's buffers.
 * Numerics: softmax outputs within 1e-3 of the float64 evaluation of the graph (lstmmodel.py:32-133).  Windows whose two
 * best classes are closer than that tolerance are evaluated a second time in the library's most precise tensor-core mode
 * before the argmax is taken, so that the labels -- and with them the revised sequence -- do not depend on the reduced-
 * precision correction passes of the fast path (NRV_REFINE=0 switches this off). */
int nrv_revise_batch(nrv_handle* h, const nrv_batch* b, nrv_result* r);

/* The same call split in two so that a host thread can keep TWO batches in flight (the reference gets its concurrency from a
 * multiprocessing.Pool of provide_fasta workers, NanoReviser.py:203-223; here one host thread per GPU pipelines instead):
 * nrv_submit_batch stages the offsets, enqueues the H2D copies on a copy stream and K1..K4 on nrv_stream(h), and returns a
 * ticket without waiting; nrv_wait_batch(ticket) copies the results out (D2H on the copy stream) and returns when they are in
 * the buffers of the nrv_result given at submit time.  Submitting batch i+1 before waiting for batch i puts the H2D of i+1
 * and the D2H of i under the kernels of the other batch.  At most 2 tickets may be outstanding (a third submit fails with
 * NRV_E_INVALID); the batch's and the result's host arrays must stay valid until nrv_wait_batch returns (use pinned memory for
 * truly asynchronous copies).  nrv_revise_batch == submit + wait. */
int nrv_submit_batch(nrv_handle* h, const nrv_batch* b, nrv_result* r, int64_t* ticket);
int nrv_wait_batch(nrv_handle* h, int64_t ticket);

/* Same as nrv_revise_batch, but the bulk arrays of `b` and the outputs of `r` are DEVICE pointers (sig_off/base_off stay
 * host).  Enqueues on nrv_stream(h) and returns without synchronising: the offsets go through a ring of pinned staging
 * slots guarded by events, so consecutive calls never drain the stream. */
int nrv_revise_batch_device(nrv_handle* h, const nrv_batch* b, nrv_result* r);

/* Diagnostic entry point (tests only): C[M][N] = A[M][K] . Bt[N][K]^T (+ bias[N]) through the split-fp16
 * tcgen05 projection GEMM that the Bi-LSTM input projections use (csrc/nrv_gemm.cu).  Host fp32 buffers;
 * N must be a multiple of 256 and K a multiple of 64. */
int nrv_debug_gemm(nrv_handle* h, int64_t M, int N, int K, const float* A, const float* Bt, const float* bias, float* C);

/* ---- A1 at GPU rate: multi-threaded native fast5 ingest ----------------------------
 * Replaces get_read_data(fast5_fn, basecall_group, basecall_subgroup) (nanorev_fast5_handeler.py:39-150) for a LIST of
 * fast5 files: own HDF5-subset reader, chunk decoding (deflate, or ONT's VBZ filter 32020 = zig-zag delta + streamvbyte +
 * zstd, the plugin the reference bundles under nanorevutils/utils/lib) and the event collapse (:84-118) on n_threads host
 * threads (0 = all cores), packing the reads that succeed straight into the CSR batch above (signal = raw[a0:],
 * NanoReviser.py:120).  Legacy Albacore <= 0.0 tables (float64 `start` seconds rescaled with the raw read's start_time,
 * :65-73) are decoded as the reference does.  A multi-read container (/read_<id>/{Raw/Signal, Analyses/...}; the reference
 * itself only opens single-read files, :132-133) yields one read per member, see nrv_ingest_read_names.
 * Host-only (no CUDA call).  Per-file status mirrors the exceptions of the reference; NRV_INGEST_UNSUPPORTED marks inputs
 * outside the native subset (other filters, float32 / integer legacy columns, big-endian types ...), which the caller must
 * route through the Python reader (nanoreviser_b200/fast5.py) -- nothing is guessed. */
#define NRV_INGEST_OK            0
#define NRV_INGEST_OPEN_FAILED   1   /* "Error opening file. Likely a corrupted file." (:59-61) */
#define NRV_INGEST_NO_EVENTS     2   /* "No events or corrupted events in file" (:76-77) */
#define NRV_INGEST_TOO_SHORT     3   /* "Events is too short or there are too much zero moves" (:127-128) */
#define NRV_INGEST_NO_SIGNAL     4   /* "No signal stored in the file" (:134-135) */
#define NRV_INGEST_SIGNAL_SHORT  5   /* "Signal is shorter than the Events" (:142-143) */
#define NRV_INGEST_CORRUPT       6   /* structurally invalid HDF5 inside the supported subset */
#define NRV_INGEST_UNSUPPORTED   7   /* valid but outside the native subset: use the Python reader */

typedef struct nrv_ingest nrv_ingest;   /* owns the packed host arrays */

int nrv_ingest_fast5(const char* const* paths, int64_t n_files, const char* basecall_group, const char* basecall_subgroup,
                     int n_threads, nrv_ingest** out);
/* Views into the result, valid until nrv_ingest_free: the batch of the reads that succeeded (in file order),
 * file_status[n_files], read_file[n_reads] (index into paths of every packed read), a0[n_reads] (abs_event_start).
 * batch->qual holds the basecaller's Phred scores of the event-collapsed bases (the Fastq dataset next to Events: the call is
 * Fastq_seq[2:-2]; extract_fastq, nanorev_fast5_handeler.py:152-171, reads the same dataset) when EVERY packed read has a
 * Fastq dataset that lines up, else NULL. */
int nrv_ingest_view(const nrv_ingest* r, nrv_batch* batch, const int32_t** file_status, const int64_t** read_file,
                    const int64_t** a0);
/* names[n_reads]: the member name (read_<id>) of every packed read that came from a multi-read container, "" for reads of
 * single-read files (those are named after their file, NanoReviser.py:137).  Valid until nrv_ingest_free. */
int nrv_ingest_read_names(const nrv_ingest* r, const char* const** names);
void nrv_ingest_free(nrv_ingest* r);

#ifdef __cplusplus
}
#endif
#endif /* NRV_H */
