/* dynamont_b200 — C ABI of the B200-native Dynamont segmentation hot path.
 *
 * This is the drop-in boundary.  The reference has no C ABI: its native boundary is the pybind11 class
 * `dynamont._dynamont.Aligner` (reference src/cpp/aligner_bindings.cpp:111-167, 180-219) over
 * `dynamont::Aligner::align / train` (reference include/dynamont/aligner.hpp:56-85).  Each entry point
 * below names the reference interface it replaces.  Plain pointers and sizes only; no C++/torch types.
 *
 * All functions are thread-safe per handle (calls on one handle are serialised internally).
 * There is no CPU fallback: creating a handle fails when no CUDA device is usable.
 */
#ifndef DYNAMONT_B200_H
#define DYNAMONT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dyn_aligner dyn_aligner;

/* Per-read status.  Values 1..6 correspond one-to-one to the exceptions the reference throws for that
 * read (message text via dyn_status_message, identical to the reference's strings). */
enum
{
	DYN_OK = 0,
	DYN_SIGNAL_EMPTY = 1,   /* aligner.cpp:151  "Signal is empty" */
	DYN_SEQ_SHORT = 2,      /* aligner.cpp:156  "Sequence shorter than model kmer size" */
	DYN_SIGNAL_SHORT = 3,   /* aligner.cpp:162  "Signal too short compared to sequence" */
	DYN_INVALID_NT = 4,     /* aligner.cpp:182,194  "Invalid nucleotide: X" (X in dyn_read_result.bad_char) */
	DYN_ALIGN_FAILED = 5,   /* NT_aligner_api.cpp:291  "Alignment failed: alignment scores do not match" */
	DYN_TRAIN_FAILED = 6,   /* NT_aligner_api.cpp:625  "Training failed: alignment scores do not match" */
	DYN_REC_OVERFLOW = 7,   /* internal, never returned: resolved by an automatic retry */
	DYN_INTERNAL = 8,
	DYN_BAND_UNSUPPORTED = 9, /* band wider than this build's ring capacity */
	DYN_NTK_TN_FAILED = 11,  /* NTK_aligner_api.cpp:335  "NTK preprocessing TN failed: alignment scores do not match" */
	DYN_NTK_TK_FAILED = 12,  /* NTK_aligner_api.cpp:381  "NTK preprocessing TK failed: alignment scores do not match" */
	DYN_NTK_ALIGN_FAILED = 13 /* NTK_aligner_api.cpp:916  "NTK alignment failed: alignment scores do not match" */
};

typedef struct
{
	int32_t status;        /* DYN_* */
	char bad_char;         /* offending character for DYN_INVALID_NT */
	char pad[3];
	double Z;              /* Result.Z (aligner.hpp:44-48): natural-log partition function (backward) */
	uint64_t seg_offset;   /* first entry of this read in the segment output arrays */
	uint64_t n_segments;   /* Kc = L - k + 1 on success with calc_probabilities, else 0 */
} dyn_read_result;

typedef struct
{
	int32_t status;
	char bad_char;
	char pad[3];
	double Z;              /* TrainingResult.Z (aligner.hpp:50-55) */
	double m1, e1, e2;     /* TrainingResult.transitions: re-estimated probabilities (NT:703-722) */
} dyn_train_result;

/* Replaces PyAligner's constructor / makeAligner (aligner_bindings.cpp:34-51,111-130) and
 * Aligner::Aligner (aligner.cpp:13-36).  pore: "rna002" | "rna004" | "dna_r9" | "dna_r10_260bps" |
 * "dna_r10_400bps"; mode: "basic" | "nt" -> NTAligner, "resquiggle" | "ntk" -> NTKAligner (served by dyn_ntk_align*); threads is accepted and
 * ignored exactly like the reference; band as in the reference (default 400);
 * device = CUDA ordinal or -1 for the current device.  Returns NULL and fills err on failure; err_kind is
 * set to 1 for the reference's std::invalid_argument cases (-> ValueError), 0 for runtime_error. */
dyn_aligner* dyn_create(const char* model_path, const char* pore, const char* mode, int threads, int band,
	int device, char* err, size_t errlen, int* err_kind);
void dyn_destroy(dyn_aligner*);

int dyn_kmer_size(const dyn_aligner*);
uint64_t dyn_num_kmers(const dyn_aligner*);
int dyn_is_rna(const dyn_aligner*);
/* model table in native index order (aligner.cpp:136-141): mean[K], stdev[K] */
void dyn_model(const dyn_aligner*, double* mean, double* stdev);
/* replace the emission table (native order); used by the training driver between iterations */
int dyn_set_model(dyn_aligner*, const double* mean, const double* stdev);
/* log transition parameters {m1, e1, e2} (NT_aligner_api.cpp:84-86) */
void dyn_transitions(const dyn_aligner*, double* log3);

/* number of output segments a batch needs: sum over reads of max(L - k + 1, 0) */
uint64_t dyn_count_segments(const dyn_aligner*, const uint64_t* seq_off, uint32_t n_reads);
/* in-band DP cells of one read (forward's trip count, NT:122-141) — the GCUPS unit */
uint64_t dyn_read_cells(const dyn_aligner*, uint64_t S, uint64_t L);
/* same for a whole batch (invalid reads count 0); per_read may be NULL; returns the total */
uint64_t dyn_batch_cells(const dyn_aligner*, const uint64_t* sig_off, const uint64_t* seq_off, uint32_t n_reads,
	uint64_t* per_read);

/* Batched Aligner::align (NT_aligner_api.cpp:230-312).  Read r has samples
 * signal[sig_off[r] .. sig_off[r+1]) and bases seq[seq_off[r] .. seq_off[r+1]) (no terminators).
 * results[n_reads]; the three segment arrays hold dyn_count_segments() entries and are filled at
 * results[r].seg_offset: Segment.sequencePosition, Segment.signalPosition, Segment.probability
 * (aligner.hpp:35-42; state is always 'M' and polish empty in basic mode).  calc_probabilities = 0 computes
 * Z only.  Returns 0, or -1 on a CUDA/runtime error (message via dyn_last_error). One bad read never fails the
 * batch (reference front end: segment.py:160-176). */
int dyn_align_batch(dyn_aligner*, const float* signal, const uint64_t* sig_off, const char* seq,
	const uint64_t* seq_off, uint32_t n_reads, int calc_probabilities, dyn_read_result* results,
	uint64_t* sequence_positions, uint64_t* signal_positions, double* probabilities);

/* Same, float64 samples as in the reference signature (converted to FP32 with round-to-nearest). */
int dyn_align_batch_f64(dyn_aligner*, const double* signal, const uint64_t* sig_off, const char* seq,
	const uint64_t* seq_off, uint32_t n_reads, int calc_probabilities, dyn_read_result* results,
	uint64_t* sequence_positions, uint64_t* signal_positions, double* probabilities);

/* Same, with the samples and bases already resident in device memory (d_signal, d_seq are device pointers;
 * offsets and outputs are host pointers). */
int dyn_align_batch_device(dyn_aligner*, const float* d_signal, const uint64_t* sig_off, const char* d_seq,
	const uint64_t* seq_off, uint32_t n_reads, int calc_probabilities, dyn_read_result* results,
	uint64_t* sequence_positions, uint64_t* signal_positions, double* probabilities);

/* Batched Aligner::train (NT_aligner_api.cpp:567-639 + runTraining :462-561 + trainTransition :641-725).
 * results[n_reads] receives per-read Z and re-estimated transitions.  Pooled sufficient statistics over all
 * successful reads of the batch are ADDED to pooled_w / pooled_x / pooled_xx (K doubles each, native kmer
 * order; NT:510-512) and to pooled_xi[2] = expected {E->M, E->E} transition counts — these are what a
 * data-parallel trainer all-reduces.  If per_read_mean/per_read_stdev are non-NULL they receive the
 * reference's per-read M-step (n_reads x K doubles each, NT:519-535: kmers without weight keep the model). */
int dyn_train_batch(dyn_aligner*, const float* signal, const uint64_t* sig_off, const char* seq,
	const uint64_t* seq_off, uint32_t n_reads, dyn_train_result* results, double* pooled_w, double* pooled_x,
	double* pooled_xx, double* pooled_xi, double* per_read_mean, double* per_read_stdev);

/* the reference's exact message for a status; "Invalid nucleotide: " lacks the char */
const char* dyn_status_message(int status);
/* message of the last runtime failure on this handle */
const char* dyn_last_error(const dyn_aligner*);

/* ---- front-end stages either side of the DP ------------------------------------------------
 * dyn_preprocess_batch: (raw - shift) / scale followed by the Hampel outlier filter (utils.py:16-43; segment.py:151-153
 *   uses window 3 / 3 sigma, train.py:168-170 window 7 / 5 sigma), computed on the GPU in float64 like the numpy code
 *   and rounded to FP32, the sample type the DP entry points take.  raw / out: concatenated samples with sig_off[n+1].
 * dyn_format_segments: utils.segmentation_to_string (utils.py:193-232) — the CSV lines of one read
 *   "readid,signalid,start,end,basepos,base,motif,state,prob(.6f),polish\n"; polishes may be NULL (basic mode: "NA").
 *   Returns the byte length (the text is written when it fits out_cap). */
int dyn_preprocess_batch(dyn_aligner*, const float* raw, const uint64_t* sig_off, uint32_t n_reads, const double* shift,
	const double* scale, int window, double n_sigmas, float* out);
int64_t dyn_format_segments(const char* readid, const char* signalid, int64_t sig_offset, int64_t last_index, const char* read,
	int kmer_size, int rna, uint64_t n_segments, const uint64_t* sequence_positions, const uint64_t* signal_positions,
	const double* probabilities, const char* states, const char* const* polishes, char* out, uint64_t out_cap);

/* ---- resquiggle ("NTK") mode, pre-pass stages (reference NTK_aligner_api.cpp:120-441; rows B1-B6 of SURVEY.md 8a).
 * A handle created with mode "resquiggle"/"ntk" is served by dyn_ntk_align (one read per call: the reference's
 * NTKAligner::align, NTK:881-927) and by the stage entry points below; the batched basic-mode entry points fail on it.
 *
 * dyn_ntk_transitions: the 14 log transition scores a1,a2,p1,p2,p3,s1,s2,s3,e1,e2,e3,e4,i1,i2 followed by
 *   log ntMatch / log ntExtend of the TN and of the TK pre-pass (NTKAligner::initializeTransitions, NTK:35-104).
 * dyn_ntk_prepass: one read.  tn_mask [T][ceil(N/32)] and tk_mask [T][ceil(K/32)] receive the posterior-mass row
 *   sets tnMap / tkMap of preProcTN / preProcTK (NTK:315-400) as bitmaps (bit i of word i/32); keys receives the sorted
 *   key list t*N*K + n*K + q of preProcTNK (NTK:402-441) when it fits keys_cap, *n_keys its length either way;
 *   z4 = {Zf, Zb} of the TN pass then of the TK pass.  T = S + 1, N = L - k + 2, K = 4^k.
 *   Returns 0, a dyn_status (input validation as in dyn_align_batch, DYN_NTK_TN_FAILED, DYN_NTK_TK_FAILED) or -1. */
void dyn_ntk_transitions(const dyn_aligner*, double* out18);
int dyn_ntk_prepass(dyn_aligner*, const float* signal, uint64_t S, const char* seq, uint64_t L, uint32_t* tn_mask,
	uint32_t* tk_mask, uint64_t* keys, uint64_t keys_cap, uint64_t* n_keys, double* z4);
/* dyn_ntk_align = NTKAligner::align (NTK_aligner_api.cpp:881-927) for one read: pre-passes, sparse 5-state forward /
 * backward over the keys (logF / logB, :443-607, each result visible to later keys at once — the repair described in
 * INTEGRATION.md; as shipped the reference throws for every input in this mode), Z check (:913-918), sparse posteriors,
 * MAP fill and traceback (:628-879).  Output arrays hold `cap` entries (cap >= S + L is always enough); segment i has
 * states[i] in {'M','P'}, sequence_positions[i], signal_positions[i], probabilities[i] and polish_kmers[i], the id of
 * the polished kmer in the aligner's native order (Aligner::intToKmer, aligner.cpp:222-239, turns it into the
 * string; RNA pores reverse it).  calc_probabilities = 0 computes Z only.  Returns 0, a dyn_status or -1. */
int dyn_ntk_align(dyn_aligner*, const float* signal, uint64_t S, const char* seq, uint64_t L, int calc_probabilities,
	double* Z, uint64_t* n_segments, char* states, uint64_t* sequence_positions, uint64_t* signal_positions,
	double* probabilities, uint32_t* polish_kmers, uint64_t cap);
/* The same for a batch of independent reads (inputs laid out like dyn_align_batch): read r writes status[r], Z[r],
 * n_segments[r] and its segments at [out_off[r], out_off[r+1]) of the output arrays (out_off[r+1] - out_off[r] >=
 * S_r + L_r + 16).  `concurrency` host threads (0 = 32), each with its own CUDA stream and device buffers, process
 * the reads so that their small grids overlap on the device.  Returns 0, or -1 on a CUDA/runtime error. */
int dyn_ntk_align_batch(dyn_aligner*, const float* signal, const uint64_t* sig_off, const char* seq, const uint64_t* seq_off,
	uint32_t n_reads, int calc_probabilities, int32_t* status, double* Z, uint64_t* n_segments, const uint64_t* out_off,
	char* states, uint64_t* sequence_positions, uint64_t* signal_positions, double* probabilities, uint32_t* polish_kmers,
	int concurrency);

/* Pooled (data-parallel) Baum-Welch with the statistics resident on the device — what dynamont-train's per-batch
 * loop (train.py:185-223) becomes when the expected counts are pooled instead of averaging per-read estimates.
 * dyn_train_accumulate ADDS the sufficient statistics of the batch's successful reads to the caller's DEVICE buffer
 * d_stats[3K + 4] (doubles): w[K], x[K], xx[K] (NT:510-512, native kmer order), then expected E->M and E->E transition
 * counts (NT:641-725), sum of Z, number of successful reads.  Nothing but the per-read status (status[n_reads], host,
 * may be NULL) returns to the host, so ranks can all-reduce d_stats in place (ncclAllReduce / torch.distributed on the
 * device pointer).  signal / seq are host pointers, or device pointers if inputs_on_device.
 * dyn_train_mstep_device applies the M-step (NT:519-535: mean, stdev with the 1e-12 variance floor, kmers without
 * weight keep the model) to d_stats on the device, installs the result as the handle's model, and returns the
 * re-estimated transitions {m1, e1, e2} (NT:703-722) in transitions3 (host, may be NULL). */
int dyn_train_accumulate(dyn_aligner*, const float* signal, const uint64_t* sig_off, const char* seq, const uint64_t* seq_off,
	uint32_t n_reads, int inputs_on_device, double* d_stats, int32_t* status);
int dyn_train_mstep_device(dyn_aligner*, const double* d_stats, double* transitions3);

/* Asynchronous form of dyn_align_batch for callers that stream batches: dyn_align_submit starts the batch on one of two
 * internal lanes (own CUDA stream and device buffers each, same device and model as the handle) and returns a ticket
 * (>= 0; -1 on error); dyn_align_wait blocks until that batch is complete and returns what dyn_align_batch would have
 * returned.  With two batches in flight the host-to-device copy of the next batch and the device-to-host copy + result
 * fan-out of the previous one overlap the kernels of the current one.  All input and output buffers of a submitted
 * batch must stay valid and untouched until its wait returns; every ticket must be waited for exactly once before
 * dyn_destroy.  Replaces the mp.Pool fan-out of the reference front end (segment.py:296-324) for one GPU. */
int64_t dyn_align_submit(dyn_aligner*, const float* signal, const uint64_t* sig_off, const char* seq, const uint64_t* seq_off,
	uint32_t n_reads, int calc_probabilities, dyn_read_result* results, uint64_t* sequence_positions,
	uint64_t* signal_positions, double* probabilities);
/* the same with the samples and bases already resident in device memory (cf. dyn_align_batch_device) */
int64_t dyn_align_submit_device(dyn_aligner*, const float* d_signal, const uint64_t* sig_off, const char* d_seq,
	const uint64_t* seq_off, uint32_t n_reads, int calc_probabilities, dyn_read_result* results, uint64_t* sequence_positions,
	uint64_t* signal_positions, double* probabilities);
int dyn_align_wait(dyn_aligner*, int64_t ticket);

/* instrumentation for bench.py: device time (ms, CUDA events on the launching stream) of the kernels of the
 * last batch call: [0] encode/emission-constant kernel, [1] main DP kernel, [2] number of kernel launches */
void dyn_last_timing(const dyn_aligner*, double* out3);
/* number of reads of the last batch call that the FP32 linear-domain kernels could not represent and that were
 * re-run by the log2-domain kernels (same GPU); results are identical either way */
uint64_t dyn_last_fallbacks(const dyn_aligner*);
/* number of reads of the last batch call that the first-tier linear-domain kernels (renormalisation every 8 rows) handed
 * to the second tier (every 4 rows); dyn_last_fallbacks counts what the second tier handed on to the log2 domain */
uint64_t dyn_last_lin_retries(const dyn_aligner*);
/* kernel build variant the last batch call ran (csrc/engine.cu: 3 = general kernels at 8 CTAs/SM, 4 = uniform-sigma
 * kernels at 8 CTAs/SM, ...); -1 before the first call */
int dyn_last_variant(const dyn_aligner*);
/* ribbon kernels (first tier: the same recurrences on a narrow window that follows the probability mass, every loss
 * checked on the device): out2[0] = reads of the last batch call they were given, out2[1] = reads they handed on to
 * the full-band kernels (window / range checks failed); results are identical either way */
void dyn_last_ribbon(const dyn_aligner*, uint64_t* out2);
/* cumulative count of the reads the linear-domain ribbon kernels lost, by reason code 1..12 (csrc/dp_ribbon.cuh,
 * ribbon_read), out16[reason]; out16[13] = posterior records per lattice row of the last batch x 1000, out16[14] = scratch
 * layout of the last batch (0: a checkpoint per group, 1: two-level checkpoints, 2: two-level + records-free),
 * out16[15] = how many of the lost reads the log2-domain ribbon (the tier before the full-band kernels) kept */
void dyn_ribbon_fault_reasons(const dyn_aligner*, uint64_t* out16);
/* run all work of this handle on the caller's CUDA stream (a cudaStream_t, e.g. torch's current stream) instead
 * of the handle's own stream, so that the caller's CUDA events bracket it */
int dyn_set_stream(dyn_aligner*, void* cuda_stream);
/* tuning / experiments, INTEGRATION.md 3c: "ribbon" (0 | 2 | 4 columns per lane), "rib_guard", "rib_two_level", "rib_gather"
 * (records-free scratch layout), "rib_log" (0: no log2-domain ribbon tier, 2: every read through it as well — test hook),
 * "arith", "variant", "warps_per_sm", "mem_fraction", "thr2", ...; returns -1 for an unknown key */
int dyn_set_option(dyn_aligner*, const char* key, double value);

#ifdef __cplusplus
}
#endif
#endif
