/*
 * deepmod_b200 -- C ABI of the B200-native DeepMod `detect` hot path.
 *
 * The reference (WGLab/DeepMod) is pure Python over TensorFlow 1.x and has no
 * FFI of its own; the seams this library replaces are (paths relative to the
 * reference tree):
 *
 *   - the session tuple  sp_options['rnn'] = (sess, X, Y, init_l, mfpred)
 *       bin/DeepMod_scripts/myDetect.py:972, consumed at :805 and :816-820
 *       -> dm_forward_windows()
 *   - get_Feature()      bin/DeepMod_scripts/myDetect.py:839-903
 *     mPredict1()        bin/DeepMod_scripts/myDetect.py:787-834
 *       -> dm_detect_batch() (features, windows, BiLSTM, label write-back, and
 *          the per-position accumulation of sum_handler :1089-1100, fused)
 *   - sum_handler() BED writer  bin/DeepMod_scripts/myDetect.py:1107-1120
 *       -> dm_hist_nonzero() / dm_write_bed()
 *   - model restore      bin/DeepMod_scripts/myDetect.py:950-956
 *       -> dm_create() takes the 14 restored tensors in the reference layout
 *
 * Conventions: plain pointers and sizes only; every pointer is HOST memory unless
 * the name ends in _d; every function returns 0 on success or a negative dm_status
 * and never throws.  A context is bound to one GPU and must be driven by one host
 * thread at a time.  The library has no CPU fallback: without a CUDA device
 * dm_create() fails with DM_ERR_CUDA.
 */
#ifndef DEEPMOD_B200_H
#define DEEPMOD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_WINDOW   21   /* --windowsize default, bin/DeepMod.py:317 */
#define DM_FNUM      7   /* --fnum default,       bin/DeepMod.py:318 */
#define DM_HIDDEN  100   /* --hidden default,     bin/DeepMod.py:319 */
#define DM_LIVE_STEPS 11 /* outputs[int(21/2)]: myMultiBiRNN.py:55  */

typedef struct dm_ctx dm_ctx;

enum dm_status {
  DM_OK = 0,
  DM_ERR_ARG = -1,
  DM_ERR_CUDA = -2,
  DM_ERR_STATE = -3,
  DM_ERR_IO = -4,
  DM_ERR_OVERFLOW = -5,  /* a per-position counter reached its limit (2^28 - 1): the accumulator is no longer exact */
  DM_ERR_NCCL = -6
};

/* arithmetic of the BiLSTM */
enum dm_precision {
  DM_FP32 = 0,   /* fp32 FMA + accurate expf/tanhf: the parity path (<=1e-4 on p1) */
  DM_BF16 = 1,   /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM; CTA pairs (cta_group::2) */
  DM_BF16_1CTA = 2, /* same arithmetic, one CTA per 128-window tile (cta_group::1); kept for A/B measurements */
  DM_F16 = 3     /* fp16 operands (weights and hidden state) on the same tensor-core kernel: same rate as DM_BF16,
                    three more mantissa bits on every operand; the throughput default */
};

/* per-read status written by dm_detect_batch (mirrors sp_param['f5status']) */
enum dm_read_status {
  DM_READ_OK = 0,
  DM_READ_MISMATCH = 1,    /* 'Error Does not match', myDetect.py:868-874 */
  DM_READ_BAD_ALIGN = 2,   /* #non-gap columns != mapped events (reference would raise) */
  DM_READ_LESS_EVENT = 3,  /* 'Less Event', myDetect.py:702-705 */
  DM_READ_NO_MATCH = 4     /* alignment without a single matching base, myDetect.py:622-627 (dm_align_upload only) */
};

/* The 14 inference tensors exactly as Saver.restore leaves them (fp32, row-major):
 * kernel[d][0] is [107,400], kernel[d][1..2] are [200,400], bias[d][l] is [400];
 * rows = [input | h], columns = [i | j | f | o] (BasicLSTMCell); d: 0 = fw, 1 = bw.
 * cls_w is `Variable` [200,2] (rows = [fw h | bw h]), cls_b is `Variable_1` [2]. */
typedef struct dm_weights {
  const float* kernel[2][3];
  const float* bias[2][3];
  const float* cls_w;
  const float* cls_b;
} dm_weights;

/* One packed batch of aligned reads (host pointers; see deepmod_b200/synth.py).
 * Columns and clips are in READ orientation ('-' strand already flipped and
 * complemented, myDetect.py:661-666). */
typedef struct dm_batch {
  int32_t        n_reads;
  const int64_t* ev_off;       /* [n_reads+1] */
  const float*   ev_mean;      /* [ev_off[n]] normalised mean,   myDetect.py:898 */
  const float*   ev_stdv;      /*             normalised stdv,   :899 */
  const float*   ev_len;       /*             raw-sample count,  :900 */
  const uint8_t* ev_base;      /* ASCII k-mer centre (model_state[2]); NULL = skip the :868 check */
  const int64_t* col_off;      /* [n_reads+1] */
  const uint8_t* col_refbase;  /* ASCII, '-' = insertion */
  const uint8_t* col_readbase; /* ASCII, '-' = deletion  */
  const int64_t* col_refpos;   /* 0-based reference position of the column */
  const int32_t* start_clip;   /* [n_reads] */
  const int32_t* end_clip;     /* [n_reads] */
  const int32_t* contig;       /* [n_reads] index into dm_set_genome's contigs */
  const int8_t*  strand;       /* [n_reads] +1 / -1 */
} dm_batch;

/* ---- lifetime ---------------------------------------------------------- */
int  dm_create(dm_ctx** out, int device, const dm_weights* w, int precision);
void dm_destroy(dm_ctx* ctx);
const char* dm_last_error(const dm_ctx* ctx);   /* ctx may be NULL: last global error */
int  dm_version(void);
/* pick the arithmetic of subsequent calls (both weight images are always resident) */
int  dm_set_precision(dm_ctx* ctx, int precision);

/* ---- model only (b1 seam) ---------------------------------------------- */
/* X is [n,21,7] fp32 windows as mPredict1 builds them (:791-803).  p1_out[n]
 * receives softmax(logits)[:,1]; pred_out[n] receives argmax (what `mfpred`
 * fetches, myMultiBiRNN.py:61).  Either output may be NULL. */
int dm_forward_windows(dm_ctx* ctx, int64_t n, const float* X, float* p1_out, uint8_t* pred_out);

/* ---- per-position accumulator (sum_handler's dict) ----------------------- */
/* Allocates and zeroes one (cov, mod, del) cell per reference position and strand
 * for `base` (the --Base of interest, bin/DeepMod.py:331). */
int dm_set_genome(dm_ctx* ctx, int32_t n_contigs, const int64_t* contig_len, char base);
int dm_hist_clear(dm_ctx* ctx);
/* Device pointer + length (in uint64 cells) of the whole accumulator (cell = cov | mod << 28 | key-created
 * flag << 56, so a sum of cells is the sum of counters; dm_reduce* is the supported way to merge GPUs). */
int dm_hist_device_ptr(dm_ctx* ctx, void** cells_d, int64_t* n_cells);

/* ---- the job's single exchange step: sum of the accumulators of all GPUs ----------------
 * Reference equivalent: the dict accumulation over all reads (myDetect.py:1089-1100) and the offline merge of
 * per-run BED files, DeepMod_tools/sum_chr_mod.py:36-63.  Integer sums: bit-exact in any order.  Afterwards
 * EVERY context holds the merged accumulator.  NCCL (libnccl.so.2) is loaded on first use; without it these
 * calls fail with DM_ERR_NCCL.  Both check max(cov) * ranks against the counter limit first and return
 * DM_ERR_OVERFLOW on every rank, without summing, if the merged counters might not fit.
 *
 * One process driving n contexts on n distinct devices (ncclCommInitAll + one grouped ncclAllReduce): */
int dm_reduce(dm_ctx** ctxs, int n);
/* One process per GPU: rank 0 makes an id (dm_reduce_unique_id), the host shares its 128 bytes with the other
 * ranks by any means, then every rank calls dm_reduce_comm(ctx, id, rank, n).  The communicator is kept in the
 * context: later calls may pass id == NULL. */
int dm_reduce_unique_id(uint8_t id_out[128]);
int dm_reduce_comm(dm_ctx* ctx, const uint8_t* id, int rank, int n_ranks);
/* Destroys the communicator dm_reduce_comm keeps in the context.  ncclCommDestroy is a COLLECTIVE: it returns once every
 * rank has called it, so call this on all ranks together, right after the last exchange (dm_destroy does it too: a
 * rank that destroys its context early waits here for the others; a rank that never does leaves them waiting). */
int dm_reduce_finalize(dm_ctx* ctx);
/* dst += src for two contexts of ONE process holding the same genome (same or different GPU of the box): counters
 * add, key-created flags OR -- the merge of DeepMod_tools/sum_chr_mod.py:47-52 without going through BED files. */
int dm_hist_merge(dm_ctx* dst, dm_ctx* src);
/* Conservation counters of the accumulator (what a merge must preserve): sum of cov, sum of mod, number of
 * existing rows (cells != 0), and a position-weighted checksum (sum over cells of (cov + 3 mod) * (index mod 65521 + 1),
 * mod 2^64) that is linear in the cells like the counters themselves.  Any output may be NULL. */
int dm_hist_totals(dm_ctx* ctx, uint64_t* sum_cov, uint64_t* sum_mod, uint64_t* n_rows, uint64_t* checksum);
/* Device time of the last exchange (CUDA events around the all-reduce on the context's stream). */
int dm_last_reduce_ms(const dm_ctx* ctx, float* ms);
/* Rows that exist in the reference's dict for (contig, strand): positions touched
 * by at least one alignment column whose refbase == base (deletions included,
 * myDetect.py:1093-1094), ascending.  Call with pos == NULL to get the count. */
int dm_hist_nonzero(dm_ctx* ctx, int32_t contig, int8_t strand, int64_t cap,
                    int64_t* pos, int32_t* cov, int32_t* mod, int64_t* n_rows);
/* Writes mod_pos.<chr><strand>.<Base>.bed in sum_handler's exact text format
 * (myDetect.py:1116-1120).  No file is created when there are no rows (:1109). */
int dm_write_bed(dm_ctx* ctx, int32_t contig, int8_t strand, const char* chrom,
                 const char* path, int64_t* n_rows);

/* sum_handler's loop over STORED per-read predictions (--predDet 0: read_pred_detail + myDetect.py:1089-1100): n
 * records of one read mapped to (contig, strand) -- the columns of its `predetail` dataset (:720-753): reference base,
 * read base ('-' = deletion), reference position, stored prediction -- accumulate exactly like dm_detect_batch's. */
int dm_accumulate_records(dm_ctx* ctx, int32_t contig, int8_t strand, int64_t n, const uint8_t* refbase,
                          const uint8_t* readbase, const int64_t* refpos, const int8_t* mod_pred);

/* Overwrite cells of (contig, strand) with summary rows read elsewhere (e.g. BED files of earlier runs:
 * DeepMod_tools/sum_chr_mod.py:36-45 readbed2); deletion-touch counters of those cells become 0. */
int dm_hist_load(dm_ctx* ctx, int32_t contig, int8_t strand, int64_t n, const int64_t* pos,
                 const int32_t* cov, const int32_t* mod);
/* Merged summary of one contig in sum_chr_mod.py's format (:47-63): both strands interleaved by position,
 * rows with mod == 0 dropped, "%s %d %d %s %d %s  %d %d 0,0,0 %d %d %d".  No file when there are no rows. */
int dm_write_merged_bed(dm_ctx* ctx, int32_t contig, const char* chrom, const char* path, int64_t* n_rows);

/* ---- CpG-cluster second pass (DeepMod_tools/hm_cluster_predict.py) ---------------------------- */
/* MLP of train_deepmod/na12878_cluster_train_mod-*: W_1[14,100] b_1[100] W_2[100,20] b_2[20] W_O[20,1] b_O[1] */
typedef struct dm_cluster_weights {
  const float *w1, *b1, *w2, *b2, *wo, *bo;
} dm_cluster_weights;
/* Motif (CpG) sites of a contig, what the reference reads from motif_<chr>_C.bed
 * (hm_cluster_predict.py:117-123, written by generate_motif_pos.py:56-71).  Replaces earlier sites. */
int dm_cluster_set_sites(dm_ctx* ctx, int32_t contig, int64_t n, const int64_t* pos, const int8_t* strand);
/* Sites = motif sites whose cell has cov > 0 (and mod > 0 when drop_unmodified, the rows sum_chr_mod.py keeps),
 * ordered by (strand '+' first, position) like the script's sorted keys.  Outputs (any may be NULL) hold up to
 * `cap` sites: features[cap,14] as fed to the model, prob = sigmoid output, pct = int(prob*100) (:170). */
int dm_cluster_predict(dm_ctx* ctx, int32_t contig, const dm_cluster_weights* w, int drop_unmodified, int64_t cap,
                       int64_t* pos, int8_t* strand, int32_t* cov, int32_t* mod, float* features, float* prob,
                       int32_t* pct, int64_t* n_sites);
/* "<merged line> <pct>" per site into <prefix>_clusterCpG.<chr>.C.bed's format (:168-170). */
int dm_write_cluster_bed(dm_ctx* ctx, int32_t contig, const dm_cluster_weights* w, int drop_unmodified,
                         const char* chrom, const char* path, int64_t* n_rows);

/* ---- the hot path ------------------------------------------------------- */
/* get_Feature + mPredict1 + reducer for a packed batch, host buffers in, host
 * results out.  p1_out / pred_out are indexed by window (= mapped event), reads
 * concatenated in order, sum(Lmap) entries; windows of rejected reads hold 0.
 * status_out[n_reads] receives dm_read_status.  Any output may be NULL. */
int dm_detect_batch(dm_ctx* ctx, const dm_batch* b, float* p1_out, uint8_t* pred_out,
                    int32_t* status_out);
/* Page-locked host memory for input staging (optional: dm_detect_batch takes any host pointer; from these buffers its
 * copies overlap the kernels without a driver-side staging pass). */
int  dm_pinned_alloc(size_t bytes, int device /* the GPU the calling thread will feed; < 0: its current device */, void** out);
void dm_pinned_free(void* p);
/* dm_detect_batch cuts a large batch into contiguous read ranges and alternates them between two
 * device slots / streams, so the host<->device copies of one range run under the kernels of the
 * other (results are identical: reads are independent, the accumulator is a sum).  parts: 0 = decide
 * by batch size (default), 1 = never, n = always n ranges.  After dm_detect_batch the resident
 * batch (dm_detect_resident, dm_build_windows, ...) is unspecified; use dm_batch_upload for those. */
int dm_set_pipeline(dm_ctx* ctx, int parts);

/* ---- event-table front-end: raw signal -> per-event statistics ----------------------------------------- */
/* mnormalized (myDetect.py:266-282) + the per-event mean / stdv loop of getFast5Info (:334-343) for a batch of
 * reads: raw int16 samples (raw_off per read), events as (start, length) in samples relative to the read's raw
 * array.  mean_out / stdv_out [ev_off[n_reads]] receive what the reference stores in the '<f4' fields. */
int dm_event_stats(dm_ctx* ctx, int32_t n_reads, const int64_t* raw_off, const int16_t* raw, const int64_t* ev_off,
                   const int64_t* ev_start, const int64_t* ev_length, float* mean_out, float* stdv_out);

/* ---- from alignment records (SAM) instead of ready-made columns ---------------------------------------- */
/* Reference sequence of a contig (upper-case ASCII, length = the contig length given to dm_set_genome); what
 * getRefSeq fetches with `samtools faidx` (myDetect.py:470-483). */
int dm_set_contig_sequence(dm_ctx* ctx, int32_t contig, const uint8_t* seq, int64_t len);

/* One best-MAPQ record per read (handle_line, myDetect.py:929-943) after the removal of leading / trailing
 * non-aligned CIGAR ops (:527-540), tokenised by the host (deepmod_b200/sam.py):
 *   clip_left/right  bases clipped at the alignment's left / right end (S, H, leading I, X ...), :527-540
 *   ref_start        0-based reference position of the first remaining op
 *   op_code/op_len   remaining ops, ASCII codes out of "MIDNSHP=X"; op_off[r]..op_off[r+1] per read
 *   seq              SEQ without the clipped bases (reference orientation); seq_off per read
 * Event arrays as in dm_batch (sequencing order). */
typedef struct dm_sam_batch {
  int32_t        n_reads;
  const int64_t* ev_off;
  const float*   ev_mean;
  const float*   ev_stdv;
  const float*   ev_len;
  const uint8_t* ev_base;
  const int32_t* contig;
  const int8_t*  strand;       /* +1 / -1 (flag & 0x10) */
  const int64_t* ref_start;
  const int32_t* clip_left;
  const int32_t* clip_right;
  const int64_t* op_off;       /* [n_reads+1] */
  const uint8_t* op_code;
  const int32_t* op_len;
  const int64_t* seq_off;      /* [n_reads+1] */
  const uint8_t* seq;
} dm_sam_batch;
/* The CIGAR walk of handle_record (myDetect.py:565-705) on the GPU; leaves the batch resident exactly as
 * dm_batch_upload would (then dm_detect_resident / dm_fetch_results). */
int dm_align_upload(dm_ctx* ctx, const dm_sam_batch* sb, int64_t* n_windows, int64_t* n_cols);
/* The alignment columns of the resident batch (= base_map_info): col_off[n_reads+1], columns, final clips in read
 * orientation.  Any pointer may be NULL. */
int dm_fetch_alignment(dm_ctx* ctx, int64_t* col_off, uint8_t* refbase, uint8_t* readbase, int64_t* refpos,
                       int32_t* start_clip, int32_t* end_clip);

/* ---- synthetic reads generated on the device (benchmark workload: SURVEY 8(d) row 2, BASELINE configs[2]) ------ */
/* Read `id` of the set is a pure function of (seed, id): Philox4x32-10 counters; length ~ Gamma(2, mean_len / 2) clipped to
 * [len_lo, len_hi], clips ~ U{0..max_clip}, all-match alignment on the contigs of dm_set_genome (iid ACGT genome, also a
 * function of the seed), one event per base with the distributions of deepmod_b200/synth.py.  Any GPU can generate any
 * range of ids, so one read set shards over 1..8 GPUs and must reduce to the same accumulator. */
typedef struct dm_synth_spec {
  uint64_t seed;
  float    mean_len;
  int32_t  len_lo, len_hi, max_clip;
  int32_t  length_kind;   /* 0: Gamma(2, mean_len / 2) clipped to [len_lo, len_hi]; 1: log-uniform in [len_lo, len_hi] (configs[3]) */
} dm_synth_spec;
/* Events and windows of reads first_read .. first_read + n_reads - 1 (for cutting ranges balanced by mapped bases). */
int dm_synth_describe(dm_ctx* ctx, const dm_synth_spec* spec, int64_t first_read, int32_t n_reads,
                      int32_t* n_events_out, int32_t* n_windows_out);
/* Generate those reads straight into the resident batch, as dm_batch_upload would leave it. */
int dm_synth_generate(dm_ctx* ctx, const dm_synth_spec* spec, int64_t first_read, int32_t n_reads, int64_t* n_windows);
/* Sizes and input arrays of the resident batch (to host buffers sized by dm_resident_sizes; any pointer may be NULL). */
int dm_resident_sizes(dm_ctx* ctx, int32_t* n_reads, int64_t* n_events, int64_t* n_cols, int64_t* n_windows);
int dm_fetch_inputs(dm_ctx* ctx, int64_t* ev_off, float* ev_mean, float* ev_stdv, float* ev_len, uint8_t* ev_base,
                    int64_t* col_off, uint8_t* col_refbase, uint8_t* col_readbase, int64_t* col_refpos,
                    int32_t* start_clip, int32_t* end_clip, int32_t* contig, int8_t* strand);

/* Same work with the batch resident in HBM: upload once, run many times. */
int dm_batch_upload(dm_ctx* ctx, const dm_batch* b, int64_t* n_windows);
int dm_detect_resident(dm_ctx* ctx, int accumulate /* 0: skip the histogram update */);
int dm_fetch_results(dm_ctx* ctx, float* p1_out, uint8_t* pred_out, int32_t* status_out);

/* Materialise the [sum(Lmap),21,7] fp32 windows of the uploaded batch exactly as
 * mPredict1 slices them (:791-803) -- parity hook for the gather kernel. */
int dm_build_windows(dm_ctx* ctx, float* windows_out);

/* ---- instrumentation ------------------------------------------------------ */
/* Number of kernels this library launched since creation, and device time of the
 * BiLSTM kernels of the last dm_detect_* call (CUDA events on the library's stream). */
int64_t dm_launch_count(const dm_ctx* ctx);
int dm_last_timing(const dm_ctx* ctx, float* lstm_ms, float* total_ms);
/* tcgen05 descriptor/layout self-test: one 128 x n x k bf16 GEMM through the same
 * UMMA + TMEM path the BiLSTM uses; returns max |err| vs fp32 in *max_err. */
int dm_selftest_umma(dm_ctx* ctx, int n, int k, float* max_err);
/* Debug hook of the tensor-core BiLSTM: run only the first `max_steps` (1..66) cell-steps
 * (wavefront order, fw then bw) on windows X[n,21,7] and copy the first tile's operand
 * region of shared memory (2 x-columns + 5 hidden tiles, 137216 bytes, UMMA canonical
 * K-major layout) into `dump`.  p1_out[n] is only meaningful for max_steps == 66. */
int dm_debug_tc_windows(dm_ctx* ctx, int64_t n, const float* X, int max_steps, uint8_t* dump,
                        int64_t dump_cap, float* p1_out);

#ifdef __cplusplus
}
#endif
#endif /* DEEPMOD_B200_H */
