// Shared declarations of the deepmod_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/deepmod_b200.h"

#define DM_FLANK 10                     // rows of context either side of a window centre
#define DM_FEAT_STRIDE 8                // floats per feature row (7 features + 1 pad) = 32 B
#define DM_GATES 400                    // 4 * DM_HIDDEN
#define DM_TILE_M 128                   // windows per tensor-core tile (UMMA M)

// One accumulator cell per (strand, reference position), packed in a uint64 so that one atomic updates a
// column and one NCCL sum merges GPUs:
//   bits  0..27  cov   (coverage, < 2^28 = 268 435 456; reaching the limit is REPORTED, never silent:
//   bits 28..55  mod    k_accumulate raises ctx->overflow, dm_reduce checks max(cov) * ranks before it sums)
//   bits 56..63  key-created flag of a deletion column (myDetect.py:1093-1094: the key exists although
//                nothing is counted): 0/1 on one GPU (atomicOr), <= #ranks after the sum, then set back to 0/1
#define DM_CELL_COV_SHIFT 0
#define DM_CELL_MOD_SHIFT 28
#define DM_CELL_DEL_SHIFT 56
#define DM_CELL_MASK 0xFFFFFFFull
#define DM_CELL_DEL_MASK 0xFFull

// ---- packed weight images ---------------------------------------------------
// fp32 image per (dir, layer), [K][400]: layer 0 rows = [x0..x6, 0, h0..h99] (108),
// layers 1,2 rows = [h_below 0..99, h_self 0..99] (200).  Columns are permuted so that the
// four units a thread of dm_lstm_fp32.cu owns are four coalesced float4 loads:
//   n' = j*100 + ug*4 + gate   for unit = j*25 + ug   (reference column = gate*100 + unit)
// The bias image uses the same column order.
#define DM_K0_F32 108
#define DM_K12_F32 200

// bf16 tensor-core image per (dir, layer): B operand of D[128,400] = A[128,K] * B^T,
// stored [k-chunk][n][8] (K-major core matrices, no swizzle), n = 4*unit + gate with
// the 0.5 pre-scale of the three sigmoid gates and the forget bias folded in:
//   layer 0   K = 112: k 0..99 h_self | 100,101 bias hi/lo | 102 mean_lo | 103 stdv_lo
//                      | 104..107 A,C,G,T | 108 mean_hi | 109 stdv_hi | 110 len_hi | 111 len_lo
//   layer 1,2 K = 208: k 0..99 h_below | 100..103 zero | 104..203 h_self | 204,205 bias hi/lo
//                      | 206,207 zero
#define DM_K0_TC 112
#define DM_K12_TC 208

struct dm_dev_weights {
  float* w32[2][3];        // fp32 images  [Kpad][400]
  float* b32[2][3];        // fp32 bias    [400]
  float* cls_w;            // [200][2]
  float* cls_b;            // [2]
  // tensor-core images (see dm_lstm_tc.cu), 16-bit words; first index = operand format (0 bf16, 1 fp16)
  uint16_t* wtc[2][2][3];    // one CTA per tile
  uint16_t* wtc2[2][2][3];   // split by CTA of a pair (cta_group::2)
  float* cls_d;            // [2][100] cls_w[:,1]-cls_w[:,0] per direction
  float cls_db;            // cls_b[1]-cls_b[0]
};

struct dm_dev_batch {
  int32_t n_reads = 0;
  int64_t n_events = 0, n_cols = 0, n_windows = 0, n_frows = 0;
  // inputs
  int64_t *ev_off = nullptr, *col_off = nullptr, *col_refpos = nullptr;
  float *ev_mean = nullptr, *ev_stdv = nullptr, *ev_len = nullptr;
  uint8_t *ev_base = nullptr, *col_refbase = nullptr, *col_readbase = nullptr;
  int32_t *start_clip = nullptr, *end_clip = nullptr, *contig = nullptr;
  int8_t* strand = nullptr;
  // derived
  int64_t *win_off = nullptr;     // [n_reads+1] windows before read r
  int64_t *col_rank = nullptr;    // [n_cols] exclusive count of non-gap columns (global)
  int64_t *win_col = nullptr;     // [n_windows] alignment column of each window centre (-1: none)
  int32_t *win_frow = nullptr;    // [n_windows_padded] first feature row of each window
  int32_t *status = nullptr;      // [n_reads]
  int32_t *align_status = nullptr;  // [n_reads] verdict of the CIGAR walk (dm_align_upload), merged into status on fetch
  float *feat = nullptr;          // [n_frows+21][8] fp32 feature rows (+21 all-zero rows)
  uint16_t* feat_tc = nullptr;    // [n_frows+21][16] 16-bit hi/lo rows for the tensor-core path (bf16 or fp16, ctx->tc_f16)
  float *p1 = nullptr;            // [n_windows_padded]
  uint8_t *pred = nullptr;        // [n_windows_padded]
  // capacity bookkeeping (buffers are grown, never shrunk)
  int64_t cap_reads = 0, cap_events = 0, cap_cols = 0, cap_windows = 0, cap_frows = 0;
  bool has_ev_base = false;
  bool prepared = false;          // derived arrays are current for the uploaded inputs
  bool from_alignment = false;    // columns were produced by dm_align_upload
};

struct dm_ctx {
  int device = 0;
  int precision = DM_FP32;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  dm_dev_weights w{};
  dm_dev_batch b{};               // the uploaded read batch
  dm_dev_batch b2{};              // second slot of the pipelined dm_detect_batch (upload k+1 under compute k)
  cudaStream_t stream2 = nullptr;
  int pipeline_parts = 0;         // dm_set_pipeline: 0 = by batch size, 1 = off, n = always n sub-batches
  void* scratch2 = nullptr; size_t scratch2_bytes = 0;
  dm_dev_batch fw{};              // dm_forward_windows' own table (explicit [n,21,7] windows)
  float* fw_x = nullptr; int64_t fw_x_cap = 0;
  size_t h2d_bytes = 0;           // bytes the last dm_batch_upload moved
  // accumulator
  int32_t n_contigs = 0;
  std::vector<int64_t> contig_len, contig_off;   // contig_off in positions (cells / 2)
  int64_t* contig_off_d = nullptr;
  unsigned long long* cells = nullptr;           // [2][total_len] (strand-major per contig)
  uint8_t* genome = nullptr;                     // [total_len] reference bases (dm_set_contig_sequence)
  uint8_t* motif = nullptr;                      // same indexing: 1 = motif (CpG) site, for the cluster second pass
  int64_t n_cells = 0;
  char base = 'C';
  int* overflow_d = nullptr;                     // set by k_accumulate when a coverage counter reaches DM_CELL_MASK
  void* nccl_comm = nullptr; int nccl_rank = 0, nccl_ranks = 0;   // dm_reduce_comm's communicator (kept across calls)
  float reduce_ms = 0.f;                         // device time of the last dm_reduce / dm_reduce_comm exchange
  // scratch
  void* scratch = nullptr; size_t scratch_bytes = 0;
  void* hbuf = nullptr; size_t hbuf_bytes = 0;   // inter-layer hidden states (tensor-core path)
  float* dpart = nullptr; size_t dpart_bytes = 0;
  void* pinned = nullptr; size_t pinned_bytes = 0;
  int64_t launches = 0;
  float lstm_ms = 0.f, total_ms = 0.f;
  bool fp32_attr_set = false, tc_attr_set = false;
  bool tc_pair = true;            // CTA-pair (cta_group::2) variant of the tensor-core kernel
  bool tc_f16 = false;            // operand format of the tensor-core path: fp16 (DM_F16) instead of bf16
  std::string err;
};

// error helpers -----------------------------------------------------------------
void dm_set_error(dm_ctx* ctx, const std::string& msg);
#define DM_CUDA(ctx, call)                                                              \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      dm_set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(e__));           \
      return DM_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// kernels / launchers (each returns a dm_status) ---------------------------------
int dm_batch_reserve(dm_ctx* ctx, int64_t n_reads, int64_t n_events, int64_t n_cols, int64_t n_windows);  // dm_api.cu
int dm_genome_sequence_upload(dm_ctx* ctx, int32_t contig, const uint8_t* seq, int64_t len);               // dm_align.cu
int dm_align_build(dm_ctx* ctx, const dm_sam_batch* sb, int64_t* n_windows_out, int64_t* n_cols_out);      // dm_align.cu
int dm_signal_event_stats(dm_ctx* ctx, int32_t n_reads, const int64_t* raw_off, const int16_t* raw, const int64_t* ev_off,
                          const int64_t* ev_start, const int64_t* ev_length, float* mean_out, float* stdv_out);   // dm_signal.cu
int dm_launch_prepare(dm_ctx* ctx);                       // dm_features.cu
int dm_launch_build_windows(dm_ctx* ctx, float* out_d);   // dm_features.cu
int dm_launch_accumulate(dm_ctx* ctx);                    // dm_hist.cu
int dm_check_overflow(dm_ctx* ctx);                       // dm_hist.cu: DM_ERR_OVERFLOW if a counter hit its limit (syncs the stream)
int dm_hist_max_cov(dm_ctx* ctx, unsigned long long* max_cov);   // dm_hist.cu
int dm_hist_normalise_flags(dm_ctx* ctx);                 // dm_hist.cu: key-created field back to 0/1 after a sum
void dm_reduce_release(dm_ctx* ctx);                      // dm_reduce.cu: destroys the context's communicator
int dm_launch_mask_rejected(dm_ctx* ctx);                 // dm_hist.cu
struct dm_cluster_result {                                 // output of the cluster second pass (dm_cluster.cu)
  std::vector<int64_t> pos;
  std::vector<int8_t> strand;
  std::vector<int32_t> cov, mod, pct;
  std::vector<float> prob, feat;
};
int dm_cluster_run(dm_ctx* ctx, int32_t contig, const dm_cluster_weights* cw, int drop_unmodified, bool want_feat,
                   dm_cluster_result& out);
int dm_cluster_sites_upload(dm_ctx* ctx, int32_t contig, int64_t n, const int64_t* pos, const int8_t* strand);
int dm_hist_load_rows(dm_ctx* ctx, int32_t contig, int8_t strand, int64_t n, const int64_t* pos, const int32_t* cov,
                      const int32_t* mod);
int dm_hist_compact(dm_ctx* ctx, int32_t contig, int8_t strand, std::vector<int64_t>& pos,
                    std::vector<int32_t>& cov, std::vector<int32_t>& mod);  // dm_hist.cu
// BiLSTM over the uploaded batch's feature table -> b.p1 / b.pred
int dm_launch_lstm_fp32(dm_ctx* ctx, const float* feat, const int32_t* win_frow, int64_t n_windows,
                        float* p1, uint8_t* pred);        // dm_lstm_fp32.cu
int dm_launch_lstm_tc(dm_ctx* ctx, const uint16_t* feat_tc, const int32_t* win_frow,
                      int64_t n_windows, float* p1, uint8_t* pred);   // dm_lstm_tc.cu
int dm_tc_selftest(dm_ctx* ctx, int n, int k, float* max_err);       // dm_lstm_tc.cu
int dm_tc_debug(dm_ctx* ctx, const uint16_t* feat_tc, const int32_t* win_frow, int64_t n_windows,
                float* p1, uint8_t* pred, int max_steps, unsigned char* dump_host, int64_t dump_cap);
// window-level entry: explicit [n,21,7] windows -> feature rows (21 per window)
int dm_launch_windows_to_rows(dm_ctx* ctx, const float* X_d, int64_t n, float* feat,
                              uint16_t* feat_tc, int32_t* win_frow);   // dm_features.cu

// result buffers are padded to whole CTA pairs (2 x 128 windows)
static inline int64_t dm_pad_windows(int64_t n) { return (n + 2 * DM_TILE_M - 1) / (2 * DM_TILE_M) * (2 * DM_TILE_M); }
