// C ABI of deepmod_b200 (include/deepmod_b200.h): context lifetime, weight packing, batch
// staging, the detect call and the per-position accumulator read-out / BED writer.
// Everything that computes runs in the CUDA kernels of this library; there is no CPU path.
#include "dm_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace {

thread_local std::string g_error;

template <typename T>
int grow(dm_ctx* ctx, T*& p, int64_t n) {
  if (p) cudaFree(p);
  p = nullptr;
  if (n <= 0) n = 1;
  DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&p), sizeof(T) * (size_t)n));
  return DM_OK;
}

#define DM_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != DM_OK) return rc__; \
  } while (0)

int fail(dm_ctx* ctx, int code, const std::string& msg) {
  dm_set_error(ctx, msg);
  return code;
}

// ---- weight images ------------------------------------------------------------------
// fp32 image: see dm_common.cuh (column permutation n' = j*100 + ug*4 + gate).
void pack_fp32(const float* kernel, const float* bias, int layer, std::vector<float>& W,
               std::vector<float>& B) {
  const int K = layer == 0 ? DM_K0_F32 : DM_K12_F32;
  W.assign((size_t)K * DM_GATES, 0.f);
  B.assign(DM_GATES, 0.f);
  for (int u = 0; u < DM_HIDDEN; ++u) {
    const int j = u / 25, ug = u % 25;
    for (int g = 0; g < 4; ++g) {
      const int n_ref = g * DM_HIDDEN + u, n_img = j * 100 + ug * 4 + g;
      B[n_img] = bias[n_ref];
      if (layer == 0) {
        for (int k = 0; k < DM_FNUM; ++k) W[(size_t)k * DM_GATES + n_img] = kernel[(size_t)k * DM_GATES + n_ref];
        for (int k = 0; k < DM_HIDDEN; ++k)
          W[(size_t)(8 + k) * DM_GATES + n_img] = kernel[(size_t)(DM_FNUM + k) * DM_GATES + n_ref];
      } else {
        for (int k = 0; k < 2 * DM_HIDDEN; ++k) W[(size_t)k * DM_GATES + n_img] = kernel[(size_t)k * DM_GATES + n_ref];
      }
    }
  }
}

}  // namespace

void dm_tc_pack_weights(const float* kernel, const float* bias, int layer, bool pair, bool f16, std::vector<uint16_t>& img);  // dm_lstm_tc.cu

static bool valid_precision(int p) { return p == DM_FP32 || p == DM_BF16 || p == DM_BF16_1CTA || p == DM_F16; }
static void apply_precision(dm_ctx* ctx, int p) {
  ctx->precision = p == DM_FP32 ? DM_FP32 : DM_BF16;      // internally: fp32 path or tensor-core path
  if (p != DM_FP32) { ctx->tc_pair = p != DM_BF16_1CTA; ctx->tc_f16 = p == DM_F16; }
}

void dm_set_error(dm_ctx* ctx, const std::string& msg) {
  g_error = msg;
  if (ctx) ctx->err = msg;
}

extern "C" {

int dm_version(void) { return 100; }

const char* dm_last_error(const dm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_error.c_str(); }

int dm_create(dm_ctx** out, int device, const dm_weights* w, int precision) {
  if (!out || !w) return fail(nullptr, DM_ERR_ARG, "dm_create: null argument");
  *out = nullptr;
  if (!valid_precision(precision)) return fail(nullptr, DM_ERR_ARG, "dm_create: bad precision");
  for (int d = 0; d < 2; ++d)
    for (int l = 0; l < 3; ++l)
      if (!w->kernel[d][l] || !w->bias[d][l]) return fail(nullptr, DM_ERR_ARG, "dm_create: missing weight tensor");
  if (!w->cls_w || !w->cls_b) return fail(nullptr, DM_ERR_ARG, "dm_create: missing classifier tensor");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, DM_ERR_CUDA, std::string("dm_create: no CUDA device (") + cudaGetErrorString(e) +
                                          "); deepmod_b200 has no CPU path");
  if (device < 0 || device >= n_dev) return fail(nullptr, DM_ERR_ARG, "dm_create: device out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, DM_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, DM_ERR_CUDA, "dm_create: this library is built for sm_100a (B200) only, found sm_" +
                                          std::to_string(prop.major) + std::to_string(prop.minor));
  dm_ctx* ctx = new dm_ctx();
  ctx->device = device;
  apply_precision(ctx, precision);
  ctx->sm_count = prop.multiProcessorCount;
  auto bail = [&](int rc) { std::string m = ctx->err; dm_destroy(ctx); g_error = m; return rc; };
#define DM_CK(call)                                                                         \
  do {                                                                                      \
    cudaError_t e2 = (call);                                                                \
    if (e2 != cudaSuccess) { dm_set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(e2)); return bail(DM_ERR_CUDA); } \
  } while (0)
  DM_CK(cudaSetDevice(device));
  DM_CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  DM_CK(cudaEventCreate(&ctx->ev0));
  DM_CK(cudaEventCreate(&ctx->ev1));
  DM_CK(cudaEventCreate(&ctx->ev2));
  DM_CK(cudaEventCreate(&ctx->ev3));
  std::vector<float> W, B;
  std::vector<uint16_t> T;
  for (int d = 0; d < 2; ++d)
    for (int l = 0; l < 3; ++l) {
      pack_fp32(w->kernel[d][l], w->bias[d][l], l, W, B);
      DM_CK(cudaMalloc(reinterpret_cast<void**>(&ctx->w.w32[d][l]), W.size() * sizeof(float)));
      DM_CK(cudaMalloc(reinterpret_cast<void**>(&ctx->w.b32[d][l]), B.size() * sizeof(float)));
      DM_CK(cudaMemcpy(ctx->w.w32[d][l], W.data(), W.size() * sizeof(float), cudaMemcpyHostToDevice));
      DM_CK(cudaMemcpy(ctx->w.b32[d][l], B.data(), B.size() * sizeof(float), cudaMemcpyHostToDevice));
      for (int f = 0; f < 2; ++f) {
        dm_tc_pack_weights(w->kernel[d][l], w->bias[d][l], l, false, f == 1, T);
        DM_CK(cudaMalloc(reinterpret_cast<void**>(&ctx->w.wtc[f][d][l]), T.size() * sizeof(uint16_t)));
        DM_CK(cudaMemcpy(ctx->w.wtc[f][d][l], T.data(), T.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        dm_tc_pack_weights(w->kernel[d][l], w->bias[d][l], l, true, f == 1, T);
        DM_CK(cudaMalloc(reinterpret_cast<void**>(&ctx->w.wtc2[f][d][l]), T.size() * sizeof(uint16_t)));
        DM_CK(cudaMemcpy(ctx->w.wtc2[f][d][l], T.data(), T.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
      }
    }
  DM_CK(cudaMalloc(reinterpret_cast<void**>(&ctx->w.cls_w), 2 * DM_HIDDEN * 2 * sizeof(float)));
  DM_CK(cudaMalloc(reinterpret_cast<void**>(&ctx->w.cls_b), 2 * sizeof(float)));
  DM_CK(cudaMemcpy(ctx->w.cls_w, w->cls_w, 2 * DM_HIDDEN * 2 * sizeof(float), cudaMemcpyHostToDevice));
  DM_CK(cudaMemcpy(ctx->w.cls_b, w->cls_b, 2 * sizeof(float), cudaMemcpyHostToDevice));
  float cd[2 * DM_HIDDEN];
  for (int i = 0; i < 2 * DM_HIDDEN; ++i) cd[i] = w->cls_w[i * 2 + 1] - w->cls_w[i * 2 + 0];
  DM_CK(cudaMalloc(reinterpret_cast<void**>(&ctx->w.cls_d), sizeof(cd)));
  DM_CK(cudaMemcpy(ctx->w.cls_d, cd, sizeof(cd), cudaMemcpyHostToDevice));
  ctx->w.cls_db = w->cls_b[1] - w->cls_b[0];
#undef DM_CK
  *out = ctx;
  return DM_OK;
}

void dm_destroy(dm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int d = 0; d < 2; ++d)
    for (int l = 0; l < 3; ++l) {
      cudaFree(ctx->w.w32[d][l]);
      cudaFree(ctx->w.b32[d][l]);
      for (int f = 0; f < 2; ++f) { cudaFree(ctx->w.wtc[f][d][l]); cudaFree(ctx->w.wtc2[f][d][l]); }
    }
  cudaFree(ctx->w.cls_w); cudaFree(ctx->w.cls_b); cudaFree(ctx->w.cls_d);
  if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
  dm_dev_batch* bs[3] = {&ctx->b, &ctx->b2, &ctx->fw};
  for (dm_dev_batch* b : bs) {
    cudaFree(b->ev_off); cudaFree(b->col_off); cudaFree(b->col_refpos);
    cudaFree(b->ev_mean); cudaFree(b->ev_stdv); cudaFree(b->ev_len);
    cudaFree(b->ev_base); cudaFree(b->col_refbase); cudaFree(b->col_readbase);
    cudaFree(b->start_clip); cudaFree(b->end_clip); cudaFree(b->contig); cudaFree(b->strand);
    cudaFree(b->win_off); cudaFree(b->col_rank); cudaFree(b->win_col); cudaFree(b->win_frow);
    cudaFree(b->status); cudaFree(b->align_status); cudaFree(b->feat); cudaFree(b->feat_tc); cudaFree(b->p1); cudaFree(b->pred);
  }
  cudaFree(ctx->scratch2);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  cudaFree(ctx->fw_x);
  dm_reduce_release(ctx);
  cudaFree(ctx->contig_off_d); cudaFree(ctx->cells); cudaFree(ctx->motif); cudaFree(ctx->genome); cudaFree(ctx->overflow_d);
  cudaFree(ctx->scratch); cudaFree(ctx->hbuf); cudaFree(ctx->dpart);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev2) cudaEventDestroy(ctx->ev2);
  if (ctx->ev3) cudaEventDestroy(ctx->ev3);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int dm_set_precision(dm_ctx* ctx, int precision) {
  if (!ctx) return DM_ERR_ARG;
  if (!valid_precision(precision)) return fail(ctx, DM_ERR_ARG, "dm_set_precision: bad precision");
  apply_precision(ctx, precision);
  return DM_OK;
}

int64_t dm_launch_count(const dm_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dm_last_timing(const dm_ctx* ctx, float* lstm_ms, float* total_ms) {
  if (!ctx) return DM_ERR_ARG;
  if (lstm_ms) *lstm_ms = ctx->lstm_ms;
  if (total_ms) *total_ms = ctx->total_ms;
  return DM_OK;
}

}  // extern "C"

// ---- internals shared by the entry points -------------------------------------------
namespace {

int run_lstm(dm_ctx* ctx, dm_dev_batch& b) {
  if (b.n_windows == 0) return DM_OK;
  if (ctx->precision == DM_FP32) return dm_launch_lstm_fp32(ctx, b.feat, b.win_frow, b.n_windows, b.p1, b.pred);
  return dm_launch_lstm_tc(ctx, b.feat_tc, b.win_frow, b.n_windows, b.p1, b.pred);
}

// result buffers + feature table for `n_windows` windows over `n_frows` feature rows
int reserve_windows(dm_ctx* ctx, dm_dev_batch& b, int64_t n_windows, int64_t n_frows) {
  const int64_t n_pad = dm_pad_windows(n_windows);
  if (n_pad > b.cap_windows) {
    const int64_t cap = n_pad + n_pad / 8;
    DM_TRY(grow(ctx, b.win_col, cap));
    DM_TRY(grow(ctx, b.win_frow, cap));
    DM_TRY(grow(ctx, b.p1, cap));
    DM_TRY(grow(ctx, b.pred, cap));
    b.cap_windows = cap;
  }
  if (n_frows + DM_WINDOW > b.cap_frows) {
    const int64_t cap = n_frows + n_frows / 8 + DM_WINDOW;
    DM_TRY(grow(ctx, b.feat, cap * DM_FEAT_STRIDE));
    DM_TRY(grow(ctx, b.feat_tc, cap * 16));
    b.cap_frows = cap;
  }
  return DM_OK;
}

}  // namespace

int dm_batch_reserve(dm_ctx* ctx, int64_t n, int64_t n_events, int64_t n_cols, int64_t n_windows) {
  dm_dev_batch& b = ctx->b;
  if (n > b.cap_reads) {
    const int64_t cap = n + n / 8 + 16;
    DM_TRY(grow(ctx, b.ev_off, cap + 1)); DM_TRY(grow(ctx, b.col_off, cap + 1)); DM_TRY(grow(ctx, b.win_off, cap + 1));
    DM_TRY(grow(ctx, b.start_clip, cap)); DM_TRY(grow(ctx, b.end_clip, cap)); DM_TRY(grow(ctx, b.contig, cap));
    DM_TRY(grow(ctx, b.strand, cap)); DM_TRY(grow(ctx, b.status, cap)); DM_TRY(grow(ctx, b.align_status, cap));
    b.cap_reads = cap;
  }
  if (n_events > b.cap_events) {
    const int64_t cap = n_events + n_events / 8;
    DM_TRY(grow(ctx, b.ev_mean, cap)); DM_TRY(grow(ctx, b.ev_stdv, cap)); DM_TRY(grow(ctx, b.ev_len, cap));
    DM_TRY(grow(ctx, b.ev_base, cap));
    b.cap_events = cap;
  }
  if (n_cols > b.cap_cols) {
    const int64_t cap = n_cols + n_cols / 8;
    DM_TRY(grow(ctx, b.col_refbase, cap)); DM_TRY(grow(ctx, b.col_readbase, cap)); DM_TRY(grow(ctx, b.col_refpos, cap));
    DM_TRY(grow(ctx, b.col_rank, cap));
    b.cap_cols = cap;
  }
  return reserve_windows(ctx, b, n_windows, n_windows + (int64_t)(2 * DM_FLANK) * n);
}

extern "C" {

int dm_event_stats(dm_ctx* ctx, int32_t n_reads, const int64_t* raw_off, const int16_t* raw, const int64_t* ev_off,
                   const int64_t* ev_start, const int64_t* ev_length, float* mean_out, float* stdv_out) {
  if (!ctx || n_reads < 0) return DM_ERR_ARG;
  if (n_reads > 0 && (!raw_off || !raw || !ev_off || !ev_start || !ev_length || !mean_out || !stdv_out))
    return fail(ctx, DM_ERR_ARG, "dm_event_stats: null array");
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  return dm_signal_event_stats(ctx, n_reads, raw_off, raw, ev_off, ev_start, ev_length, mean_out, stdv_out);
}

int dm_set_contig_sequence(dm_ctx* ctx, int32_t contig, const uint8_t* seq, int64_t len) {
  if (!ctx || !seq || len < 0) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  return dm_genome_sequence_upload(ctx, contig, seq, len);
}

int dm_align_upload(dm_ctx* ctx, const dm_sam_batch* sb, int64_t* n_windows, int64_t* n_cols) {
  if (!ctx || !sb) return DM_ERR_ARG;
  if (sb->n_reads < 0) return fail(ctx, DM_ERR_ARG, "dm_align_upload: negative n_reads");
  if (sb->n_reads > 0 && (!sb->ev_off || !sb->contig || !sb->strand || !sb->ref_start || !sb->clip_left || !sb->clip_right ||
                          !sb->op_off || !sb->op_code || !sb->op_len || !sb->seq_off || !sb->seq || !sb->ev_mean ||
                          !sb->ev_stdv || !sb->ev_len))
    return fail(ctx, DM_ERR_ARG, "dm_align_upload: null array");
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  return dm_align_build(ctx, sb, n_windows, n_cols);
}

int dm_fetch_alignment(dm_ctx* ctx, int64_t* col_off, uint8_t* refbase, uint8_t* readbase, int64_t* refpos,
                       int32_t* start_clip, int32_t* end_clip) {
  if (!ctx) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_dev_batch& b = ctx->b;
  cudaStream_t s = ctx->stream;
  const auto D2H = cudaMemcpyDeviceToHost;
  if (b.n_reads > 0) {
    if (col_off) DM_CUDA(ctx, cudaMemcpyAsync(col_off, b.col_off, sizeof(int64_t) * (b.n_reads + 1), D2H, s));
    if (start_clip) DM_CUDA(ctx, cudaMemcpyAsync(start_clip, b.start_clip, sizeof(int32_t) * b.n_reads, D2H, s));
    if (end_clip) DM_CUDA(ctx, cudaMemcpyAsync(end_clip, b.end_clip, sizeof(int32_t) * b.n_reads, D2H, s));
  }
  if (b.n_cols > 0) {
    if (refbase) DM_CUDA(ctx, cudaMemcpyAsync(refbase, b.col_refbase, b.n_cols, D2H, s));
    if (readbase) DM_CUDA(ctx, cudaMemcpyAsync(readbase, b.col_readbase, b.n_cols, D2H, s));
    if (refpos) DM_CUDA(ctx, cudaMemcpyAsync(refpos, b.col_refpos, sizeof(int64_t) * b.n_cols, D2H, s));
  }
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  return DM_OK;
}

int dm_batch_upload(dm_ctx* ctx, const dm_batch* hb, int64_t* n_windows_out) {
  if (!ctx || !hb) return DM_ERR_ARG;
  if (hb->n_reads < 0) return fail(ctx, DM_ERR_ARG, "dm_batch_upload: negative n_reads");
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_dev_batch& b = ctx->b;
  const int n = hb->n_reads;
  b.n_reads = n;
  b.n_events = b.n_cols = b.n_windows = b.n_frows = 0;
  if (n_windows_out) *n_windows_out = 0;
  if (n == 0) return DM_OK;
  if (!hb->ev_off || !hb->col_off || !hb->start_clip || !hb->end_clip || !hb->contig || !hb->strand)
    return fail(ctx, DM_ERR_ARG, "dm_batch_upload: null per-read array");
  if (hb->ev_off[0] != 0 || hb->col_off[0] != 0) return fail(ctx, DM_ERR_ARG, "dm_batch_upload: offsets must start at 0");
  // window offsets: a read contributes L - start_clip - end_clip windows, none if that is
  // below 50 ('Less Event', myDetect.py:702-705)
  std::vector<int64_t> win_off((size_t)n + 1, 0);
  for (int r = 0; r < n; ++r) {
    const int64_t L = hb->ev_off[r + 1] - hb->ev_off[r], C = hb->col_off[r + 1] - hb->col_off[r];
    if (L < 0 || C < 0 || hb->start_clip[r] < 0 || hb->end_clip[r] < 0)
      return fail(ctx, DM_ERR_ARG, "dm_batch_upload: negative length or clip in read " + std::to_string(r));
    const int64_t lmap = L - hb->start_clip[r] - hb->end_clip[r];
    win_off[r + 1] = win_off[r] + (lmap >= 50 ? lmap : 0);
  }
  const int64_t n_events = hb->ev_off[n], n_cols = hb->col_off[n], n_windows = win_off[n];
  const int64_t n_frows = n_windows + (int64_t)(2 * DM_FLANK) * n;
  if (n_frows + DM_WINDOW >= (int64_t)INT32_MAX) return fail(ctx, DM_ERR_ARG, "dm_batch_upload: batch too large (>2^31 rows)");
  if (n_events > 0 && (!hb->ev_mean || !hb->ev_stdv || !hb->ev_len)) return fail(ctx, DM_ERR_ARG, "dm_batch_upload: null event array");
  if (n_cols > 0 && (!hb->col_refbase || !hb->col_readbase || !hb->col_refpos)) return fail(ctx, DM_ERR_ARG, "dm_batch_upload: null column array");
  DM_TRY(dm_batch_reserve(ctx, n, n_events, n_cols, n_windows));
  b.from_alignment = false;
  cudaStream_t s = ctx->stream;
  const auto H2D = cudaMemcpyHostToDevice;
  DM_CUDA(ctx, cudaMemcpyAsync(b.ev_off, hb->ev_off, sizeof(int64_t) * (n + 1), H2D, s));
  DM_CUDA(ctx, cudaMemcpyAsync(b.col_off, hb->col_off, sizeof(int64_t) * (n + 1), H2D, s));
  DM_CUDA(ctx, cudaMemcpyAsync(b.win_off, win_off.data(), sizeof(int64_t) * (n + 1), H2D, s));
  DM_CUDA(ctx, cudaMemcpyAsync(b.start_clip, hb->start_clip, sizeof(int32_t) * n, H2D, s));
  DM_CUDA(ctx, cudaMemcpyAsync(b.end_clip, hb->end_clip, sizeof(int32_t) * n, H2D, s));
  DM_CUDA(ctx, cudaMemcpyAsync(b.contig, hb->contig, sizeof(int32_t) * n, H2D, s));
  DM_CUDA(ctx, cudaMemcpyAsync(b.strand, hb->strand, sizeof(int8_t) * n, H2D, s));
  if (n_events > 0) {
    DM_CUDA(ctx, cudaMemcpyAsync(b.ev_mean, hb->ev_mean, sizeof(float) * n_events, H2D, s));
    DM_CUDA(ctx, cudaMemcpyAsync(b.ev_stdv, hb->ev_stdv, sizeof(float) * n_events, H2D, s));
    DM_CUDA(ctx, cudaMemcpyAsync(b.ev_len, hb->ev_len, sizeof(float) * n_events, H2D, s));
    if (hb->ev_base) DM_CUDA(ctx, cudaMemcpyAsync(b.ev_base, hb->ev_base, n_events, H2D, s));
  }
  b.has_ev_base = hb->ev_base != nullptr;
  if (n_cols > 0) {
    DM_CUDA(ctx, cudaMemcpyAsync(b.col_refbase, hb->col_refbase, n_cols, H2D, s));
    DM_CUDA(ctx, cudaMemcpyAsync(b.col_readbase, hb->col_readbase, n_cols, H2D, s));
    DM_CUDA(ctx, cudaMemcpyAsync(b.col_refpos, hb->col_refpos, sizeof(int64_t) * n_cols, H2D, s));
  }
  // win_off.data() is pageable stack/heap memory: the copies above are complete (staged)
  // only after this synchronize
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  b.n_events = n_events; b.n_cols = n_cols; b.n_windows = n_windows; b.n_frows = n_frows;
  b.prepared = false;
  ctx->h2d_bytes = sizeof(int64_t) * 3 * (size_t)(n + 1) + (size_t)n * 13 + (size_t)n_events * (hb->ev_base ? 13 : 12) +
                   (size_t)n_cols * 10;
  if (n_windows_out) *n_windows_out = n_windows;
  return DM_OK;
}

int dm_detect_resident(dm_ctx* ctx, int accumulate) {
  if (!ctx) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_dev_batch& b = ctx->b;
  if (b.n_reads == 0) return DM_OK;
  if (accumulate && ctx->cells == nullptr) return fail(ctx, DM_ERR_STATE, "dm_detect_resident: dm_set_genome not called");
  cudaStream_t s = ctx->stream;
  DM_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
  DM_TRY(dm_launch_prepare(ctx));
  b.prepared = true;
  DM_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
  DM_TRY(run_lstm(ctx, b));
  DM_CUDA(ctx, cudaEventRecord(ctx->ev2, s));
  DM_TRY(dm_launch_mask_rejected(ctx));
  if (accumulate) DM_TRY(dm_launch_accumulate(ctx));
  DM_CUDA(ctx, cudaEventRecord(ctx->ev3, s));
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  DM_CUDA(ctx, cudaEventElapsedTime(&ctx->lstm_ms, ctx->ev1, ctx->ev2));
  DM_CUDA(ctx, cudaEventElapsedTime(&ctx->total_ms, ctx->ev0, ctx->ev3));
  return accumulate ? dm_check_overflow(ctx) : DM_OK;
}

int dm_fetch_results(dm_ctx* ctx, float* p1_out, uint8_t* pred_out, int32_t* status_out) {
  if (!ctx) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_dev_batch& b = ctx->b;
  cudaStream_t s = ctx->stream;
  const auto D2H = cudaMemcpyDeviceToHost;
  if (p1_out && b.n_windows > 0) DM_CUDA(ctx, cudaMemcpyAsync(p1_out, b.p1, sizeof(float) * b.n_windows, D2H, s));
  if (pred_out && b.n_windows > 0) DM_CUDA(ctx, cudaMemcpyAsync(pred_out, b.pred, b.n_windows, D2H, s));
  if (status_out && b.n_reads > 0) DM_CUDA(ctx, cudaMemcpyAsync(status_out, b.status, sizeof(int32_t) * b.n_reads, D2H, s));
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  if (status_out && b.n_reads > 0 && b.from_alignment) {     // a read dropped by the CIGAR walk keeps that verdict
    std::vector<int32_t> al((size_t)b.n_reads);
    DM_CUDA(ctx, cudaMemcpy(al.data(), b.align_status, sizeof(int32_t) * b.n_reads, D2H));
    for (int r = 0; r < b.n_reads; ++r)
      if (al[r] != DM_READ_OK) status_out[r] = al[r];
  }
  return DM_OK;
}

}  // extern "C"

namespace {

// Puts the second pipeline slot where every launcher looks (ctx->b, ctx->stream, ctx->scratch) for the
// lifetime of the object.
struct SlotSwap {
  dm_ctx* c;
  bool on;
  SlotSwap(dm_ctx* ctx, bool second) : c(ctx), on(second) { flip(); }
  ~SlotSwap() { flip(); }
  void flip() {
    if (!on) return;
    std::swap(c->b, c->b2);
    std::swap(c->stream, c->stream2);
    std::swap(c->scratch, c->scratch2);
    std::swap(c->scratch_bytes, c->scratch2_bytes);
  }
};

// dm_detect_batch for a large batch: contiguous read ranges balanced by events, alternating between two
// slots (device batch + stream + scratch each), so that the host->device copy of part k+1 and the
// device->host copy of part k-1 run under the kernels of part k.  Parts accumulate into the same
// per-position cells; results land at their window / read offsets in the caller's arrays.
int detect_batch_pipelined(dm_ctx* ctx, const dm_batch* hb, int parts, float* p1_out, uint8_t* pred_out,
                           int32_t* status_out) {
  const int n = hb->n_reads;
  if (!hb->ev_off || !hb->col_off || !hb->start_clip || !hb->end_clip || !hb->contig || !hb->strand)
    return fail(ctx, DM_ERR_ARG, "dm_detect_batch: null per-read array");
  if (hb->ev_off[0] != 0 || hb->col_off[0] != 0) return fail(ctx, DM_ERR_ARG, "dm_detect_batch: offsets must start at 0");
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->stream2) DM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
  const bool accumulate = ctx->cells != nullptr;
  // read ranges with about the same number of events
  std::vector<int> cut(1, 0);
  const int64_t total_ev = hb->ev_off[n];
  for (int k = 1, r = 0; k < parts; ++k) {
    const int64_t want = total_ev * k / parts;
    while (r < n && hb->ev_off[r] < want) ++r;
    if (r > cut.back() && r < n) cut.push_back(r);
  }
  cut.push_back(n);
  const int np = (int)cut.size() - 1;
  std::vector<cudaEvent_t> ev((size_t)np * 4, nullptr);
  struct Finish {          // whatever happens, nothing of this call is still in flight when it returns
    dm_ctx* c; std::vector<cudaEvent_t>& ev;
    ~Finish() {
      cudaStreamSynchronize(c->stream);
      if (c->stream2) cudaStreamSynchronize(c->stream2);
      for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
    }
  } finish{ctx, ev};
  for (auto& e : ev) DM_CUDA(ctx, cudaEventCreate(&e));
  int64_t win_base = 0;
  size_t h2d = 0;
  struct Part { int r0 = 0, m = 0; int64_t nw = 0, win0 = 0; };
  std::vector<Part> part((size_t)np);
  // results of part k go out on its own stream; issued one iteration late so that a blocking copy into
  // pageable host memory waits with the next part already queued
  auto fetch = [&](int k) -> int {
    const Part& pt = part[(size_t)k];
    SlotSwap slot(ctx, (k & 1) != 0);
    cudaStream_t s = ctx->stream;
    const dm_dev_batch& b = ctx->b;
    const auto D2H = cudaMemcpyDeviceToHost;
    if (p1_out && pt.nw > 0) DM_CUDA(ctx, cudaMemcpyAsync(p1_out + pt.win0, b.p1, sizeof(float) * pt.nw, D2H, s));
    if (pred_out && pt.nw > 0) DM_CUDA(ctx, cudaMemcpyAsync(pred_out + pt.win0, b.pred, pt.nw, D2H, s));
    if (status_out && pt.m > 0) DM_CUDA(ctx, cudaMemcpyAsync(status_out + pt.r0, b.status, sizeof(int32_t) * pt.m, D2H, s));
    return DM_OK;
  };
  for (int k = 0; k < np; ++k) {
    const int r0 = cut[k], r1 = cut[k + 1], m = r1 - r0;
    const int64_t e0 = hb->ev_off[r0], c0 = hb->col_off[r0];
    std::vector<int64_t> ev_off((size_t)m + 1), col_off((size_t)m + 1);
    for (int r = 0; r <= m; ++r) { ev_off[r] = hb->ev_off[r0 + r] - e0; col_off[r] = hb->col_off[r0 + r] - c0; }
    dm_batch sub{};
    sub.n_reads = m;
    sub.ev_off = ev_off.data(); sub.col_off = col_off.data();
    sub.ev_mean = hb->ev_mean ? hb->ev_mean + e0 : nullptr;
    sub.ev_stdv = hb->ev_stdv ? hb->ev_stdv + e0 : nullptr;
    sub.ev_len = hb->ev_len ? hb->ev_len + e0 : nullptr;
    sub.ev_base = hb->ev_base ? hb->ev_base + e0 : nullptr;
    sub.col_refbase = hb->col_refbase ? hb->col_refbase + c0 : nullptr;
    sub.col_readbase = hb->col_readbase ? hb->col_readbase + c0 : nullptr;
    sub.col_refpos = hb->col_refpos ? hb->col_refpos + c0 : nullptr;
    sub.start_clip = hb->start_clip + r0; sub.end_clip = hb->end_clip + r0;
    sub.contig = hb->contig + r0; sub.strand = hb->strand + r0;
    {
      SlotSwap slot(ctx, (k & 1) != 0);
      cudaStream_t s = ctx->stream;
      DM_CUDA(ctx, cudaStreamSynchronize(s));       // this slot's previous part is fetched: its buffers are free
      int64_t nw = 0;
      DM_TRY(dm_batch_upload(ctx, &sub, &nw));      // returns once the part is staged; the other slot keeps computing
      h2d += ctx->h2d_bytes;
      dm_dev_batch& b = ctx->b;
      DM_CUDA(ctx, cudaEventRecord(ev[4 * k + 0], s));
      DM_TRY(dm_launch_prepare(ctx));
      b.prepared = true;
      DM_CUDA(ctx, cudaEventRecord(ev[4 * k + 1], s));
      DM_TRY(run_lstm(ctx, b));
      DM_CUDA(ctx, cudaEventRecord(ev[4 * k + 2], s));
      DM_TRY(dm_launch_mask_rejected(ctx));
      if (accumulate) DM_TRY(dm_launch_accumulate(ctx));
      DM_CUDA(ctx, cudaEventRecord(ev[4 * k + 3], s));
      part[(size_t)k].r0 = r0; part[(size_t)k].m = m; part[(size_t)k].nw = nw; part[(size_t)k].win0 = win_base;
      win_base += nw;
    }
    if (k > 0) DM_TRY(fetch(k - 1));
  }
  DM_TRY(fetch(np - 1));
  DM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  DM_CUDA(ctx, cudaStreamSynchronize(ctx->stream2));
  ctx->lstm_ms = 0.f;
  for (int k = 0; k < np; ++k) {
    float ms = 0.f;
    DM_CUDA(ctx, cudaEventElapsedTime(&ms, ev[4 * k + 1], ev[4 * k + 2]));
    ctx->lstm_ms += ms;
  }
  DM_CUDA(ctx, cudaEventElapsedTime(&ctx->total_ms, ev[0], ev[4 * (np - 1) + 3]));
  ctx->h2d_bytes = h2d;
  return accumulate ? dm_check_overflow(ctx) : DM_OK;
}

}  // namespace

extern "C" {

// page-locked host memory for callers that stage their input files themselves (a loader thread reading
// file k+1 into one of these while dm_detect_batch works on file k): copies from it are truly asynchronous
int dm_pinned_alloc(size_t bytes, int device, void** out) {
  if (!out) return DM_ERR_ARG;
  *out = nullptr;
  // (the allocating thread needs a current device: name it, or a loader thread would open a context on device 0)
  cudaError_t e = device >= 0 ? cudaSetDevice(device) : cudaSuccess;
  if (e == cudaSuccess) e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) { dm_set_error(nullptr, std::string("dm_pinned_alloc: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  return DM_OK;
}

void dm_pinned_free(void* p) {
  if (p) cudaFreeHost(p);
}

int dm_set_pipeline(dm_ctx* ctx, int parts) {
  if (!ctx) return DM_ERR_ARG;
  if (parts < 0 || parts > 64) return fail(ctx, DM_ERR_ARG, "dm_set_pipeline: parts must be in [0, 64]");
  ctx->pipeline_parts = parts;
  return DM_OK;
}

int dm_detect_batch(dm_ctx* ctx, const dm_batch* hb, float* p1_out, uint8_t* pred_out, int32_t* status_out) {
  if (!ctx || !hb) return DM_ERR_ARG;
  // large batches are cut into sub-batches whose transfers hide under each other's kernels
  int parts = ctx->pipeline_parts;
  if (parts == 0) {
    const int64_t n_ev = (hb->n_reads > 0 && hb->ev_off) ? hb->ev_off[hb->n_reads] : 0;
    parts = (int)std::min<int64_t>(8, n_ev / 1500000);        // >= 1.5 M events (~15 ms of BiLSTM) per part
  }
  if (parts >= 2 && hb->n_reads >= 2) return detect_batch_pipelined(ctx, hb, parts, p1_out, pred_out, status_out);
  int64_t nw = 0;
  DM_TRY(dm_batch_upload(ctx, hb, &nw));
  DM_TRY(dm_detect_resident(ctx, ctx->cells != nullptr));
  return dm_fetch_results(ctx, p1_out, pred_out, status_out);
}

int dm_build_windows(dm_ctx* ctx, float* windows_out) {
  if (!ctx || !windows_out) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_dev_batch& b = ctx->b;
  if (b.n_reads == 0 || b.n_windows == 0) return DM_OK;
  if (!b.prepared) { DM_TRY(dm_launch_prepare(ctx)); b.prepared = true; }
  const size_t bytes = sizeof(float) * (size_t)b.n_windows * DM_WINDOW * DM_FNUM;
  float* tmp = nullptr;
  DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&tmp), bytes));
  int rc = dm_launch_build_windows(ctx, tmp);
  cudaError_t e = cudaSuccess;
  if (rc == DM_OK) {
    e = cudaMemcpyAsync(windows_out, tmp, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(ctx, DM_ERR_CUDA, std::string("dm_build_windows: ") + cudaGetErrorString(e));
  return rc;
}

int dm_forward_windows(dm_ctx* ctx, int64_t n, const float* X, float* p1_out, uint8_t* pred_out) {
  if (!ctx || (n > 0 && !X)) return DM_ERR_ARG;
  if (n < 0) return fail(ctx, DM_ERR_ARG, "dm_forward_windows: negative n");
  if (n == 0) return DM_OK;
  if ((n + 1) * DM_WINDOW >= (int64_t)INT32_MAX) return fail(ctx, DM_ERR_ARG, "dm_forward_windows: too many windows");
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_dev_batch& f = ctx->fw;
  DM_TRY(reserve_windows(ctx, f, n, n * DM_WINDOW));
  if (n > ctx->fw_x_cap) {
    DM_TRY(grow(ctx, ctx->fw_x, (n + n / 8) * DM_WINDOW * DM_FNUM));
    ctx->fw_x_cap = n + n / 8;
  }
  cudaStream_t s = ctx->stream;
  DM_CUDA(ctx, cudaMemcpyAsync(ctx->fw_x, X, sizeof(float) * (size_t)n * DM_WINDOW * DM_FNUM, cudaMemcpyHostToDevice, s));
  f.n_windows = n;
  f.n_frows = n * DM_WINDOW;
  DM_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
  DM_TRY(dm_launch_windows_to_rows(ctx, ctx->fw_x, n, f.feat, f.feat_tc, f.win_frow));
  DM_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
  DM_TRY(run_lstm(ctx, f));
  DM_CUDA(ctx, cudaEventRecord(ctx->ev2, s));
  if (p1_out) DM_CUDA(ctx, cudaMemcpyAsync(p1_out, f.p1, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
  if (pred_out) DM_CUDA(ctx, cudaMemcpyAsync(pred_out, f.pred, n, cudaMemcpyDeviceToHost, s));
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  DM_CUDA(ctx, cudaEventElapsedTime(&ctx->lstm_ms, ctx->ev1, ctx->ev2));
  DM_CUDA(ctx, cudaEventElapsedTime(&ctx->total_ms, ctx->ev0, ctx->ev2));
  return DM_OK;
}

// ---- accumulator -----------------------------------------------------------------------
int dm_set_genome(dm_ctx* ctx, int32_t n_contigs, const int64_t* contig_len, char base) {
  if (!ctx || n_contigs <= 0 || !contig_len) return DM_ERR_ARG;
  if (base != 'A' && base != 'C' && base != 'G' && base != 'T') return fail(ctx, DM_ERR_ARG, "dm_set_genome: base must be one of ACGT");
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> off((size_t)n_contigs + 1, 0);
  for (int i = 0; i < n_contigs; ++i) {
    if (contig_len[i] < 0) return fail(ctx, DM_ERR_ARG, "dm_set_genome: negative contig length");
    off[i + 1] = off[i] + contig_len[i];
  }
  cudaFree(ctx->cells); ctx->cells = nullptr;
  cudaFree(ctx->motif); ctx->motif = nullptr;
  cudaFree(ctx->genome); ctx->genome = nullptr;
  cudaFree(ctx->contig_off_d); ctx->contig_off_d = nullptr;
  ctx->n_cells = 2 * off[n_contigs];
  if (!ctx->overflow_d) DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->overflow_d), sizeof(int)));
  DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->cells), sizeof(unsigned long long) * (size_t)std::max<int64_t>(ctx->n_cells, 1)));
  DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->contig_off_d), sizeof(int64_t) * off.size()));
  DM_CUDA(ctx, cudaMemcpy(ctx->contig_off_d, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice));
  ctx->n_contigs = n_contigs;
  ctx->contig_len.assign(contig_len, contig_len + n_contigs);
  ctx->contig_off = off;
  ctx->base = base;
  return dm_hist_clear(ctx);
}

int dm_hist_clear(dm_ctx* ctx) {
  if (!ctx) return DM_ERR_ARG;
  if (!ctx->cells) return fail(ctx, DM_ERR_STATE, "dm_hist_clear: dm_set_genome not called");
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  DM_CUDA(ctx, cudaMemsetAsync(ctx->cells, 0, sizeof(unsigned long long) * (size_t)ctx->n_cells, ctx->stream));
  DM_CUDA(ctx, cudaMemsetAsync(ctx->overflow_d, 0, sizeof(int), ctx->stream));
  DM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DM_OK;
}

int dm_hist_device_ptr(dm_ctx* ctx, void** cells_d, int64_t* n_cells) {
  if (!ctx || !cells_d || !n_cells) return DM_ERR_ARG;
  if (!ctx->cells) return fail(ctx, DM_ERR_STATE, "dm_hist_device_ptr: dm_set_genome not called");
  *cells_d = ctx->cells;
  *n_cells = ctx->n_cells;
  return DM_OK;
}

int dm_hist_nonzero(dm_ctx* ctx, int32_t contig, int8_t strand, int64_t cap, int64_t* pos, int32_t* cov,
                    int32_t* mod, int64_t* n_rows) {
  if (!ctx || !n_rows) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> p; std::vector<int32_t> c, m;
  DM_TRY(dm_hist_compact(ctx, contig, strand, p, c, m));
  *n_rows = (int64_t)p.size();
  if (pos && cov && mod) {
    const int64_t k = std::min<int64_t>(cap, (int64_t)p.size());
    if (k > 0) {
      memcpy(pos, p.data(), sizeof(int64_t) * k);
      memcpy(cov, c.data(), sizeof(int32_t) * k);
      memcpy(mod, m.data(), sizeof(int32_t) * k);
    }
  }
  return DM_OK;
}

int dm_write_bed(dm_ctx* ctx, int32_t contig, int8_t strand, const char* chrom, const char* path, int64_t* n_rows) {
  if (!ctx || !chrom || !path) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> p; std::vector<int32_t> c, m;
  DM_TRY(dm_hist_compact(ctx, contig, strand, p, c, m));
  if (n_rows) *n_rows = (int64_t)p.size();
  if (p.empty()) return DM_OK;                    // myDetect.py:1109: no keys, no file
  FILE* fh = fopen(path, "w");
  if (!fh) return fail(ctx, DM_ERR_IO, std::string("dm_write_bed: cannot open ") + path);
  const char sc = strand >= 0 ? '+' : '-';
  // myDetect.py:1116-1120: ' '.join([chr, pos, pos+1, base, min(cov,1000), strand, pos, pos+1,
  //                                 '0,0,0', cov, '%d' % (100*mod/(cov or 1)), mod, '\n'])
  // formatted by hand into a 1 MB buffer: a whole-genome summary is ~10^8 rows, and fprintf is the bottleneck
  const size_t clen = strlen(chrom);
  std::vector<char> buf((size_t)1 << 20);
  size_t used = 0;
  auto put_num = [&](long long v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) buf[used++] = tmp[--n];
    buf[used++] = ' ';
  };
  bool ok = true;
  for (size_t i = 0; i < p.size() && ok; ++i) {
    if (used + clen + 256 > buf.size()) { ok = fwrite(buf.data(), 1, used, fh) == used; used = 0; }
    const long long pos = p[i], cov = c[i], mod = m[i];
    memcpy(&buf[used], chrom, clen); used += clen; buf[used++] = ' ';
    put_num(pos); put_num(pos + 1);
    buf[used++] = ctx->base; buf[used++] = ' ';
    put_num(cov > 1000 ? 1000LL : cov);
    buf[used++] = sc; buf[used++] = ' ';
    put_num(pos); put_num(pos + 1);
    memcpy(&buf[used], "0,0,0 ", 6); used += 6;
    put_num(cov); put_num((100 * mod) / (cov > 0 ? cov : 1)); put_num(mod);
    buf[used++] = '\n';
  }
  if (ok && used) ok = fwrite(buf.data(), 1, used, fh) == used;
  if (fclose(fh) != 0 || !ok) return fail(ctx, DM_ERR_IO, std::string("dm_write_bed: write failed for ") + path);
  return DM_OK;
}

int dm_debug_tc_windows(dm_ctx* ctx, int64_t n, const float* X, int max_steps, uint8_t* dump, int64_t dump_cap,
                        float* p1_out) {
  if (!ctx || n <= 0 || !X) return DM_ERR_ARG;
  if ((n + 1) * DM_WINDOW >= (int64_t)INT32_MAX) return fail(ctx, DM_ERR_ARG, "dm_debug_tc_windows: too many windows");
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_dev_batch& f = ctx->fw;
  DM_TRY(reserve_windows(ctx, f, n, n * DM_WINDOW));
  if (n > ctx->fw_x_cap) {
    DM_TRY(grow(ctx, ctx->fw_x, (n + n / 8) * DM_WINDOW * DM_FNUM));
    ctx->fw_x_cap = n + n / 8;
  }
  cudaStream_t s = ctx->stream;
  DM_CUDA(ctx, cudaMemcpyAsync(ctx->fw_x, X, sizeof(float) * (size_t)n * DM_WINDOW * DM_FNUM, cudaMemcpyHostToDevice, s));
  f.n_windows = n;
  f.n_frows = n * DM_WINDOW;
  DM_TRY(dm_launch_windows_to_rows(ctx, ctx->fw_x, n, f.feat, f.feat_tc, f.win_frow));
  DM_TRY(dm_tc_debug(ctx, f.feat_tc, f.win_frow, n, f.p1, f.pred, max_steps, dump, dump_cap));
  if (p1_out) DM_CUDA(ctx, cudaMemcpy(p1_out, f.p1, sizeof(float) * n, cudaMemcpyDeviceToHost));
  return DM_OK;
}

// ---- merged summary + cluster second pass ------------------------------------------------

int dm_hist_load(dm_ctx* ctx, int32_t contig, int8_t strand, int64_t n, const int64_t* pos, const int32_t* cov,
                 const int32_t* mod) {
  if (!ctx || (n > 0 && (!pos || !cov || !mod))) return DM_ERR_ARG;
  for (int64_t i = 0; i < n; ++i)       // the reference's ints are unbounded; ours say so instead of wrapping
    if (cov[i] < 0 || mod[i] < 0 || (uint64_t)cov[i] > DM_CELL_MASK || (uint64_t)mod[i] > DM_CELL_MASK)
      return fail(ctx, DM_ERR_OVERFLOW, "dm_hist_load: row " + std::to_string(i) + " has a count outside [0, 2^28)");
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  return dm_hist_load_rows(ctx, contig, strand, n, pos, cov, mod);
}

static void merged_line(FILE* fh, const char* chrom, long long pos, char base, char strand, long long cov, long long mod) {
  // sum_chr_mod.py:63 (two spaces after the strand; percentage = int(mod*100/cov))
  fprintf(fh, "%s %lld %lld %c %lld %c  %lld %lld 0,0,0 %lld %lld %lld", chrom, pos, pos + 1, base, cov < 1000 ? cov : 1000LL,
          strand, pos, pos + 1, cov, cov > 0 ? (mod * 100) / cov : 0LL, mod);
}

int dm_write_merged_bed(dm_ctx* ctx, int32_t contig, const char* chrom, const char* path, int64_t* n_rows) {
  if (!ctx || !chrom || !path) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> pp, pm; std::vector<int32_t> cp, mp, cm, mm;
  DM_TRY(dm_hist_compact(ctx, contig, 1, pp, cp, mp));
  DM_TRY(dm_hist_compact(ctx, contig, -1, pm, cm, mm));
  int64_t rows = 0;
  FILE* fh = nullptr;
  size_t i = 0, j = 0;
  while (i < pp.size() || j < pm.size()) {       // keys sorted by (pos, strand), '+' < '-'
    const bool plus = j >= pm.size() || (i < pp.size() && pp[i] <= pm[j]);
    const long long pos = plus ? pp[i] : pm[j], cov = plus ? cp[i] : cm[j], mod = plus ? mp[i] : mm[j];
    if (plus) ++i; else ++j;
    if (mod == 0) continue;                      // sum_chr_mod.py:55-57
    if (!fh) {
      fh = fopen(path, "w");
      if (!fh) return fail(ctx, DM_ERR_IO, std::string("dm_write_merged_bed: cannot open ") + path);
    }
    merged_line(fh, chrom, pos, ctx->base, plus ? '+' : '-', cov, mod);
    fputc('\n', fh);
    ++rows;
  }
  if (fh && fclose(fh) != 0) return fail(ctx, DM_ERR_IO, std::string("dm_write_merged_bed: write failed for ") + path);
  if (n_rows) *n_rows = rows;
  return DM_OK;
}

int dm_cluster_set_sites(dm_ctx* ctx, int32_t contig, int64_t n, const int64_t* pos, const int8_t* strand) {
  if (!ctx || n < 0 || (n > 0 && (!pos || !strand))) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  return dm_cluster_sites_upload(ctx, contig, n, pos, strand);
}

int dm_cluster_predict(dm_ctx* ctx, int32_t contig, const dm_cluster_weights* w, int drop_unmodified, int64_t cap,
                       int64_t* pos, int8_t* strand, int32_t* cov, int32_t* mod, float* features, float* prob,
                       int32_t* pct, int64_t* n_sites) {
  if (!ctx || !w || !n_sites || !w->w1 || !w->b1 || !w->w2 || !w->b2 || !w->wo || !w->bo) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_cluster_result r;
  DM_TRY(dm_cluster_run(ctx, contig, w, drop_unmodified, features != nullptr, r));
  *n_sites = (int64_t)r.pos.size();
  const size_t k = (size_t)std::min<int64_t>(cap, (int64_t)r.pos.size());
  if (k > 0) {
    if (pos) memcpy(pos, r.pos.data(), k * sizeof(int64_t));
    if (strand) memcpy(strand, r.strand.data(), k);
    if (cov) memcpy(cov, r.cov.data(), k * sizeof(int32_t));
    if (mod) memcpy(mod, r.mod.data(), k * sizeof(int32_t));
    if (features) memcpy(features, r.feat.data(), k * 14 * sizeof(float));
    if (prob) memcpy(prob, r.prob.data(), k * sizeof(float));
    if (pct) memcpy(pct, r.pct.data(), k * sizeof(int32_t));
  }
  return DM_OK;
}

int dm_write_cluster_bed(dm_ctx* ctx, int32_t contig, const dm_cluster_weights* w, int drop_unmodified, const char* chrom,
                         const char* path, int64_t* n_rows) {
  if (!ctx || !w || !chrom || !path) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_cluster_result r;
  DM_TRY(dm_cluster_run(ctx, contig, w, drop_unmodified, false, r));
  if (n_rows) *n_rows = (int64_t)r.pos.size();
  if (r.pos.empty()) return DM_OK;
  FILE* fh = fopen(path, "w");
  if (!fh) return fail(ctx, DM_ERR_IO, std::string("dm_write_cluster_bed: cannot open ") + path);
  for (size_t i = 0; i < r.pos.size(); ++i) {
    merged_line(fh, chrom, r.pos[i], ctx->base, r.strand[i] >= 0 ? '+' : '-', r.cov[i], r.mod[i]);
    fprintf(fh, " %d\n", r.pct[i]);
  }
  if (fclose(fh) != 0) return fail(ctx, DM_ERR_IO, std::string("dm_write_cluster_bed: write failed for ") + path);
  return DM_OK;
}

int dm_selftest_umma(dm_ctx* ctx, int n, int k, float* max_err) {
  if (!ctx || !max_err) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  return dm_tc_selftest(ctx, n, k, max_err);
}

}  // extern "C"
