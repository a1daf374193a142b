// SAM/CIGAR walk on the GPU: alignment records -> the packed alignment columns (`base_map_info`) of the
// resident read batch.
//
// Reference behaviour restated (bin/DeepMod_scripts/myDetect.py, handle_record):
//   :565-621  CIGAR expansion: M/=/X columns (ref, read), I ('-', read), D/N (ref, '-'); S inside the alignment
//             consumes a read base without a column; first/last matching column / read index / position
//   :622-627  no matching base at all -> the read is dropped
//   :630-657  clips grow by the unmatched ends; base_map_info is cut to [first_al_match, last_al_match]
//   :661-666  '-' strand: reverse, complement, swap the clips
//   :680-700  CpG gap swap ("C-G" written as C,gap..,G with the read's G moved next to the C)
//   :702-705  fewer than 50 events left -> 'Less Event'
// The host (deepmod_b200/sam.py) only tokenises the SAM text: best-MAPQ record per read (handle_line :929-943),
// CIGAR -> (op, length) arrays, and the removal of leading/trailing non-aligned ops (:527-540).
#include "dm_common.cuh"

#include <algorithm>
#include <numeric>

namespace {

__device__ __forceinline__ int find_seg(const int64_t* __restrict__ off, int64_t n, int64_t x) {
  int64_t lo = 0, hi = n;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (off[mid] <= x) lo = mid; else hi = mid;
  }
  return (int)lo;
}

__device__ __forceinline__ uint8_t complement(uint8_t b) {      // myCom.na_bp; anything else maps to itself
  switch (b) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return b;
  }
}

struct ReadAcc {            // per read, filled by k_expand with atomics
  int first_read, last_read;       // firstmatch / lastmatch (read index)
  int first_col, last_col;         // first_al_match / last_al_match (raw column index within the read)
};

// one thread per RAW alignment column
__global__ void k_expand(int64_t n_raw, int64_t n_ops, int n_reads, const int64_t* __restrict__ op_col /*[n_ops+1]*/,
                         const int64_t* __restrict__ op_off /*[n_reads+1]*/, const uint8_t* __restrict__ op_code,
                         const int64_t* __restrict__ op_read /*read index at op start, per read*/,
                         const int64_t* __restrict__ op_ref /*ref offset at op start, per read*/,
                         const int32_t* __restrict__ contig, const int64_t* __restrict__ ref_start,
                         const int64_t* __restrict__ seq_off, const uint8_t* __restrict__ seq,
                         const uint8_t* __restrict__ genome, const int64_t* __restrict__ contig_off, int n_contigs,
                         uint8_t* __restrict__ raw_ref, uint8_t* __restrict__ raw_read, int64_t* __restrict__ raw_pos,
                         ReadAcc* __restrict__ acc) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_raw) return;
  const int o = find_seg(op_col, n_ops, c);
  const int r = find_seg(op_off, n_reads, o);
  const int64_t k = c - op_col[o];
  const uint8_t op = op_code[o];
  const bool adv_ref = op == 'M' || op == 'D' || op == 'N' || op == '=' || op == 'X';
  const bool adv_read = op == 'M' || op == 'I' || op == '=' || op == 'X';
  const int64_t rp = ref_start[r] + op_ref[o] + (adv_ref ? k : 0);
  const int64_t ri = op_read[o] + (adv_read ? k : 0);
  uint8_t rb = '-', qb = '-';
  if (op != 'I') {
    const int ct = contig[r];
    const int64_t len = (ct >= 0 && ct < n_contigs) ? contig_off[ct + 1] - contig_off[ct] : 0;
    rb = (rp >= 0 && rp < len) ? genome[contig_off[ct] + rp] : 'N';
  }
  if (adv_read) qb = seq[seq_off[r] + ri];
  raw_ref[c] = rb;
  raw_read[c] = qb;
  raw_pos[c] = rp;
  const bool match = op == '=' || (op == 'M' && rb == qb);
  if (match) {
    // columns are ordered inside a read, so only the first / last matching lane of a same-read run in the warp
    // can improve the read's minimum / maximum
    const int col = (int)(c - op_col[op_off[r]]);
    const unsigned peers = __match_any_sync(__activemask(), r);
    const int lane = threadIdx.x & 31;
    if (lane == __ffs(peers) - 1) { atomicMin(&acc[r].first_read, (int)ri); atomicMin(&acc[r].first_col, col); }
    if (lane == 31 - __clz(peers)) { atomicMax(&acc[r].last_read, (int)ri); atomicMax(&acc[r].last_col, col); }
  }
}

// one thread per read: final clips, kept column range, status (:622-657, :661-666, :702-705)
__global__ void k_read_ranges(int n_reads, const ReadAcc* __restrict__ acc, const int64_t* __restrict__ ev_off,
                              const int64_t* __restrict__ op_off, const int64_t* __restrict__ op_col,
                              const int8_t* __restrict__ strand, const int32_t* __restrict__ clip_left,
                              const int32_t* __restrict__ clip_right, int32_t* __restrict__ start_clip,
                              int32_t* __restrict__ end_clip, int32_t* __restrict__ col_lo, int32_t* __restrict__ n_keep,
                              int32_t* __restrict__ n_win, int32_t* __restrict__ align_status) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const int n_cols = (int)(op_col[op_off[r + 1]] - op_col[op_off[r]]);
  const ReadAcc a = acc[r];
  int left = clip_left[r], right = clip_right[r];
  const int L = (int)(ev_off[r + 1] - ev_off[r]);
  const int n_ev = L - left - right;                       // len(m_event) after the clip removal
  if (a.first_read == INT32_MAX || a.last_read < 0) {      // no matching base
    start_clip[r] = L; end_clip[r] = 0; col_lo[r] = 0; n_keep[r] = 0; n_win[r] = 0;
    align_status[r] = DM_READ_NO_MATCH;
    return;
  }
  const int firstmatch = a.first_read, lastmatch = a.last_read;
  const int tail = n_ev - lastmatch > 1 ? n_ev - lastmatch - 1 : 0;
  if (strand[r] >= 0) { left += firstmatch; right += tail; }
  else                { right += firstmatch; left += tail; }
  int lo = 0, keep = n_cols;
  if (firstmatch > 0 || n_cols - a.last_col > 1) {
    if (n_cols - a.last_col > 1) { lo = a.first_col; keep = a.last_col + 1 - a.first_col; }
    else if (a.first_col > 0)    { lo = a.first_col; keep = n_cols - a.first_col; }
  }
  if (strand[r] < 0) { const int t = left; left = right; right = t; }
  const int lmap = L - left - right;
  start_clip[r] = left;
  end_clip[r] = right;
  col_lo[r] = lo;
  n_keep[r] = keep;
  n_win[r] = lmap >= 50 ? lmap : 0;
  align_status[r] = DM_READ_OK;                            // 'Less Event' is recognised by the detect pass itself
}

// one thread per KEPT column: cut, reverse + complement for '-' reads
__global__ void k_emit_columns(int64_t n_out, int n_reads, const int64_t* __restrict__ col_off /*[n_reads+1] out*/,
                               const int64_t* __restrict__ op_off, const int64_t* __restrict__ op_col,
                               const int32_t* __restrict__ col_lo, const int8_t* __restrict__ strand,
                               const uint8_t* __restrict__ raw_ref, const uint8_t* __restrict__ raw_read,
                               const int64_t* __restrict__ raw_pos, uint8_t* __restrict__ out_ref,
                               uint8_t* __restrict__ out_read, int64_t* __restrict__ out_pos) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_out) return;
  const int r = find_seg(col_off, n_reads, c);
  const int64_t i = c - col_off[r], keep = col_off[r + 1] - col_off[r];
  const int64_t base = op_col[op_off[r]] + col_lo[r];
  if (strand[r] >= 0) {
    const int64_t s = base + i;
    out_ref[c] = raw_ref[s]; out_read[c] = raw_read[s]; out_pos[c] = raw_pos[s];
  } else {
    const int64_t s = base + (keep - 1 - i);
    out_ref[c] = complement(raw_ref[s]); out_read[c] = complement(raw_read[s]); out_pos[c] = raw_pos[s];
  }
}

// CpG gap swap (:680-700), one thread per column.  The reference walks the columns sequentially, but a swap of the
// first rule only rewrites G-reference columns behind a (C,C) column and a swap of the second rule only rewrites
// C-reference columns in front of a (G,G) column; neither can create, destroy or feed a trigger of the other kind or
// of a later one of its own kind, so every trigger can be decided on its own and the swaps never overlap.
__global__ void k_gap_swap(int64_t n_out, int n_reads, const int64_t* __restrict__ col_off, const uint8_t* __restrict__ refb,
                           uint8_t* __restrict__ readb) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_out) return;
  const uint8_t rfc = refb[c], rdc = readb[c];
  const bool t1 = rfc == 'C' && rdc == 'C', t2 = rfc == 'G' && rdc == 'G';
  if (!t1 && !t2) return;
  const int r = find_seg(col_off, n_reads, c);
  const int64_t b = col_off[r], n = col_off[r + 1] - b, ali = c - b;
  const uint8_t* rf = refb + b;
  uint8_t* rd = readb + b;
  if (t1) {
    if (ali + 1 < n && rd[ali + 1] == '-' && rf[ali + 1] == 'G') {
      int64_t add = 2;
      while (ali + add < n && rd[ali + add] == '-' && rf[ali + add] == 'G') ++add;
      if (ali + add < n && rd[ali + add] == 'G' && rf[ali + add] == 'G') {
        const uint8_t t = rd[ali + 1]; rd[ali + 1] = rd[ali + add]; rd[ali + add] = t;
      }
    }
  } else {
    if (ali - 1 > -1 && rd[ali - 1] == '-' && rf[ali - 1] == 'C') {
      int64_t add = 2;
      while (ali - add > -1 && rd[ali - add] == '-' && rf[ali - add] == 'C') ++add;
      if (ali - add > -1 && rd[ali - add] == 'C' && rf[ali - add] == 'C') {
        const uint8_t t = rd[ali - 1]; rd[ali - 1] = rd[ali - add]; rd[ali - add] = t;
      }
    }
  }
}

__global__ void k_init_acc(int n, ReadAcc* acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { acc[i].first_read = INT32_MAX; acc[i].last_read = -1; acc[i].first_col = INT32_MAX; acc[i].last_col = -1; }
}

inline unsigned nb(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

template <typename T>
struct Dev {                      // stream-ordered scratch (pool allocator: no device-wide synchronisation)
  T* p = nullptr;
  cudaStream_t st = nullptr;
  ~Dev() { if (p) cudaFreeAsync(p, st); }
  cudaError_t alloc(size_t n, cudaStream_t s = nullptr) {
    st = s;
    return cudaMallocAsync(reinterpret_cast<void**>(&p), sizeof(T) * std::max<size_t>(n, 1), s);
  }
  cudaError_t upload(const T* h, size_t n, cudaStream_t s) {
    cudaError_t e = alloc(n, s);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(p, h, sizeof(T) * n, cudaMemcpyHostToDevice, s);
    return e;
  }
};

}  // namespace

int dm_genome_sequence_upload(dm_ctx* ctx, int32_t contig, const uint8_t* seq, int64_t len) {
  if (!ctx->cells) { dm_set_error(ctx, "dm_set_genome not called"); return DM_ERR_STATE; }
  if (contig < 0 || contig >= ctx->n_contigs) { dm_set_error(ctx, "contig out of range"); return DM_ERR_ARG; }
  if (len != ctx->contig_len[contig]) { dm_set_error(ctx, "sequence length differs from the contig length given to dm_set_genome"); return DM_ERR_ARG; }
  if (!ctx->genome) {
    const int64_t total = ctx->contig_off[ctx->n_contigs];
    DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->genome), (size_t)std::max<int64_t>(total, 1)));
    DM_CUDA(ctx, cudaMemsetAsync(ctx->genome, 'N', (size_t)total, ctx->stream));
  }
  DM_CUDA(ctx, cudaMemcpyAsync(ctx->genome + ctx->contig_off[contig], seq, (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
  DM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DM_OK;
}

// Builds the resident batch (ctx->b) from SAM-level records.  Returns window / column totals through the pointers.
int dm_align_build(dm_ctx* ctx, const dm_sam_batch* sb, int64_t* n_windows_out, int64_t* n_cols_out) {
  dm_dev_batch& b = ctx->b;
  const int n = sb->n_reads;
  cudaStream_t s = ctx->stream;
  if (!ctx->genome) { dm_set_error(ctx, "dm_set_contig_sequence not called"); return DM_ERR_STATE; }
  b.n_reads = n; b.n_events = b.n_cols = b.n_windows = b.n_frows = 0; b.prepared = false;
  if (n_windows_out) *n_windows_out = 0;
  if (n_cols_out) *n_cols_out = 0;
  if (n == 0) return DM_OK;
  const int64_t n_ops = sb->op_off[n], n_events = sb->ev_off[n], n_seq = sb->seq_off[n];
  // op-level prefix sums on the host (ops are ~100x fewer than bases)
  std::vector<int64_t> op_col((size_t)n_ops + 1, 0), op_read((size_t)n_ops, 0), op_ref((size_t)n_ops, 0);
  for (int r = 0; r < n; ++r) {
    int64_t rd = 0, rf = 0;
    for (int64_t o = sb->op_off[r]; o < sb->op_off[r + 1]; ++o) {
      const uint8_t op = sb->op_code[o];
      const int64_t len = sb->op_len[o];
      if (len < 0) { dm_set_error(ctx, "negative CIGAR length"); return DM_ERR_ARG; }
      op_read[o] = rd; op_ref[o] = rf;
      const bool col = op == 'M' || op == 'I' || op == 'D' || op == 'N' || op == '=' || op == 'X';
      op_col[o + 1] = op_col[o] + (col ? len : 0);
      if (op == 'M' || op == 'I' || op == 'S' || op == '=' || op == 'X') rd += len;
      if (op == 'M' || op == 'D' || op == 'N' || op == '=' || op == 'X') rf += len;
    }
    if (rd != sb->seq_off[r + 1] - sb->seq_off[r]) {
      dm_set_error(ctx, "read " + std::to_string(r) + ": CIGAR consumes " + std::to_string(rd) + " bases, SEQ has " +
                            std::to_string(sb->seq_off[r + 1] - sb->seq_off[r]));
      return DM_ERR_ARG;
    }
  }
  const int64_t n_raw = op_col[n_ops];
  Dev<int64_t> d_op_col, d_op_off, d_op_read, d_op_ref, d_ref_start, d_seq_off, d_raw_pos;
  Dev<uint8_t> d_op_code, d_seq, d_raw_ref, d_raw_read;
  Dev<int32_t> d_clip_l, d_clip_r, d_col_lo, d_keep, d_nwin;
  Dev<ReadAcc> d_acc;
  DM_CUDA(ctx, d_op_col.upload(op_col.data(), op_col.size(), s));
  DM_CUDA(ctx, d_op_off.upload(sb->op_off, (size_t)n + 1, s));
  DM_CUDA(ctx, d_op_read.upload(op_read.data(), op_read.size(), s));
  DM_CUDA(ctx, d_op_ref.upload(op_ref.data(), op_ref.size(), s));
  DM_CUDA(ctx, d_op_code.upload(sb->op_code, (size_t)n_ops, s));
  DM_CUDA(ctx, d_ref_start.upload(sb->ref_start, (size_t)n, s));
  DM_CUDA(ctx, d_seq_off.upload(sb->seq_off, (size_t)n + 1, s));
  DM_CUDA(ctx, d_seq.upload(sb->seq, (size_t)n_seq, s));
  DM_CUDA(ctx, d_clip_l.upload(sb->clip_left, (size_t)n, s));
  DM_CUDA(ctx, d_clip_r.upload(sb->clip_right, (size_t)n, s));
  DM_CUDA(ctx, d_raw_ref.alloc(n_raw, s)); DM_CUDA(ctx, d_raw_read.alloc(n_raw, s)); DM_CUDA(ctx, d_raw_pos.alloc(n_raw, s));
  DM_CUDA(ctx, d_acc.alloc(n, s)); DM_CUDA(ctx, d_col_lo.alloc(n, s)); DM_CUDA(ctx, d_keep.alloc(n, s)); DM_CUDA(ctx, d_nwin.alloc(n, s));
  // per-read arrays of the resident batch
  int rc = dm_batch_reserve(ctx, n, n_events, 0, 0);
  if (rc != DM_OK) return rc;
  const auto H2D = cudaMemcpyHostToDevice;
  DM_CUDA(ctx, cudaMemcpyAsync(b.ev_off, sb->ev_off, sizeof(int64_t) * (n + 1), H2D, s));
  DM_CUDA(ctx, cudaMemcpyAsync(b.contig, sb->contig, sizeof(int32_t) * n, H2D, s));
  DM_CUDA(ctx, cudaMemcpyAsync(b.strand, sb->strand, n, H2D, s));
  if (n_events > 0) {
    DM_CUDA(ctx, cudaMemcpyAsync(b.ev_mean, sb->ev_mean, sizeof(float) * n_events, H2D, s));
    DM_CUDA(ctx, cudaMemcpyAsync(b.ev_stdv, sb->ev_stdv, sizeof(float) * n_events, H2D, s));
    DM_CUDA(ctx, cudaMemcpyAsync(b.ev_len, sb->ev_len, sizeof(float) * n_events, H2D, s));
    if (sb->ev_base) DM_CUDA(ctx, cudaMemcpyAsync(b.ev_base, sb->ev_base, (size_t)n_events, H2D, s));
  }
  b.has_ev_base = sb->ev_base != nullptr;
  DM_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
  k_init_acc<<<nb(n, 256), 256, 0, s>>>(n, d_acc.p);
  if (n_raw > 0)
    k_expand<<<nb(n_raw, 256), 256, 0, s>>>(n_raw, n_ops, n, d_op_col.p, d_op_off.p, d_op_code.p, d_op_read.p, d_op_ref.p,
                                            b.contig, d_ref_start.p, d_seq_off.p, d_seq.p, ctx->genome, ctx->contig_off_d,
                                            ctx->n_contigs, d_raw_ref.p, d_raw_read.p, d_raw_pos.p, d_acc.p);
  k_read_ranges<<<nb(n, 128), 128, 0, s>>>(n, d_acc.p, b.ev_off, d_op_off.p, d_op_col.p, b.strand, d_clip_l.p, d_clip_r.p,
                                           b.start_clip, b.end_clip, d_col_lo.p, d_keep.p, d_nwin.p, b.align_status);
  ctx->launches += 3;
  DM_CUDA(ctx, cudaGetLastError());
  std::vector<int32_t> keep((size_t)n), nwin((size_t)n);
  DM_CUDA(ctx, cudaMemcpyAsync(keep.data(), d_keep.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  DM_CUDA(ctx, cudaMemcpyAsync(nwin.data(), d_nwin.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  std::vector<int64_t> col_off((size_t)n + 1, 0), win_off((size_t)n + 1, 0);
  for (int r = 0; r < n; ++r) { col_off[r + 1] = col_off[r] + keep[r]; win_off[r + 1] = win_off[r] + nwin[r]; }
  const int64_t n_cols = col_off[n], n_windows = win_off[n], n_frows = n_windows + (int64_t)(2 * DM_FLANK) * n;
  if (n_frows + DM_WINDOW >= (int64_t)INT32_MAX) { dm_set_error(ctx, "dm_align_upload: batch too large (>2^31 rows)"); return DM_ERR_ARG; }
  rc = dm_batch_reserve(ctx, n, n_events, n_cols, n_windows);
  if (rc != DM_OK) return rc;
  DM_CUDA(ctx, cudaMemcpyAsync(b.col_off, col_off.data(), sizeof(int64_t) * (n + 1), H2D, s));
  DM_CUDA(ctx, cudaMemcpyAsync(b.win_off, win_off.data(), sizeof(int64_t) * (n + 1), H2D, s));
  if (n_cols > 0) {
    k_emit_columns<<<nb(n_cols, 256), 256, 0, s>>>(n_cols, n, b.col_off, d_op_off.p, d_op_col.p, d_col_lo.p, b.strand,
                                                   d_raw_ref.p, d_raw_read.p, d_raw_pos.p, b.col_refbase, b.col_readbase,
                                                   b.col_refpos);
    k_gap_swap<<<nb(n_cols, 256), 256, 0, s>>>(n_cols, n, b.col_off, b.col_refbase, b.col_readbase);
    ctx->launches += 2;
  }
  DM_CUDA(ctx, cudaGetLastError());
  DM_CUDA(ctx, cudaEventRecord(ctx->ev3, s));
  DM_CUDA(ctx, cudaStreamSynchronize(s));      // host vectors above are pageable
  DM_CUDA(ctx, cudaEventElapsedTime(&ctx->total_ms, ctx->ev0, ctx->ev3));   // walk kernels + the one size round trip
  ctx->lstm_ms = 0.f;
  b.n_events = n_events; b.n_cols = n_cols; b.n_windows = n_windows; b.n_frows = n_frows;
  b.from_alignment = true;
  if (n_windows_out) *n_windows_out = n_windows;
  if (n_cols_out) *n_cols_out = n_cols;
  return DM_OK;
}
