// Label write-back + per-position accumulation, fused.
//
// Reference behaviour (bin/DeepMod_scripts/myDetect.py):
//   mPredict1 :822-833  prediction m of a read belongs to its m-th non-gap alignment column;
//   sum_handler :1089-1100  for every column whose refbase == --Base: create the key
//       (chr, strand, refpos); if readbase != '-': cov += 1, and mod += 1 when mod_pred == 1.
// A key exists as soon as a column touches it, even a deletion, so the cell keeps a
// key-created flag next to cov and mod: row exists <=> cell != 0 (layout: dm_common.cuh).
#include "dm_common.cuh"

#include <algorithm>

namespace {

__device__ __forceinline__ int find_segment(const int64_t* __restrict__ off, int n, int64_t x) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void k_accumulate(int n_reads, int64_t n_cols, const int64_t* __restrict__ col_off,
                             const int64_t* __restrict__ col_rank, const uint8_t* __restrict__ refbase,
                             const uint8_t* __restrict__ readbase, const int64_t* __restrict__ refpos,
                             const int32_t* __restrict__ contig, const int8_t* __restrict__ strand,
                             const int32_t* __restrict__ status, const int64_t* __restrict__ win_off,
                             const uint8_t* __restrict__ pred, const int64_t* __restrict__ contig_off, int n_contigs,
                             unsigned long long* __restrict__ cells, uint8_t base, int* __restrict__ overflow) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  if (refbase[c] != base) return;                          // :1091 (base is one of A,C,G,T)
  int r = find_segment(col_off, n_reads, c);
  if (status[r] != DM_READ_OK) return;                     // read skipped at :712 / :702
  int ct = contig[r];
  if (ct < 0 || ct >= n_contigs) return;
  int64_t p = refpos[c];
  int64_t len = contig_off[ct + 1] - contig_off[ct];
  if (p < 0 || p >= len) return;
  // cells: per contig, [+ strand | - strand] blocks of contig length
  int64_t cell = 2 * contig_off[ct] + (strand[r] >= 0 ? 0 : len) + p;
  if (readbase[c] == '-') {
    atomicOr(&cells[cell], 1ull << DM_CELL_DEL_SHIFT);     // key created, nothing counted
  } else {
    int64_t k = col_rank[c] - col_rank[col_off[r]];
    unsigned long long add = 1ull << DM_CELL_COV_SHIFT;
    if (pred[win_off[r] + k] == 1) add |= 1ull << DM_CELL_MOD_SHIFT;
    const unsigned long long old = atomicAdd(&cells[cell], add);
    // the reference's python ints cannot overflow; ours report it (the carry went into the next field)
    if (((old >> DM_CELL_COV_SHIFT) & DM_CELL_MASK) == DM_CELL_MASK) *overflow = 1;
  }
}

// sum_handler's loop over STORED per-read records (myDetect.py:1089-1100, --predDet 0): same rules as k_accumulate,
// the prediction comes with the record instead of from the BiLSTM
__global__ void k_accumulate_records(int64_t n, const uint8_t* __restrict__ refbase, const uint8_t* __restrict__ readbase,
                                     const int64_t* __restrict__ refpos, const int8_t* __restrict__ mod_pred, int64_t len,
                                     unsigned long long* __restrict__ blk, uint8_t base, int* __restrict__ overflow) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || refbase[i] != base) return;
  const int64_t p = refpos[i];
  if (p < 0 || p >= len) return;
  if (readbase[i] == '-') {
    atomicOr(&blk[p], 1ull << DM_CELL_DEL_SHIFT);
  } else {
    unsigned long long add = 1ull << DM_CELL_COV_SHIFT;
    if (mod_pred[i] == 1) add |= 1ull << DM_CELL_MOD_SHIFT;
    const unsigned long long old = atomicAdd(&blk[p], add);
    if (((old >> DM_CELL_COV_SHIFT) & DM_CELL_MASK) == DM_CELL_MASK) *overflow = 1;
  }
}

// conservation counters of a cell range (dm_hist_totals) and the largest coverage (dm_reduce's pre-check)
__global__ void k_cell_totals(const unsigned long long* __restrict__ cells, int64_t n, unsigned long long* __restrict__ out) {
  unsigned long long cov = 0, mod = 0, rows = 0, chk = 0, mx = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long v = cells[i];
    if (v == 0ull) continue;
    const unsigned long long c = (v >> DM_CELL_COV_SHIFT) & DM_CELL_MASK, m = (v >> DM_CELL_MOD_SHIFT) & DM_CELL_MASK;
    cov += c; mod += m; rows += 1;
    chk += (c + 3ull * m) * (unsigned long long)(i % 65521 + 1);
    mx = c > mx ? c : mx;
  }
  for (int o = 16; o; o >>= 1) {
    cov += __shfl_xor_sync(0xffffffffu, cov, o);
    mod += __shfl_xor_sync(0xffffffffu, mod, o);
    rows += __shfl_xor_sync(0xffffffffu, rows, o);
    chk += __shfl_xor_sync(0xffffffffu, chk, o);
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = t > mx ? t : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], cov); atomicAdd(&out[1], mod); atomicAdd(&out[2], rows); atomicAdd(&out[3], chk);
    atomicMax(&out[4], mx);
  }
}

// dst += src for two accumulators of the same genome: counters add, key-created flags OR
__global__ void k_merge_cells(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, int64_t n,
                              int* __restrict__ overflow) {
  const unsigned long long low = (1ull << DM_CELL_DEL_SHIFT) - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long b = src[i];
    if (b == 0ull) continue;
    const unsigned long long a = dst[i];
    if (((a >> DM_CELL_COV_SHIFT) & DM_CELL_MASK) + ((b >> DM_CELL_COV_SHIFT) & DM_CELL_MASK) > DM_CELL_MASK) *overflow = 1;
    dst[i] = ((a & low) + (b & low)) | ((a | b) & ~low);
  }
}

// after a sum over ranks the key-created field holds the number of ranks that set it: back to 0/1
__global__ void k_normalise_flags(unsigned long long* __restrict__ cells, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long v = cells[i];
    if ((v >> DM_CELL_DEL_SHIFT) > 1ull) cells[i] = (v & ((1ull << DM_CELL_DEL_SHIFT) - 1)) | (1ull << DM_CELL_DEL_SHIFT);
  }
}

// Windows of reads rejected on the device (mismatch / bad alignment) report p1 = 0, pred = 0:
// the reference never predicts them (myDetect.py:712).  One CTA per read.
__global__ void k_mask_rejected(const int32_t* __restrict__ status, const int64_t* __restrict__ win_off,
                                float* __restrict__ p1, uint8_t* __restrict__ pred) {
  const int r = blockIdx.x;
  if (status[r] == DM_READ_OK) return;
  for (int64_t w = win_off[r] + threadIdx.x; w < win_off[r + 1]; w += blockDim.x) {
    p1[w] = 0.f;
    pred[w] = 0;
  }
}

// two-pass compaction of the non-zero cells of one (contig, strand) block
__global__ void k_count_nonzero(const unsigned long long* __restrict__ cells, int64_t n,
                                int* __restrict__ block_cnt) {
  __shared__ int wsum[8];
  int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) cnt += (i + j < n) && (cells[i + j] != 0ull);
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += wsum[w];
    block_cnt[blockIdx.x] = s;
  }
}

__global__ void k_scan_blocks(int* __restrict__ block_cnt, int n_blocks, int64_t* __restrict__ block_off,
                              int64_t* __restrict__ total) {
  // single thread block, sequential chunks of 1024
  __shared__ int64_t buf[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    int i = base + threadIdx.x;
    int64_t v = i < n_blocks ? block_cnt[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int64_t t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n_blocks) block_off[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void k_emit_nonzero(const unsigned long long* __restrict__ cells, int64_t n,
                               const int64_t* __restrict__ block_off, int64_t cap,
                               int64_t* __restrict__ pos, int32_t* __restrict__ cov, int32_t* __restrict__ mod) {
  __shared__ int wtot[8];
  int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  unsigned long long v[4];
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] = (i + j < n) ? cells[i + j] : 0ull;
    cnt += v[j] != 0ull;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wtot[warp] = incl;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += wtot[w];
  int64_t o = block_off[blockIdx.x] + wbase + incl - cnt;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (v[j] != 0ull) {
      if (o < cap) {
        pos[o] = i + j;
        cov[o] = (int32_t)((v[j] >> DM_CELL_COV_SHIFT) & DM_CELL_MASK);
        mod[o] = (int32_t)((v[j] >> DM_CELL_MOD_SHIFT) & DM_CELL_MASK);
      }
      ++o;
    }
  }
}

}  // namespace

int dm_launch_accumulate(dm_ctx* ctx) {
  dm_dev_batch& b = ctx->b;
  if (b.n_cols == 0 || ctx->cells == nullptr) return DM_OK;
  unsigned blocks = (unsigned)((b.n_cols + 255) / 256);
  k_accumulate<<<blocks, 256, 0, ctx->stream>>>(
      b.n_reads, b.n_cols, b.col_off, b.col_rank, b.col_refbase, b.col_readbase, b.col_refpos,
      b.contig, b.strand, b.status, b.win_off, b.pred, ctx->contig_off_d, ctx->n_contigs,
      ctx->cells, (uint8_t)ctx->base, ctx->overflow_d);
  ctx->launches += 1;
  DM_CUDA(ctx, cudaGetLastError());
  return DM_OK;
}

int dm_launch_mask_rejected(dm_ctx* ctx) {
  dm_dev_batch& b = ctx->b;
  if (b.n_reads == 0 || b.n_windows == 0) return DM_OK;
  k_mask_rejected<<<b.n_reads, 128, 0, ctx->stream>>>(b.status, b.win_off, b.p1, b.pred);
  ctx->launches += 1;
  DM_CUDA(ctx, cudaGetLastError());
  return DM_OK;
}

int dm_hist_compact(dm_ctx* ctx, int32_t contig, int8_t strand, std::vector<int64_t>& pos,
                    std::vector<int32_t>& cov, std::vector<int32_t>& mod) {
  pos.clear(); cov.clear(); mod.clear();
  if (ctx->cells == nullptr) { dm_set_error(ctx, "dm_set_genome not called"); return DM_ERR_STATE; }
  if (contig < 0 || contig >= ctx->n_contigs) { dm_set_error(ctx, "contig out of range"); return DM_ERR_ARG; }
  const int64_t len = ctx->contig_len[contig];
  if (len == 0) return DM_OK;
  const unsigned long long* blk = ctx->cells + 2 * ctx->contig_off[contig] + (strand >= 0 ? 0 : len);
  const int n_blocks = (int)((len + 1023) / 1024);
  int* block_cnt = nullptr;
  int64_t* block_off = nullptr;
  DM_CUDA(ctx, cudaMalloc(&block_cnt, sizeof(int) * (size_t)n_blocks));
  DM_CUDA(ctx, cudaMalloc(&block_off, sizeof(int64_t) * (size_t)(n_blocks + 1)));
  int64_t* total_d = block_off + n_blocks;
  k_count_nonzero<<<n_blocks, 256, 0, ctx->stream>>>(blk, len, block_cnt);
  k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(block_cnt, n_blocks, block_off, total_d);
  ctx->launches += 2;
  int64_t total = 0;
  cudaError_t e = cudaMemcpyAsync(&total, total_d, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  int rc = DM_OK;
  int64_t *pos_d = nullptr; int32_t *cov_d = nullptr, *mod_d = nullptr;
  if (e == cudaSuccess && total > 0) {
    e = cudaMalloc(&pos_d, sizeof(int64_t) * (size_t)total);
    if (e == cudaSuccess) e = cudaMalloc(&cov_d, sizeof(int32_t) * (size_t)total);
    if (e == cudaSuccess) e = cudaMalloc(&mod_d, sizeof(int32_t) * (size_t)total);
    if (e == cudaSuccess) {
      k_emit_nonzero<<<n_blocks, 256, 0, ctx->stream>>>(blk, len, block_off, total, pos_d, cov_d, mod_d);
      ctx->launches += 1;
      pos.resize(total); cov.resize(total); mod.resize(total);
      e = cudaMemcpyAsync(pos.data(), pos_d, sizeof(int64_t) * total, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(cov.data(), cov_d, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(mod.data(), mod_d, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
  }
  if (e != cudaSuccess) { dm_set_error(ctx, std::string("dm_hist_compact: ") + cudaGetErrorString(e)); rc = DM_ERR_CUDA; }
  cudaFree(pos_d); cudaFree(cov_d); cudaFree(mod_d); cudaFree(block_cnt); cudaFree(block_off);
  return rc;
}

int dm_check_overflow(dm_ctx* ctx) {
  if (!ctx->overflow_d) return DM_OK;
  int flag = 0;
  DM_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->overflow_d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (flag) {
    dm_set_error(ctx, "per-position coverage counter overflow (>= 2^28 reads over one position): the accumulator is not exact any more");
    return DM_ERR_OVERFLOW;
  }
  return DM_OK;
}

static int cell_totals(dm_ctx* ctx, unsigned long long out[5]) {
  if (ctx->cells == nullptr) { dm_set_error(ctx, "dm_set_genome not called"); return DM_ERR_STATE; }
  unsigned long long* out_d = nullptr;
  DM_CUDA(ctx, cudaMalloc(&out_d, 5 * sizeof(unsigned long long)));
  cudaMemsetAsync(out_d, 0, 5 * sizeof(unsigned long long), ctx->stream);
  const int grid = ctx->sm_count * 8;
  k_cell_totals<<<grid, 256, 0, ctx->stream>>>(ctx->cells, ctx->n_cells, out_d);
  ctx->launches += 1;
  cudaError_t e = cudaMemcpyAsync(out, out_d, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(out_d);
  if (e != cudaSuccess) { dm_set_error(ctx, std::string("cell totals: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  return DM_OK;
}

int dm_hist_max_cov(dm_ctx* ctx, unsigned long long* max_cov) {
  unsigned long long t[5];
  int rc = cell_totals(ctx, t);
  if (rc == DM_OK) *max_cov = t[4];
  return rc;
}

int dm_hist_normalise_flags(dm_ctx* ctx) {
  if (ctx->cells == nullptr) return DM_OK;
  k_normalise_flags<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->cells, ctx->n_cells);
  ctx->launches += 1;
  DM_CUDA(ctx, cudaGetLastError());
  return DM_OK;
}

extern "C" int dm_hist_totals(dm_ctx* ctx, uint64_t* sum_cov, uint64_t* sum_mod, uint64_t* n_rows, uint64_t* checksum) {
  if (!ctx) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  unsigned long long t[5];
  int rc = cell_totals(ctx, t);
  if (rc != DM_OK) return rc;
  if (sum_cov) *sum_cov = t[0];
  if (sum_mod) *sum_mod = t[1];
  if (n_rows) *n_rows = t[2];
  if (checksum) *checksum = t[3];
  return DM_OK;
}

extern "C" int dm_accumulate_records(dm_ctx* ctx, int32_t contig, int8_t strand, int64_t n, const uint8_t* refbase,
                                     const uint8_t* readbase, const int64_t* refpos, const int8_t* mod_pred) {
  if (!ctx || n < 0 || (n > 0 && (!refbase || !readbase || !refpos || !mod_pred))) return DM_ERR_ARG;
  if (!ctx->cells) { dm_set_error(ctx, "dm_accumulate_records: dm_set_genome not called"); return DM_ERR_STATE; }
  if (contig < 0 || contig >= ctx->n_contigs) { dm_set_error(ctx, "dm_accumulate_records: contig out of range"); return DM_ERR_ARG; }
  if (n == 0) return DM_OK;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int64_t len = ctx->contig_len[contig];
  unsigned long long* blk = ctx->cells + 2 * ctx->contig_off[contig] + (strand >= 0 ? 0 : len);
  uint8_t *rb = nullptr, *qb = nullptr; int64_t* rp = nullptr; int8_t* mp = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&rb), (size_t)n, s);
  if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&qb), (size_t)n, s);
  if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&rp), sizeof(int64_t) * (size_t)n, s);
  if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&mp), (size_t)n, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(rb, refbase, (size_t)n, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(qb, readbase, (size_t)n, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(rp, refpos, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(mp, mod_pred, (size_t)n, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    k_accumulate_records<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, rb, qb, rp, mp, len, blk, (uint8_t)ctx->base, ctx->overflow_d);
    ctx->launches += 1;
    e = cudaGetLastError();
  }
  if (rb) cudaFreeAsync(rb, s);
  if (qb) cudaFreeAsync(qb, s);
  if (rp) cudaFreeAsync(rp, s);
  if (mp) cudaFreeAsync(mp, s);
  if (e != cudaSuccess) { dm_set_error(ctx, std::string("dm_accumulate_records: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  return dm_check_overflow(ctx);
}

extern "C" int dm_hist_merge(dm_ctx* dst, dm_ctx* src) {
  if (!dst || !src || dst == src) return DM_ERR_ARG;
  if (!dst->cells || !src->cells) { dm_set_error(dst, "dm_hist_merge: dm_set_genome not called"); return DM_ERR_STATE; }
  if (dst->n_cells != src->n_cells || dst->base != src->base) { dm_set_error(dst, "dm_hist_merge: different genomes / bases"); return DM_ERR_ARG; }
  DM_CUDA(src, cudaSetDevice(src->device));
  DM_CUDA(src, cudaStreamSynchronize(src->stream));               // src's pending updates are in
  DM_CUDA(dst, cudaSetDevice(dst->device));
  const int grid = dst->sm_count * 8;
  if (dst->device == src->device) {
    k_merge_cells<<<grid, 256, 0, dst->stream>>>(dst->cells, src->cells, dst->n_cells, dst->overflow_d);
    dst->launches += 1;
  } else {
    // another GPU of the box: stage through a bounded buffer over NVLink (cudaMemcpyPeer works with or without P2P)
    const int64_t chunk = std::min<int64_t>(dst->n_cells, (int64_t)8 << 20);
    unsigned long long* tmp = nullptr;
    DM_CUDA(dst, cudaMalloc(&tmp, sizeof(unsigned long long) * (size_t)chunk));
    cudaError_t e = cudaSuccess;
    for (int64_t o = 0; o < dst->n_cells && e == cudaSuccess; o += chunk) {
      const int64_t m = std::min(chunk, dst->n_cells - o);
      e = cudaMemcpyPeerAsync(tmp, dst->device, src->cells + o, src->device, sizeof(unsigned long long) * (size_t)m, dst->stream);
      if (e == cudaSuccess) {
        k_merge_cells<<<grid, 256, 0, dst->stream>>>(dst->cells + o, tmp, m, dst->overflow_d);
        dst->launches += 1;
      }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(dst->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { dm_set_error(dst, std::string("dm_hist_merge: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  }
  DM_CUDA(dst, cudaGetLastError());
  return dm_check_overflow(dst);
}
