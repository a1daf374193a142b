// CpG-cluster second pass on the per-position accumulator (BASELINE configs[4]).
//
// Reference behaviour restated on the GPU:
//   DeepMod_tools/sum_chr_mod.py:47-63      merged summary: (cov, mod) summed per (chr, pos, strand); rows with
//                                           mod == 0 are dropped; line format with two spaces after the strand
//   DeepMod_tools/hm_cluster_predict.py     :43-72  sites = merged rows that are motif (CpG) sites with cov > 0,
//                                                   methylation fraction = round(percent / 100, 3)
//                                           :128-154 14 features per site: own fraction, partner-strand fraction,
//                                                   number of other CpG sites within +-25 bp, 11-bin histogram of
//                                                   their fractions (bin = int(frac / 0.1 + 0.5)) normalised by
//                                                   the count and rounded to 3 decimals
//                                           :158-170 MLP 14 -> 100 (relu) -> 20 (relu) -> 1 (sigmoid), dropout is
//                                                   the identity at keep_prob = 1; appended column int(p * 100)
// The reference walks python dicts keyed by (chr, strand, pos); here the accumulator is dense by position, so the
// +-25 bp neighbourhood is 2 x 51 direct cell reads and the whole pass is one HBM-bound sweep.
#include "dm_common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int NB = 25;                 // nbsize, hm_cluster_predict.py:83
constexpr int NF = 14;
constexpr int RATIO_N = 52;            // at most 49 neighbours contribute

__device__ __forceinline__ bool cell_in_pred(unsigned long long v, uint8_t motif, int drop_unmodified) {
  const unsigned cov = (unsigned)((v >> DM_CELL_COV_SHIFT) & DM_CELL_MASK), mod = (unsigned)((v >> DM_CELL_MOD_SHIFT) & DM_CELL_MASK);
  return motif != 0 && cov > 0 && (!drop_unmodified || mod > 0);
}
__device__ __forceinline__ int cell_pct(unsigned long long v) {
  const unsigned cov = (unsigned)((v >> DM_CELL_COV_SHIFT) & DM_CELL_MASK), mod = (unsigned)((v >> DM_CELL_MOD_SHIFT) & DM_CELL_MASK);
  return (int)((100ull * mod) / cov);             // int(mod * 100 / cov), sum_chr_mod.py:63
}

__global__ void k_set_sites(int64_t n, const int64_t* __restrict__ pos, const int8_t* __restrict__ strand, int64_t len,
                            uint8_t* __restrict__ motif_blk) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = pos[i];
  if (p < 0 || p >= len) return;
  motif_blk[(strand[i] >= 0 ? 0 : len) + p] = 1;
}

__global__ void k_hist_load(int64_t n, const int64_t* __restrict__ pos, const int32_t* __restrict__ cov,
                            const int32_t* __restrict__ mod, int64_t len, unsigned long long* __restrict__ blk) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = pos[i];
  if (p < 0 || p >= len) return;
  blk[p] = ((unsigned long long)(cov[i] & DM_CELL_MASK) << DM_CELL_COV_SHIFT) |
           ((unsigned long long)(mod[i] & DM_CELL_MASK) << DM_CELL_MOD_SHIFT);
}

// flag[i] = 1 iff cell i of the contig's [+ | -] block is a row of hm_cluster_predict's preddict
__global__ void k_site_flags(const unsigned long long* __restrict__ cells, const uint8_t* __restrict__ motif, int64_t n,
                             int drop_unmodified, uint8_t* __restrict__ flag, int* __restrict__ block_cnt) {
  __shared__ int wsum[8];
  int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (i + j < n) {
      const uint8_t f = cell_in_pred(cells[i + j], motif[i + j], drop_unmodified) ? 1 : 0;
      flag[i + j] = f;
      cnt += f;
    }
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += wsum[w];
    block_cnt[blockIdx.x] = s;
  }
}

__global__ void k_scan_counts(const int* __restrict__ block_cnt, int n_blocks, int64_t* __restrict__ block_off,
                              int64_t* __restrict__ total) {
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

__global__ void k_emit_sites(const uint8_t* __restrict__ flag, int64_t n, const int64_t* __restrict__ block_off,
                             int64_t* __restrict__ site_idx) {
  __shared__ int wtot[8];
  int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  int f[4], cnt = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) { f[j] = (i + j < n) ? flag[i + j] : 0; cnt += f[j]; }
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
  for (int j = 0; j < 4; ++j)
    if (f[j]) site_idx[o++] = i + j;
}

struct ClusterW {
  float w1[NF * 100], b1[100], w2[100 * 20], b2[20], wo[20], bo;
};

// one thread per site: features (the reference's float64 arithmetic, exactly) -> fp32 MLP
__global__ void __launch_bounds__(128)
k_cluster_sites(const unsigned long long* __restrict__ cells, const uint8_t* __restrict__ flag, int64_t len,
                const int64_t* __restrict__ site_idx, int64_t n_sites, const ClusterW* __restrict__ wg,
                const float* __restrict__ ratio_lut, int64_t* __restrict__ pos_out, int8_t* __restrict__ strand_out,
                int32_t* __restrict__ cov_out, int32_t* __restrict__ mod_out, float* __restrict__ feat_out,
                float* __restrict__ prob_out, int32_t* __restrict__ pct_out) {
  __shared__ ClusterW w;
  for (int i = threadIdx.x; i < (int)(sizeof(ClusterW) / 4); i += blockDim.x)
    reinterpret_cast<float*>(&w)[i] = reinterpret_cast<const float*>(wg)[i];
  __syncthreads();
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_sites) return;
  const int64_t idx = site_idx[k];
  const int s = idx >= len ? 1 : 0;                 // 0: '+', 1: '-'
  const int64_t p = idx - (s ? len : 0);
  const unsigned long long v = cells[idx];
  const int64_t partner = s == 0 ? p + 1 : p - 1;   // opposite strand, hm_cluster_predict.py:132
  const int64_t pidx = (s == 0 ? len : 0) + partner;
  float x[NF];
  x[0] = (float)((double)cell_pct(v) / 100.0);
  x[1] = (partner >= 0 && partner < len && flag[pidx]) ? (float)((double)cell_pct(cells[pidx]) / 100.0) : 0.f;
  int hist[11];
#pragma unroll
  for (int i = 0; i < 11; ++i) hist[i] = 0;
  int total = 0;
  for (int64_t r = p - NB; r <= p + NB; ++r) {
    if (r == p || r == partner || r < 0 || r >= len) continue;
    int64_t j = -1;
    if (flag[r]) j = r;                             // '+' strand first, then '-' (the script's if / elif)
    else if (flag[len + r]) j = len + r;
    if (j < 0) continue;
    const double frac = (double)cell_pct(cells[j]) / 100.0;
    const int b = (int)(frac / 0.1 + 0.5);
#pragma unroll
    for (int i = 0; i < 11; ++i) hist[i] += (b == i);
    ++total;
  }
  x[2] = (float)total;
#pragma unroll
  for (int i = 0; i < 11; ++i) x[3 + i] = total > 0 ? ratio_lut[hist[i] * RATIO_N + total] : 0.f;
  // MLP: layer 1 unit by unit, folded straight into the 20 accumulators of layer 2
  float a2[20];
#pragma unroll
  for (int c = 0; c < 20; ++c) a2[c] = 0.f;
  for (int j = 0; j < 100; ++j) {
    float h = 0.f;
#pragma unroll
    for (int i = 0; i < NF; ++i) h = fmaf(x[i], w.w1[i * 100 + j], h);
    h = fmaxf(h + w.b1[j], 0.f);
#pragma unroll
    for (int c = 0; c < 20; ++c) a2[c] = fmaf(h, w.w2[j * 20 + c], a2[c]);
  }
  float z = 0.f;
#pragma unroll
  for (int c = 0; c < 20; ++c) z = fmaf(fmaxf(a2[c] + w.b2[c], 0.f), w.wo[c], z);
  z += w.bo;
  const float prob = 1.0f / (1.0f + expf(-z));
  pos_out[k] = p;
  strand_out[k] = s ? -1 : 1;
  cov_out[k] = (int32_t)((v >> DM_CELL_COV_SHIFT) & DM_CELL_MASK);
  mod_out[k] = (int32_t)((v >> DM_CELL_MOD_SHIFT) & DM_CELL_MASK);
  if (feat_out)
#pragma unroll
    for (int i = 0; i < NF; ++i) feat_out[k * NF + i] = x[i];
  prob_out[k] = prob;
  pct_out[k] = (int32_t)__fmul_rn(prob, 100.0f);    // int(float32(p) * 100), hm_cluster_predict.py:170
}

inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

int ensure_motif(dm_ctx* ctx) {
  if (!ctx->cells) { dm_set_error(ctx, "dm_set_genome not called"); return DM_ERR_STATE; }
  if (!ctx->motif) {
    DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->motif), (size_t)std::max<int64_t>(ctx->n_cells, 1)));
    DM_CUDA(ctx, cudaMemsetAsync(ctx->motif, 0, (size_t)ctx->n_cells, ctx->stream));
  }
  return DM_OK;
}

template <typename T>
struct DevBuf {                   // stream-ordered scratch (pool allocator: no device-wide synchronisation)
  T* p = nullptr;
  cudaStream_t st = nullptr;
  ~DevBuf() { if (p) cudaFreeAsync(p, st); }
  cudaError_t alloc(size_t n, cudaStream_t s) {
    st = s;
    return cudaMallocAsync(reinterpret_cast<void**>(&p), sizeof(T) * std::max<size_t>(n, 1), s);
  }
};

}  // namespace


int dm_cluster_run(dm_ctx* ctx, int32_t contig, const dm_cluster_weights* cw, int drop_unmodified, bool want_feat,
                   dm_cluster_result& out) {
  if (!ctx->cells) { dm_set_error(ctx, "dm_set_genome not called"); return DM_ERR_STATE; }
  if (contig < 0 || contig >= ctx->n_contigs) { dm_set_error(ctx, "contig out of range"); return DM_ERR_ARG; }
  if (!ctx->motif) { dm_set_error(ctx, "dm_cluster_set_sites not called"); return DM_ERR_STATE; }
  const int64_t len = ctx->contig_len[contig], n = 2 * len;
  out = dm_cluster_result();
  if (len == 0) return DM_OK;
  cudaStream_t s = ctx->stream;
  const unsigned long long* cells = ctx->cells + 2 * ctx->contig_off[contig];
  const uint8_t* motif = ctx->motif + 2 * ctx->contig_off[contig];
  const int n_blocks = (int)((n + 1023) / 1024);
  DevBuf<uint8_t> flag;
  DevBuf<int> block_cnt;
  DevBuf<int64_t> block_off, site_idx, pos_d;
  DevBuf<int8_t> strand_d;
  DevBuf<int32_t> cov_d, mod_d, pct_d;
  DevBuf<float> prob_d, feat_d, lut_d;
  DevBuf<ClusterW> w_d;
  DM_CUDA(ctx, flag.alloc(n, s));
  DM_CUDA(ctx, block_cnt.alloc(n_blocks, s));
  DM_CUDA(ctx, block_off.alloc(n_blocks + 1, s));
  DM_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
  k_site_flags<<<n_blocks, 256, 0, s>>>(cells, motif, n, drop_unmodified, flag.p, block_cnt.p);
  k_scan_counts<<<1, 1024, 0, s>>>(block_cnt.p, n_blocks, block_off.p, block_off.p + n_blocks);
  ctx->launches += 2;
  int64_t total = 0;
  DM_CUDA(ctx, cudaMemcpyAsync(&total, block_off.p + n_blocks, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  if (total == 0) return DM_OK;
  // host-built tables: round(i / float(t), 3) exactly as python rounds (correctly rounded decimal, then back)
  std::vector<float> lut((size_t)RATIO_N * RATIO_N, 0.f);
  for (int i = 0; i < RATIO_N; ++i)
    for (int t = 1; t < RATIO_N; ++t) {
      char tmp[64];
      snprintf(tmp, sizeof(tmp), "%.3f", (double)i / (double)t);
      lut[(size_t)i * RATIO_N + t] = (float)strtod(tmp, nullptr);
    }
  ClusterW hw;
  memcpy(hw.w1, cw->w1, sizeof(hw.w1)); memcpy(hw.b1, cw->b1, sizeof(hw.b1));
  memcpy(hw.w2, cw->w2, sizeof(hw.w2)); memcpy(hw.b2, cw->b2, sizeof(hw.b2));
  memcpy(hw.wo, cw->wo, sizeof(hw.wo)); hw.bo = cw->bo[0];
  DM_CUDA(ctx, lut_d.alloc(lut.size(), s));
  DM_CUDA(ctx, w_d.alloc(1, s));
  DM_CUDA(ctx, cudaMemcpyAsync(lut_d.p, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  DM_CUDA(ctx, cudaMemcpyAsync(w_d.p, &hw, sizeof(hw), cudaMemcpyHostToDevice, s));
  DM_CUDA(ctx, site_idx.alloc(total, s)); DM_CUDA(ctx, pos_d.alloc(total, s)); DM_CUDA(ctx, strand_d.alloc(total, s));
  DM_CUDA(ctx, cov_d.alloc(total, s)); DM_CUDA(ctx, mod_d.alloc(total, s)); DM_CUDA(ctx, pct_d.alloc(total, s));
  DM_CUDA(ctx, prob_d.alloc(total, s));
  if (want_feat) DM_CUDA(ctx, feat_d.alloc((size_t)total * NF, s));
  k_emit_sites<<<n_blocks, 256, 0, s>>>(flag.p, n, block_off.p, site_idx.p);
  k_cluster_sites<<<nblk(total, 128), 128, 0, s>>>(cells, flag.p, len, site_idx.p, total, w_d.p, lut_d.p, pos_d.p,
                                                   strand_d.p, cov_d.p, mod_d.p, want_feat ? feat_d.p : nullptr,
                                                   prob_d.p, pct_d.p);
  ctx->launches += 2;
  DM_CUDA(ctx, cudaGetLastError());
  DM_CUDA(ctx, cudaEventRecord(ctx->ev3, s));
  DM_CUDA(ctx, cudaEventSynchronize(ctx->ev3));
  DM_CUDA(ctx, cudaEventElapsedTime(&ctx->total_ms, ctx->ev0, ctx->ev3));   // flags + compaction + features + MLP
  ctx->lstm_ms = 0.f;
  out.pos.resize(total); out.strand.resize(total); out.cov.resize(total); out.mod.resize(total);
  out.pct.resize(total); out.prob.resize(total);
  const auto D2H = cudaMemcpyDeviceToHost;
  DM_CUDA(ctx, cudaMemcpyAsync(out.pos.data(), pos_d.p, sizeof(int64_t) * total, D2H, s));
  DM_CUDA(ctx, cudaMemcpyAsync(out.strand.data(), strand_d.p, total, D2H, s));
  DM_CUDA(ctx, cudaMemcpyAsync(out.cov.data(), cov_d.p, sizeof(int32_t) * total, D2H, s));
  DM_CUDA(ctx, cudaMemcpyAsync(out.mod.data(), mod_d.p, sizeof(int32_t) * total, D2H, s));
  DM_CUDA(ctx, cudaMemcpyAsync(out.pct.data(), pct_d.p, sizeof(int32_t) * total, D2H, s));
  DM_CUDA(ctx, cudaMemcpyAsync(out.prob.data(), prob_d.p, sizeof(float) * total, D2H, s));
  if (want_feat) {
    out.feat.resize((size_t)total * NF);
    DM_CUDA(ctx, cudaMemcpyAsync(out.feat.data(), feat_d.p, sizeof(float) * total * NF, D2H, s));
  }
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  return DM_OK;
}

int dm_cluster_sites_upload(dm_ctx* ctx, int32_t contig, int64_t n, const int64_t* pos, const int8_t* strand) {
  int rc = ensure_motif(ctx);
  if (rc != DM_OK) return rc;
  if (contig < 0 || contig >= ctx->n_contigs) { dm_set_error(ctx, "contig out of range"); return DM_ERR_ARG; }
  const int64_t len = ctx->contig_len[contig];
  uint8_t* blk = ctx->motif + 2 * ctx->contig_off[contig];
  DM_CUDA(ctx, cudaMemsetAsync(blk, 0, (size_t)(2 * len), ctx->stream));
  if (n > 0) {
    DevBuf<int64_t> p;
    DevBuf<int8_t> st;
    DM_CUDA(ctx, p.alloc(n, ctx->stream)); DM_CUDA(ctx, st.alloc(n, ctx->stream));
    DM_CUDA(ctx, cudaMemcpyAsync(p.p, pos, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    DM_CUDA(ctx, cudaMemcpyAsync(st.p, strand, n, cudaMemcpyHostToDevice, ctx->stream));
    k_set_sites<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, p.p, st.p, len, blk);
    ctx->launches += 1;
    DM_CUDA(ctx, cudaGetLastError());
    DM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  } else {
    DM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return DM_OK;
}

int dm_hist_load_rows(dm_ctx* ctx, int32_t contig, int8_t strand, int64_t n, const int64_t* pos, const int32_t* cov,
                      const int32_t* mod) {
  if (!ctx->cells) { dm_set_error(ctx, "dm_set_genome not called"); return DM_ERR_STATE; }
  if (contig < 0 || contig >= ctx->n_contigs) { dm_set_error(ctx, "contig out of range"); return DM_ERR_ARG; }
  if (n <= 0) return DM_OK;
  const int64_t len = ctx->contig_len[contig];
  unsigned long long* blk = ctx->cells + 2 * ctx->contig_off[contig] + (strand >= 0 ? 0 : len);
  DevBuf<int64_t> p;
  DevBuf<int32_t> c, m;
  DM_CUDA(ctx, p.alloc(n, ctx->stream)); DM_CUDA(ctx, c.alloc(n, ctx->stream)); DM_CUDA(ctx, m.alloc(n, ctx->stream));
  DM_CUDA(ctx, cudaMemcpyAsync(p.p, pos, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  DM_CUDA(ctx, cudaMemcpyAsync(c.p, cov, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  DM_CUDA(ctx, cudaMemcpyAsync(m.p, mod, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  k_hist_load<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, p.p, c.p, m.p, len, blk);
  ctx->launches += 1;
  DM_CUDA(ctx, cudaGetLastError());
  DM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DM_OK;
}
