// Feature-table builder and window gather for the packed read batch.
//
// Reference behaviour restated on the GPU (bin/DeepMod_scripts/myDetect.py):
//   get_Feature :839-903  -> one row per event: one-hot(reference base of the event's
//                            alignment column) + mean/stdv/length; rows of clipped events
//                            keep the signal stats but a zero one-hot; rows outside the
//                            read are all zero.  The reference keeps a +-100 row flank,
//                            only +-10 rows can reach a window, so the table keeps +-10.
//   mPredict1   :791-803  -> window m of a read = rows [m, m+21) of that read's table.
// The table is the only materialised form: 32 B per event instead of 588 B per window;
// the BiLSTM kernels read rows through win_frow[] (first row of each window).
#include "dm_common.cuh"

#include <cuda_fp16.h>

namespace {

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;                       // columns per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;  // 2048 columns per block

__device__ __forceinline__ int find_segment(const int64_t* __restrict__ off, int n, int64_t x) {
  // largest r in [0,n) with off[r] <= x   (off is ascending, off[n] > x)
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// pass 1: number of non-gap read bases per 2048-column tile
__global__ void k_nongap_count(const uint8_t* __restrict__ readbase, int64_t n_cols,
                               int64_t* __restrict__ tile_sum) {
  __shared__ int warp_sum[SCAN_BLOCK / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t c = base + i * SCAN_BLOCK + threadIdx.x;      // coalesced byte loads
    if (c < n_cols) cnt += (readbase[c] != '-');
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < SCAN_BLOCK / 32; ++w) s += warp_sum[w];
    tile_sum[blockIdx.x] = s;
  }
}

// pass 2: exclusive scan of the tile sums (one block, sequential over chunks of 1024)
__global__ void k_scan_tiles(int64_t* __restrict__ tile_sum, int n_tiles) {
  __shared__ int64_t buf[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += 1024) {
    int i = base + threadIdx.x;
    int64_t v = i < n_tiles ? tile_sum[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int64_t t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n_tiles) tile_sum[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
}

// pass 3: per-column exclusive rank of non-gap read bases (global numbering)
__global__ void k_nongap_rank(const uint8_t* __restrict__ readbase, int64_t n_cols,
                              const int64_t* __restrict__ tile_off, int64_t* __restrict__ col_rank) {
  __shared__ int warp_tot[SCAN_BLOCK / 32];
  // blocked arrangement: thread t owns columns [t*8, t*8+8) of the tile
  int64_t c0 = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int flags[SCAN_ITEMS];
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    flags[i] = (c0 + i < n_cols) ? (readbase[c0 + i] != '-') : 0;
    cnt += flags[i];
  }
  int incl = cnt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int warp_base = 0;
  for (int w = 0; w < warp; ++w) warp_base += warp_tot[w];
  int64_t run = tile_off[blockIdx.x] + warp_base + incl - cnt;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (c0 + i < n_cols) col_rank[c0 + i] = run;
    run += flags[i];
  }
}

// per read: status from the counts (Less Event :702-705, bad alignment), reset win_col
__global__ void k_read_status(int n_reads, const int64_t* __restrict__ ev_off,
                              const int64_t* __restrict__ col_off, const int64_t* __restrict__ col_rank,
                              const uint8_t* __restrict__ readbase, int64_t n_cols,
                              const int32_t* __restrict__ start_clip, const int32_t* __restrict__ end_clip,
                              int32_t* __restrict__ status) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  int64_t lmap = (ev_off[r + 1] - ev_off[r]) - start_clip[r] - end_clip[r];
  int64_t c0 = col_off[r], c1 = col_off[r + 1];
  int64_t nongap = 0;
  if (c1 > c0) {
    int64_t last = col_rank[c1 - 1] + (readbase[c1 - 1] != '-');
    nongap = last - col_rank[c0];
  }
  int st = DM_READ_OK;
  if (lmap < 50) st = DM_READ_LESS_EVENT;
  else if (nongap != lmap) st = DM_READ_BAD_ALIGN;
  status[r] = st;
}

// per column: window <- column map, and the read-base vs k-mer-centre check (:868-874)
__global__ void k_map_columns(int n_reads, int64_t n_cols, const int64_t* __restrict__ col_off,
                              const int64_t* __restrict__ col_rank, const uint8_t* __restrict__ readbase,
                              const int64_t* __restrict__ ev_off, const uint8_t* __restrict__ ev_base,
                              const int32_t* __restrict__ start_clip, const int64_t* __restrict__ win_off,
                              int64_t* __restrict__ win_col, int32_t* __restrict__ status) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  uint8_t rb = readbase[c];
  if (rb == '-') return;
  int r = find_segment(col_off, n_reads, c);
  int64_t k = col_rank[c] - col_rank[col_off[r]];          // mapped-event index inside the read
  int64_t lmap = win_off[r + 1] - win_off[r];
  if (k >= lmap) return;                                   // bad alignment, flagged by k_read_status
  win_col[win_off[r] + k] = c;
  if (ev_base != nullptr) {
    uint8_t eb = ev_base[ev_off[r] + start_clip[r] + k];
    if (eb != rb) atomicMax(&status[r], (int)DM_READ_MISMATCH);   // OK(0) -> MISMATCH(1) only
  }
}

// 16-bit hi/lo split row for the tensor-core path, in the K order of the layer-0 weight image:
// [A C G T mean_hi stdv_hi len_hi len_lo | 1 1 mean_lo stdv_lo | pad]; bf16 or fp16 operands
__device__ __forceinline__ void tc_split_row(const float (&f)[8], uint16_t* dst, bool f16) {
  auto rn = [f16](float x) -> uint16_t {
    return f16 ? __half_as_ushort(__float2half_rn(x)) : __bfloat16_as_ushort(__float2bfloat16(x));
  };
  auto back = [f16](uint16_t b) -> float {
    return f16 ? __half2float(__ushort_as_half(b)) : __bfloat162float(__ushort_as_bfloat16(b));
  };
  __align__(16) uint16_t h[16];
  for (int i = 0; i < 4; ++i) h[i] = rn(f[i]);
  const uint16_t mh = rn(f[4]), sh = rn(f[5]), lh = rn(f[6]);
  h[4] = mh; h[5] = sh; h[6] = lh;
  h[7] = rn(f[6] - back(lh));
  h[8] = h[9] = rn(1.f);
  h[10] = rn(f[4] - back(mh));
  h[11] = rn(f[5] - back(sh));
  h[12] = h[13] = h[14] = h[15] = 0;
  uint4* ot = reinterpret_cast<uint4*>(dst);
  const uint4* hs = reinterpret_cast<const uint4*>(h);
  ot[0] = hs[0];
  ot[1] = hs[1];
}

// One thread per feature row: ie = start_clip - 10 + j walks the +-10 flank (:855-900).
__global__ void k_feature_rows(int n_reads, int64_t n_frows, const int64_t* __restrict__ win_off,
                               const int64_t* __restrict__ ev_off, const float* __restrict__ ev_mean,
                               const float* __restrict__ ev_stdv, const float* __restrict__ ev_len,
                               const int32_t* __restrict__ start_clip, const int32_t* __restrict__ end_clip,
                               const int64_t* __restrict__ win_col, const uint8_t* __restrict__ refbase,
                               float* __restrict__ feat, uint16_t* __restrict__ feat_tc, bool f16) {
  int64_t fr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (fr >= n_frows + DM_WINDOW) return;       // rows >= n_frows: the shared all-zero window
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (fr < n_frows) {
    // frow_off[r] = win_off[r] + 2*FLANK*r : invert by bisection on that affine-shifted key
    int lo = 0, hi = n_reads;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (win_off[mid] + (int64_t)(2 * DM_FLANK) * mid <= fr) lo = mid; else hi = mid;
    }
    const int r = lo;
    const int64_t j = fr - (win_off[r] + (int64_t)(2 * DM_FLANK) * r);
    const int64_t L = ev_off[r + 1] - ev_off[r];
    const int64_t sc = start_clip[r], ec = end_clip[r];
    const int64_t ie = sc - DM_FLANK + j;
    if (ie >= 0 && ie < L) {
      const int64_t e = ev_off[r] + ie;
      f[4] = ev_mean[e];
      f[5] = ev_stdv[e];
      f[6] = ev_len[e];
      // reads with fewer than 50 mapped events own no windows (and no win_col entries)
      if (ie >= sc && ie - sc < win_off[r + 1] - win_off[r]) {
        int64_t c = win_col[win_off[r] + (ie - sc)];
        if (c >= 0) {
          uint8_t b = refbase[c];
          if (b == 'A') f[0] = 1.f; else if (b == 'C') f[1] = 1.f;
          else if (b == 'G') f[2] = 1.f; else if (b == 'T') f[3] = 1.f;
        }
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(feat + fr * DM_FEAT_STRIDE);
  o[0] = make_float4(f[0], f[1], f[2], f[3]);
  o[1] = make_float4(f[4], f[5], f[6], 0.f);
  if (feat_tc != nullptr) tc_split_row(f, feat_tc + fr * 16, f16);
}

// win_frow[w] = first feature row of window w; padded tail -> the zero row
__global__ void k_window_rows(int n_reads, int64_t n_windows, int64_t n_padded, int64_t n_frows,
                              const int64_t* __restrict__ win_off, int32_t* __restrict__ win_frow) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_padded) return;
  if (w >= n_windows) { win_frow[w] = (int32_t)n_frows; return; }   // padded tail -> zero rows
  int r = find_segment(win_off, n_reads, w);
  win_frow[w] = (int32_t)(w + (int64_t)(2 * DM_FLANK) * r);
}

// explicit windows [n,21,7] (the b1 seam) -> 21 feature rows each
__global__ void k_windows_to_rows(const float* __restrict__ X, int64_t n, int64_t n_padded,
                                  float* __restrict__ feat, uint16_t* __restrict__ feat_tc, bool f16,
                                  int32_t* __restrict__ win_frow) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // row index
  int64_t n_rows = n * DM_WINDOW;
  if (i < n_padded) win_frow[i] = (int32_t)((i < n ? i : n) * DM_WINDOW);   // tail -> zero rows
  if (i >= n_rows + DM_WINDOW) return;
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (i < n_rows)
    for (int k = 0; k < DM_FNUM; ++k) f[k] = X[i * DM_FNUM + k];
  float4* o = reinterpret_cast<float4*>(feat + i * DM_FEAT_STRIDE);
  o[0] = make_float4(f[0], f[1], f[2], f[3]);
  o[1] = make_float4(f[4], f[5], f[6], 0.f);
  // (general inputs: the four base columns are taken as-is, they are 0/1 in every reference window)
  if (feat_tc != nullptr) tc_split_row(f, feat_tc + i * 16, f16);
}

// coalesced gather: thread per output float of the [n_windows,21,7] tensor
__global__ void k_gather_windows(int64_t n_windows, const int32_t* __restrict__ win_frow,
                                 const float* __restrict__ feat, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n_windows * (DM_WINDOW * DM_FNUM);
  if (i >= total) return;
  int64_t w = i / (DM_WINDOW * DM_FNUM);
  int rem = (int)(i - w * (DM_WINDOW * DM_FNUM));
  int t = rem / DM_FNUM, k = rem - t * DM_FNUM;
  out[i] = feat[((int64_t)win_frow[w] + t) * DM_FEAT_STRIDE + k];
}

__global__ void k_fill_i64(int64_t* p, int64_t n, int64_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

static inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// Builds every derived array of the uploaded batch (ctx->b): ranks, window<-column map,
// per-read status, feature rows (fp32 and bf16 hi/lo), window row index.
int dm_launch_prepare(dm_ctx* ctx) {
  dm_dev_batch& b = ctx->b;
  cudaStream_t s = ctx->stream;
  if (b.n_reads == 0) return DM_OK;
  const int n_tiles = (int)((b.n_cols + SCAN_TILE - 1) / SCAN_TILE);
  size_t need = sizeof(int64_t) * (size_t)(n_tiles + 1);
  if (need > ctx->scratch_bytes) {
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    DM_CUDA(ctx, cudaMalloc(&ctx->scratch, need * 2));
    ctx->scratch_bytes = need * 2;
  }
  int64_t* tile_sum = static_cast<int64_t*>(ctx->scratch);
  if (b.n_cols > 0) {
    k_nongap_count<<<n_tiles, SCAN_BLOCK, 0, s>>>(b.col_readbase, b.n_cols, tile_sum);
    k_scan_tiles<<<1, 1024, 0, s>>>(tile_sum, n_tiles);
    k_nongap_rank<<<n_tiles, SCAN_BLOCK, 0, s>>>(b.col_readbase, b.n_cols, tile_sum, b.col_rank);
    ctx->launches += 3;
  }
  k_read_status<<<blocks_for(b.n_reads, 128), 128, 0, s>>>(b.n_reads, b.ev_off, b.col_off, b.col_rank,
                                                           b.col_readbase, b.n_cols, b.start_clip,
                                                           b.end_clip, b.status);
  ctx->launches += 1;
  if (b.n_windows > 0) {
    k_fill_i64<<<blocks_for(b.n_windows, 256), 256, 0, s>>>(b.win_col, b.n_windows, -1);
    ctx->launches += 1;
  }
  if (b.n_cols > 0) {
    k_map_columns<<<blocks_for(b.n_cols, 256), 256, 0, s>>>(
        b.n_reads, b.n_cols, b.col_off, b.col_rank, b.col_readbase, b.ev_off,
        b.has_ev_base ? b.ev_base : nullptr, b.start_clip, b.win_off, b.win_col, b.status);
    ctx->launches += 1;
  }
  k_feature_rows<<<blocks_for(b.n_frows + DM_WINDOW, 256), 256, 0, s>>>(
      b.n_reads, b.n_frows, b.win_off, b.ev_off, b.ev_mean, b.ev_stdv, b.ev_len, b.start_clip,
      b.end_clip, b.win_col, b.col_refbase, b.feat, b.feat_tc, ctx->tc_f16);
  const int64_t n_pad = dm_pad_windows(b.n_windows);
  if (n_pad > 0) {
    k_window_rows<<<blocks_for(n_pad, 256), 256, 0, s>>>(b.n_reads, b.n_windows, n_pad, b.n_frows,
                                                         b.win_off, b.win_frow);
    ctx->launches += 1;
  }
  ctx->launches += 1;
  DM_CUDA(ctx, cudaGetLastError());
  return DM_OK;
}

int dm_launch_build_windows(dm_ctx* ctx, float* out_d) {
  dm_dev_batch& b = ctx->b;
  int64_t total = b.n_windows * (DM_WINDOW * DM_FNUM);
  if (total == 0) return DM_OK;
  k_gather_windows<<<blocks_for(total, 256), 256, 0, ctx->stream>>>(b.n_windows, b.win_frow, b.feat, out_d);
  ctx->launches += 1;
  DM_CUDA(ctx, cudaGetLastError());
  return DM_OK;
}

int dm_launch_windows_to_rows(dm_ctx* ctx, const float* X_d, int64_t n, float* feat,
                              uint16_t* feat_tc, int32_t* win_frow) {
  int64_t n_pad = dm_pad_windows(n);
  int64_t threads = (n + 1) * DM_WINDOW;
  if (threads < n_pad) threads = n_pad;
  k_windows_to_rows<<<blocks_for(threads, 256), 256, 0, ctx->stream>>>(X_d, n, n_pad, feat, feat_tc, ctx->tc_f16, win_frow);
  ctx->launches += 1;
  DM_CUDA(ctx, cudaGetLastError());
  return DM_OK;
}
