// Event-table front-end on the GPU: raw-signal normalisation + per-event mean / stdv.
//
// Reference behaviour restated (bin/DeepMod_scripts/myDetect.py):
//   mnormalized :266-282  over the event span [start[0], start[-1]+length[-1]):
//                         shift = median(raw), scale = median(|raw - shift|); signal = (raw - shift) / scale;
//                         med / mad of the standardised span; clip the WHOLE signal at med +- 5 mad; np.round(.., 3)
//   getFast5Info :334-343 per event: mean = round(np.mean(seg), 3), stdv = round(np.std(seg), 3) -> '<f4' fields
// Everything is float64 in the reference.  Raw samples are int16, so every order statistic is taken EXACTLY from a
// 65536-bin histogram of the span (the standardised value is a monotone function of the raw value) and the
// per-sample value is recomputed on the fly; np.mean / np.std are reproduced with numpy's pairwise summation
// (8 accumulators up to 128 elements, recursive halving above) and IEEE operations without fma contraction.
#include "dm_common.cuh"

#include <algorithm>

namespace {

constexpr int NBIN = 65536;
constexpr int NT = 1024;

struct ReadNorm { double shift, scale, lower, upper; };

__device__ __forceinline__ double norm_value(int k, const ReadNorm& p) {
  double x = __ddiv_rn(__dsub_rn((double)k, p.shift), p.scale);
  x = x > p.upper ? p.upper : (x < p.lower ? p.lower : x);           // :282
  return __ddiv_rn(rint(__dmul_rn(x, 1000.0)), 1000.0);              // np.round(x, 3)
}

// number of span samples whose |2k - c2| <= x  (P = exclusive prefix sums of the histogram, bins = value + 32768)
__device__ __forceinline__ long long count_within(const int* __restrict__ P, long long c2, long long x) {
  long long lo2 = c2 - x, hi2 = c2 + x;
  long long k_lo = lo2 >= 0 ? (lo2 + 1) / 2 : -((-lo2) / 2);        // ceil(lo2 / 2)
  long long k_hi = hi2 >= 0 ? hi2 / 2 : -((-hi2 + 1) / 2);          // floor(hi2 / 2)
  k_lo = max(k_lo, -32768LL); k_hi = min(k_hi, 32767LL);
  if (k_hi < k_lo) return 0;
  return (long long)P[k_hi + 32768 + 1] - (long long)P[k_lo + 32768];
}

// one CTA per read (grid-stride): histogram -> medians -> clip limits
__global__ void __launch_bounds__(NT)
k_signal_norm(int n_reads, const int64_t* __restrict__ raw_off, const int16_t* __restrict__ raw,
              const int64_t* __restrict__ ev_off, const int64_t* __restrict__ ev_start,
              const int64_t* __restrict__ ev_length, int* __restrict__ scratch /*[grid][2][NBIN+1]*/,
              ReadNorm* __restrict__ out) {
  __shared__ int part[NT];
  int* hist = scratch + (size_t)blockIdx.x * 2 * (NBIN + 1);
  int* P = hist + (NBIN + 1);
  for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
    const int64_t e0 = ev_off[r], e1 = ev_off[r + 1];
    ReadNorm res = {0.0, 1.0, 0.0, 0.0};
    if (e1 <= e0) { if (threadIdx.x == 0) out[r] = res; continue; }
    // mdata[mean_start:mean_end] (myDetect.py:266-270): numpy slicing clips the span to the read's raw array
    const int64_t raw_len = raw_off[r + 1] - raw_off[r];
    const int64_t s0 = min(max(ev_start[e0], (int64_t)0), raw_len);
    const int64_t s1 = min(max(ev_start[e1 - 1] + ev_length[e1 - 1], s0), raw_len);
    const int16_t* x = raw + raw_off[r];
    const int64_t n = s1 - s0;
    if (n <= 0) { if (threadIdx.x == 0) out[r] = res; continue; }       // (uniform per CTA) empty span: identity normalisation
    for (int i = threadIdx.x; i < NBIN; i += NT) hist[i] = 0;
    __syncthreads();
    for (int64_t i = s0 + threadIdx.x; i < s1; i += NT) atomicAdd(&hist[(int)x[i] + 32768], 1);
    __syncthreads();
    // exclusive prefix sums over the bins: 64 consecutive bins per thread
    int sum = 0;
    for (int j = 0; j < NBIN / NT; ++j) sum += hist[threadIdx.x * (NBIN / NT) + j];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < NT; o <<= 1) {
      int t = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
      __syncthreads();
      part[threadIdx.x] += t;
      __syncthreads();
    }
    int run = part[threadIdx.x] - sum;
    for (int j = 0; j < NBIN / NT; ++j) {
      const int b = threadIdx.x * (NBIN / NT) + j;
      P[b] = run;
      run += hist[b];
    }
    if (threadIdx.x == NT - 1) P[NBIN] = run;
    __syncthreads();
    if (threadIdx.x == 0 && n > 0) {
      const long long r0 = (n - 1) / 2, r1 = n / 2;              // the one or two middle ranks (np.median)
      auto bin_of_rank = [&](long long rk) {                      // largest b with P[b] <= rk
        int lo = 0, hi = NBIN;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (P[mid] <= rk) lo = mid; else hi = mid; }
        return lo - 32768;
      };
      const int ka = bin_of_rank(r0), kb = bin_of_rank(r1);
      const double shift = __ddiv_rn(__dadd_rn((double)ka, (double)kb), 2.0);
      const long long c2 = (long long)ka + kb;                   // 2 * shift, exact
      auto d2_of_rank = [&](long long rk) {                       // smallest x with count_within(x) >= rk + 1
        long long lo = -1, hi = 2LL * NBIN;
        while (hi - lo > 1) { long long mid = (lo + hi) >> 1; if (count_within(P, c2, mid) >= rk + 1) hi = mid; else lo = mid; }
        return hi;
      };
      const long long da = d2_of_rank(r0), db = d2_of_rank(r1);
      const double scale = __ddiv_rn(__dadd_rn((double)da / 2.0, (double)db / 2.0), 2.0);
      res.shift = shift; res.scale = scale;
      const double sa = __ddiv_rn(__dsub_rn((double)ka, shift), scale), sb = __ddiv_rn(__dsub_rn((double)kb, shift), scale);
      const double med = __ddiv_rn(__dadd_rn(sa, sb), 2.0);
      // |standardised - med| at rank rk: the bucket of distance d2 holds at most two raw values (left / right of the
      // centre); med is 0 up to rounding, so buckets keep their order and only the order inside a bucket depends on it
      auto dev_of_rank = [&](long long rk, long long d2) {
        const long long before = d2 > 0 ? count_within(P, c2, d2 - 1) : 0;
        const long long kl2 = c2 - d2, kr2 = c2 + d2;            // 2 * raw value on either side
        double v[2]; long long cnt[2]; int m = 0;
        if ((kl2 & 1) == 0) {
          const long long k = kl2 / 2;
          if (k >= -32768 && k <= 32767 && hist[k + 32768] > 0) {
            v[m] = fabs(__dsub_rn(__ddiv_rn(__dsub_rn((double)k, shift), scale), med)); cnt[m] = hist[k + 32768]; ++m;
          }
        }
        if (d2 > 0 && (kr2 & 1) == 0) {
          const long long k = kr2 / 2;
          if (k >= -32768 && k <= 32767 && hist[k + 32768] > 0) {
            v[m] = fabs(__dsub_rn(__ddiv_rn(__dsub_rn((double)k, shift), scale), med)); cnt[m] = hist[k + 32768]; ++m;
          }
        }
        if (m == 2 && v[1] < v[0]) { double tv = v[0]; v[0] = v[1]; v[1] = tv; long long tc = cnt[0]; cnt[0] = cnt[1]; cnt[1] = tc; }
        return (m == 2 && rk - before >= cnt[0]) ? v[1] : v[0];
      };
      const double ma = dev_of_rank(r0, da), mb = dev_of_rank(r1, db);
      const double mad = __ddiv_rn(__dadd_rn(ma, mb), 2.0);
      res.lower = __dsub_rn(med, __dmul_rn(mad, 5.0));
      res.upper = __dadd_rn(med, __dmul_rn(mad, 5.0));
      out[r] = res;
    }
    __syncthreads();
  }
}

// numpy's pairwise summation of f(i), i in [0, n)  (numpy/core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum)
template <typename F>
__device__ double pairwise_sum(F f, int64_t base, int64_t n) {
  if (n < 8) {
    double res = 0.0;
    for (int64_t i = 0; i < n; ++i) res = __dadd_rn(res, f(base + i));
    return res;
  }
  if (n <= 128) {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(base + j);
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8)
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f(base + i + j));
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, f(base + i));
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(pairwise_sum(f, base, n2), pairwise_sum(f, base + n2, n - n2));
}

// one thread per event
__global__ void k_event_stats(int n_reads, int64_t n_events, const int64_t* __restrict__ raw_off,
                              const int16_t* __restrict__ raw, const int64_t* __restrict__ ev_off,
                              const int64_t* __restrict__ ev_start, const int64_t* __restrict__ ev_length,
                              const ReadNorm* __restrict__ norm, float* __restrict__ mean_out, float* __restrict__ stdv_out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_events) return;
  int lo = 0, hi = n_reads;
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (ev_off[mid] <= e) lo = mid; else hi = mid; }
  const int r = lo;
  const ReadNorm p = norm[r];
  const int16_t* x = raw + raw_off[r] + ev_start[e];
  const int64_t n = ev_length[e], avail = raw_off[r + 1] - raw_off[r] - ev_start[e];
  const int64_t m = n < avail ? n : (avail > 0 ? avail : 0);        // a slice past the end is shorter (python slicing)
  if (m <= 0) { mean_out[e] = 0.f; stdv_out[e] = 0.f; return; }
  auto val = [&](int64_t i) { return norm_value((int)x[i], p); };
  const double mean = __ddiv_rn(pairwise_sum(val, 0, m), (double)m);
  auto sq = [&](int64_t i) { const double d = __dsub_rn(norm_value((int)x[i], p), mean); return __dmul_rn(d, d); };
  const double var = __ddiv_rn(pairwise_sum(sq, 0, m), (double)m);
  const double sd = __dsqrt_rn(var);
  mean_out[e] = (float)__ddiv_rn(rint(__dmul_rn(mean, 1000.0)), 1000.0);      // round(np.mean(..), 3) -> '<f4'
  stdv_out[e] = (float)__ddiv_rn(rint(__dmul_rn(sd, 1000.0)), 1000.0);
}

template <typename T>
struct Buf {
  T* p = nullptr;
  cudaStream_t st = nullptr;
  ~Buf() { if (p) cudaFreeAsync(p, st); }
  cudaError_t alloc(size_t n, cudaStream_t s) { st = s; return cudaMallocAsync(reinterpret_cast<void**>(&p), sizeof(T) * std::max<size_t>(n, 1), s); }
  cudaError_t upload(const T* h, size_t n, cudaStream_t s) {
    cudaError_t e = alloc(n, s);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(p, h, sizeof(T) * n, cudaMemcpyHostToDevice, s);
    return e;
  }
};

}  // namespace

int dm_signal_event_stats(dm_ctx* ctx, int32_t n_reads, const int64_t* raw_off, const int16_t* raw, const int64_t* ev_off,
                          const int64_t* ev_start, const int64_t* ev_length, float* mean_out, float* stdv_out) {
  if (n_reads == 0) return DM_OK;
  cudaStream_t s = ctx->stream;
  const int64_t n_raw = raw_off[n_reads], n_ev = ev_off[n_reads];
  if (n_ev == 0) return DM_OK;
  for (int64_t e = 0; e < n_ev; ++e)
    if (ev_start[e] < 0 || ev_length[e] < 0) { dm_set_error(ctx, "dm_event_stats: negative event start/length"); return DM_ERR_ARG; }
  Buf<int64_t> d_raw_off, d_ev_off, d_start, d_len;
  Buf<int16_t> d_raw;
  Buf<int> d_scratch;
  Buf<ReadNorm> d_norm;
  Buf<float> d_mean, d_stdv;
  const int grid = std::min(n_reads, 2 * ctx->sm_count);
  DM_CUDA(ctx, d_raw_off.upload(raw_off, (size_t)n_reads + 1, s));
  DM_CUDA(ctx, d_ev_off.upload(ev_off, (size_t)n_reads + 1, s));
  DM_CUDA(ctx, d_start.upload(ev_start, (size_t)n_ev, s));
  DM_CUDA(ctx, d_len.upload(ev_length, (size_t)n_ev, s));
  DM_CUDA(ctx, d_raw.upload(raw, (size_t)n_raw, s));
  DM_CUDA(ctx, d_scratch.alloc((size_t)grid * 2 * (NBIN + 1), s));
  DM_CUDA(ctx, d_norm.alloc(n_reads, s));
  DM_CUDA(ctx, d_mean.alloc(n_ev, s));
  DM_CUDA(ctx, d_stdv.alloc(n_ev, s));
  DM_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
  k_signal_norm<<<grid, NT, 0, s>>>(n_reads, d_raw_off.p, d_raw.p, d_ev_off.p, d_start.p, d_len.p, d_scratch.p, d_norm.p);
  k_event_stats<<<(unsigned)((n_ev + 127) / 128), 128, 0, s>>>(n_reads, n_ev, d_raw_off.p, d_raw.p, d_ev_off.p, d_start.p,
                                                              d_len.p, d_norm.p, d_mean.p, d_stdv.p);
  ctx->launches += 2;
  DM_CUDA(ctx, cudaGetLastError());
  DM_CUDA(ctx, cudaEventRecord(ctx->ev3, s));
  DM_CUDA(ctx, cudaMemcpyAsync(mean_out, d_mean.p, sizeof(float) * n_ev, cudaMemcpyDeviceToHost, s));
  DM_CUDA(ctx, cudaMemcpyAsync(stdv_out, d_stdv.p, sizeof(float) * n_ev, cudaMemcpyDeviceToHost, s));
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  DM_CUDA(ctx, cudaEventElapsedTime(&ctx->total_ms, ctx->ev0, ctx->ev3));
  ctx->lstm_ms = 0.f;
  return DM_OK;
}
