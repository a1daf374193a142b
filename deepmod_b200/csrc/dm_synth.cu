// Device-side synthetic read generator (benchmark workload, SURVEY 8(d) row 2 / BASELINE configs[2]):
// "10 M synthetic reads (~8 kb), all-match alignments, generated on-device from a counter-based RNG per
// read id, never materialised on the host".  Distributions are the ones of deepmod_b200/synth.py
// (SURVEY 8(d) row 0): length ~ Gamma(2, mean/2) clipped to [lo, hi]; clips ~ U{0..max_clip}; strand +-1;
// contig ~ length; start uniform; per event mean ~ round(clip(N(0, 1.4), +-5), 3), stdv ~ round(|N(0.25, 0.12)|, 3),
// length = 2 + Geometric(0.12); the genome is iid uniform ACGT, a pure function of (contig, position).
//
// Every value is a function of (seed, read id, index) only -- Philox4x32-10 -- so any GPU generates any read
// range identically: a read set of R reads can be sharded over 1, 2, 4 or 8 GPUs (strong scaling) and must
// reduce to the same accumulator.  The generator writes the resident batch (ctx->b) exactly as dm_batch_upload
// leaves it; dm_detect_resident / dm_fetch_results / dm_fetch_inputs follow.
#include "dm_common.cuh"

#include <algorithm>
#include <cmath>

namespace {

struct Philox {
  static __host__ __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
  }
  // 4 x 32 random bits for counter (a, b, c, d) under `seed`
  static __host__ __device__ __forceinline__ void draw(uint64_t seed, uint64_t a, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)a, (uint32_t)(a >> 32), c2, c3};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

__device__ __forceinline__ float u01(uint32_t x) { return ((float)x + 0.5f) * 2.3283064365386963e-10f; }   // (0, 1)

constexpr uint32_t STREAM_READ = 0, STREAM_READ2 = 1, STREAM_EVENT = 2, STREAM_GENOME = 3;

__device__ __forceinline__ uint8_t genome_base(uint64_t seed, int contig, int64_t pos) {
  uint32_t w[4];
  Philox::draw(seed ^ 0x67656e6f6d65ull, (uint64_t)(pos >> 2), (uint32_t)contig, STREAM_GENOME, w);
  // 4 words x 16 two-bit bases per counter would be overkill: one word serves 4 consecutive positions
  const uint32_t two = (w[pos & 3] >> 7) & 3u;
  return (uint8_t)("ACGT"[two]);
}
__device__ __forceinline__ uint8_t complement(uint8_t b) { return b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : 'A'; }

struct ReadParam { int32_t L, sc, ec, contig; int64_t start; int8_t strand; };

__device__ __forceinline__ ReadParam read_param(const dm_synth_spec sp, int64_t read_id, const int64_t* __restrict__ contig_off,
                                                int n_contigs) {
  uint32_t w[4], v[4];
  Philox::draw(sp.seed, (uint64_t)read_id, 0u, STREAM_READ, w);
  Philox::draw(sp.seed, (uint64_t)read_id, 0u, STREAM_READ2, v);
  ReadParam p;
  float g;
  if (sp.length_kind == 1) g = expf(logf((float)sp.len_lo) + u01(w[0]) * (logf((float)sp.len_hi) - logf((float)sp.len_lo)));   // log-uniform
  else g = -0.5f * sp.mean_len * (logf(u01(w[0])) + logf(u01(w[1])));               // Gamma(2, mean / 2)
  int64_t L = (int64_t)fminf(fmaxf(g, (float)sp.len_lo), (float)sp.len_hi);
  p.sc = (int32_t)(w[2] % (uint32_t)(sp.max_clip + 1));
  p.ec = (int32_t)(w[3] % (uint32_t)(sp.max_clip + 1));
  p.strand = (v[0] & 1u) ? 1 : -1;
  const int64_t total = contig_off[n_contigs];
  const uint64_t r64 = ((uint64_t)v[1] << 32) | v[2];
  const int64_t g_at = (int64_t)(((unsigned __int128)r64 * (unsigned __int128)total) >> 64);   // uniform position in the genome
  int c = 0;
  while (c + 1 < n_contigs && contig_off[c + 1] <= g_at) ++c;
  p.contig = c;
  const int64_t clen = contig_off[c + 1] - contig_off[c];
  if (L - p.sc - p.ec > clen) L = clen + p.sc + p.ec;                                  // a read cannot be longer than its contig
  p.L = (int32_t)L;
  const int64_t lmap = L - p.sc - p.ec;
  const int64_t room = clen - lmap + 1;
  p.start = room > 1 ? (int64_t)(((uint64_t)v[3] * (uint64_t)room) >> 32) : 0;
  return p;
}

__global__ void k_synth_reads(dm_synth_spec sp, int64_t first_read, int n, const int64_t* __restrict__ contig_off, int n_contigs,
                              int32_t* __restrict__ L_out, int32_t* __restrict__ sc, int32_t* __restrict__ ec,
                              int32_t* __restrict__ contig, int8_t* __restrict__ strand, int64_t* __restrict__ start) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const ReadParam p = read_param(sp, first_read + r, contig_off, n_contigs);
  L_out[r] = p.L; sc[r] = p.sc; ec[r] = p.ec; contig[r] = p.contig; strand[r] = p.strand;
  if (start) start[r] = p.start;
}

// one thread per event; an all-match alignment has exactly one column per mapped event
__global__ void k_synth_events(dm_synth_spec sp, int64_t first_read, int n, int64_t n_events, const int64_t* __restrict__ ev_off,
                               const int64_t* __restrict__ col_off, const int32_t* __restrict__ sc_a, const int32_t* __restrict__ ec_a,
                               const int32_t* __restrict__ contig_a, const int8_t* __restrict__ strand_a,
                               const int64_t* __restrict__ start_a, float* __restrict__ ev_mean, float* __restrict__ ev_stdv,
                               float* __restrict__ ev_len, uint8_t* __restrict__ ev_base, uint8_t* __restrict__ col_refbase,
                               uint8_t* __restrict__ col_readbase, int64_t* __restrict__ col_refpos) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_events) return;
  int lo = 0, hi = n;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ev_off[mid] <= e) lo = mid; else hi = mid; }
  const int r = lo;
  const int64_t ie = e - ev_off[r];
  uint32_t w[4];
  Philox::draw(sp.seed, (uint64_t)(first_read + r), (uint32_t)ie, STREAM_EVENT, w);
  // Box-Muller pair
  const float rad = sqrtf(-2.0f * logf(u01(w[0])));
  float sn, cs;
  sincospif(2.0f * u01(w[1]), &sn, &cs);
  const double m = fmin(fmax(1.4 * (double)(rad * cs), -5.0), 5.0);
  const double s = fabs(0.25 + 0.12 * (double)(rad * sn));
  ev_mean[e] = (float)(rint(m * 1000.0) / 1000.0);
  ev_stdv[e] = (float)(rint(s * 1000.0) / 1000.0);
  // numpy's geometric(p): ceil(log(u) / log1p(-p)), support 1, 2, ...
  ev_len[e] = 2.0f + fmaxf(1.0f, ceilf(logf(u01(w[2])) / -0.12783337150988489f));
  const int64_t L = ev_off[r + 1] - ev_off[r];
  const int sc = sc_a[r], ec = ec_a[r];
  uint8_t base = (uint8_t)("ACGT"[w[3] & 3u]);
  if (ie >= sc && ie < L - ec) {
    const int64_t k = ie - sc, lmap = L - sc - ec;
    const bool fwd = strand_a[r] >= 0;
    const int64_t pos = fwd ? start_a[r] + k : start_a[r] + lmap - 1 - k;         // read orientation (myDetect.py:661-666)
    const uint8_t gb = genome_base(sp.seed, contig_a[r], pos);
    base = fwd ? gb : complement(gb);
    const int64_t c = col_off[r] + k;
    col_refbase[c] = base;
    col_readbase[c] = base;
    col_refpos[c] = pos;
  }
  ev_base[e] = base;
}

int check_spec(dm_ctx* ctx, const dm_synth_spec* sp, int64_t first_read, int32_t n) {
  if (!sp || first_read < 0 || n < 0) return DM_ERR_ARG;
  if (sp->length_kind != 0 && sp->length_kind != 1) { dm_set_error(ctx, "dm_synth: length_kind must be 0 (gamma) or 1 (log-uniform)"); return DM_ERR_ARG; }
  if (!(sp->mean_len > 0.f) || sp->len_lo < 1 || sp->len_hi < sp->len_lo || sp->max_clip < 0 || 2 * sp->max_clip >= sp->len_lo) {
    dm_set_error(ctx, "dm_synth: need mean_len > 0, 1 <= len_lo <= len_hi, 0 <= 2 * max_clip < len_lo");
    return DM_ERR_ARG;
  }
  if (!ctx->cells) { dm_set_error(ctx, "dm_synth: dm_set_genome not called (reads are placed on its contigs)"); return DM_ERR_STATE; }
  return DM_OK;
}

}  // namespace

extern "C" {

int dm_synth_describe(dm_ctx* ctx, const dm_synth_spec* sp, int64_t first_read, int32_t n_reads, int32_t* n_events_out,
                      int32_t* n_windows_out) {
  if (!ctx) return DM_ERR_ARG;
  int rc = check_spec(ctx, sp, first_read, n_reads);
  if (rc != DM_OK) return rc;
  if (n_reads == 0) return DM_OK;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  int32_t* d = nullptr;
  int8_t* st = nullptr;
  DM_CUDA(ctx, cudaMalloc(&d, sizeof(int32_t) * 4 * (size_t)n_reads));
  DM_CUDA(ctx, cudaMalloc(&st, (size_t)n_reads));
  k_synth_reads<<<(n_reads + 127) / 128, 128, 0, s>>>(*sp, first_read, n_reads, ctx->contig_off_d, ctx->n_contigs, d, d + n_reads,
                                                      d + 2 * (size_t)n_reads, d + 3 * (size_t)n_reads, st, nullptr);
  ctx->launches += 1;
  std::vector<int32_t> h(3 * (size_t)n_reads);
  cudaError_t e = cudaMemcpyAsync(h.data(), d, sizeof(int32_t) * 3 * (size_t)n_reads, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d); cudaFree(st);
  if (e != cudaSuccess) { dm_set_error(ctx, std::string("dm_synth_describe: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  for (int r = 0; r < n_reads; ++r) {
    const int32_t lmap = h[r] - h[n_reads + r] - h[2 * (size_t)n_reads + r];
    if (n_events_out) n_events_out[r] = h[r];
    if (n_windows_out) n_windows_out[r] = lmap >= 50 ? lmap : 0;
  }
  return DM_OK;
}

int dm_synth_generate(dm_ctx* ctx, const dm_synth_spec* sp, int64_t first_read, int32_t n_reads, int64_t* n_windows_out) {
  if (!ctx) return DM_ERR_ARG;
  int rc = check_spec(ctx, sp, first_read, n_reads);
  if (rc != DM_OK) return rc;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  dm_dev_batch& b = ctx->b;
  cudaStream_t s = ctx->stream;
  const int n = n_reads;
  b.n_reads = n;
  b.n_events = b.n_cols = b.n_windows = b.n_frows = 0;
  if (n_windows_out) *n_windows_out = 0;
  if (n == 0) return DM_OK;
  // per-read parameters first (their sizes decide every offset); the per-read arrays of the batch receive them directly
  rc = dm_batch_reserve(ctx, n, 0, 0, 0);
  if (rc != DM_OK) return rc;
  int32_t* L_d = nullptr;
  int64_t* start_d = nullptr;
  DM_CUDA(ctx, cudaMalloc(&L_d, sizeof(int32_t) * (size_t)n));
  DM_CUDA(ctx, cudaMalloc(&start_d, sizeof(int64_t) * (size_t)n));
  auto release = [&]() { cudaFree(L_d); cudaFree(start_d); };
  k_synth_reads<<<(n + 127) / 128, 128, 0, s>>>(*sp, first_read, n, ctx->contig_off_d, ctx->n_contigs, L_d, b.start_clip, b.end_clip,
                                                b.contig, b.strand, start_d);
  ctx->launches += 1;
  std::vector<int32_t> L((size_t)n), sc((size_t)n), ec((size_t)n);
  cudaError_t e = cudaMemcpyAsync(L.data(), L_d, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(sc.data(), b.start_clip, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ec.data(), b.end_clip, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { release(); dm_set_error(ctx, std::string("dm_synth_generate: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  std::vector<int64_t> ev_off((size_t)n + 1, 0), col_off((size_t)n + 1, 0), win_off((size_t)n + 1, 0);
  for (int r = 0; r < n; ++r) {
    const int64_t lmap = (int64_t)L[r] - sc[r] - ec[r];
    ev_off[r + 1] = ev_off[r] + L[r];
    col_off[r + 1] = col_off[r] + lmap;
    win_off[r + 1] = win_off[r] + (lmap >= 50 ? lmap : 0);                 // 'Less Event', myDetect.py:702-705
  }
  const int64_t n_events = ev_off[n], n_cols = col_off[n], n_windows = win_off[n];
  const int64_t n_frows = n_windows + (int64_t)(2 * DM_FLANK) * n;
  if (n_frows + DM_WINDOW >= (int64_t)INT32_MAX) { release(); dm_set_error(ctx, "dm_synth_generate: batch too large (>2^31 rows)"); return DM_ERR_ARG; }
  // growing the per-read arrays again would free what the kernel above just wrote: n is already reserved, only the
  // event / column / window arrays grow here
  rc = dm_batch_reserve(ctx, n, n_events, n_cols, n_windows);
  if (rc != DM_OK) { release(); return rc; }
  e = cudaMemcpyAsync(b.ev_off, ev_off.data(), sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(b.col_off, col_off.data(), sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(b.win_off, win_off.data(), sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && n_events > 0) {
    k_synth_events<<<(unsigned)((n_events + 255) / 256), 256, 0, s>>>(*sp, first_read, n, n_events, b.ev_off, b.col_off, b.start_clip,
                                                                      b.end_clip, b.contig, b.strand, start_d, b.ev_mean, b.ev_stdv,
                                                                      b.ev_len, b.ev_base, b.col_refbase, b.col_readbase, b.col_refpos);
    ctx->launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);          // the offset vectors above are pageable host memory
  release();
  if (e != cudaSuccess) { dm_set_error(ctx, std::string("dm_synth_generate: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  b.n_events = n_events; b.n_cols = n_cols; b.n_windows = n_windows; b.n_frows = n_frows;
  b.has_ev_base = true;
  b.from_alignment = false;
  b.prepared = false;
  ctx->h2d_bytes = 0;
  if (n_windows_out) *n_windows_out = n_windows;
  return DM_OK;
}

int dm_resident_sizes(dm_ctx* ctx, int32_t* n_reads, int64_t* n_events, int64_t* n_cols, int64_t* n_windows) {
  if (!ctx) return DM_ERR_ARG;
  if (n_reads) *n_reads = ctx->b.n_reads;
  if (n_events) *n_events = ctx->b.n_events;
  if (n_cols) *n_cols = ctx->b.n_cols;
  if (n_windows) *n_windows = ctx->b.n_windows;
  return DM_OK;
}

int dm_fetch_inputs(dm_ctx* ctx, int64_t* ev_off, float* ev_mean, float* ev_stdv, float* ev_len, uint8_t* ev_base,
                    int64_t* col_off, uint8_t* col_refbase, uint8_t* col_readbase, int64_t* col_refpos, int32_t* start_clip,
                    int32_t* end_clip, int32_t* contig, int8_t* strand) {
  if (!ctx) return DM_ERR_ARG;
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  const dm_dev_batch& b = ctx->b;
  cudaStream_t s = ctx->stream;
  const auto D2H = cudaMemcpyDeviceToHost;
  const size_t n = (size_t)b.n_reads, ne = (size_t)b.n_events, nc = (size_t)b.n_cols;
  if (n > 0) {
    if (ev_off) DM_CUDA(ctx, cudaMemcpyAsync(ev_off, b.ev_off, sizeof(int64_t) * (n + 1), D2H, s));
    if (col_off) DM_CUDA(ctx, cudaMemcpyAsync(col_off, b.col_off, sizeof(int64_t) * (n + 1), D2H, s));
    if (start_clip) DM_CUDA(ctx, cudaMemcpyAsync(start_clip, b.start_clip, sizeof(int32_t) * n, D2H, s));
    if (end_clip) DM_CUDA(ctx, cudaMemcpyAsync(end_clip, b.end_clip, sizeof(int32_t) * n, D2H, s));
    if (contig) DM_CUDA(ctx, cudaMemcpyAsync(contig, b.contig, sizeof(int32_t) * n, D2H, s));
    if (strand) DM_CUDA(ctx, cudaMemcpyAsync(strand, b.strand, n, D2H, s));
  }
  if (ne > 0) {
    if (ev_mean) DM_CUDA(ctx, cudaMemcpyAsync(ev_mean, b.ev_mean, sizeof(float) * ne, D2H, s));
    if (ev_stdv) DM_CUDA(ctx, cudaMemcpyAsync(ev_stdv, b.ev_stdv, sizeof(float) * ne, D2H, s));
    if (ev_len) DM_CUDA(ctx, cudaMemcpyAsync(ev_len, b.ev_len, sizeof(float) * ne, D2H, s));
    if (ev_base && b.has_ev_base) DM_CUDA(ctx, cudaMemcpyAsync(ev_base, b.ev_base, ne, D2H, s));
  }
  if (nc > 0) {
    if (col_refbase) DM_CUDA(ctx, cudaMemcpyAsync(col_refbase, b.col_refbase, nc, D2H, s));
    if (col_readbase) DM_CUDA(ctx, cudaMemcpyAsync(col_readbase, b.col_readbase, nc, D2H, s));
    if (col_refpos) DM_CUDA(ctx, cudaMemcpyAsync(col_refpos, b.col_refpos, sizeof(int64_t) * nc, D2H, s));
  }
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  return DM_OK;
}

}  // extern "C"
