// fp32 parity path of the BiLSTM: plain FFMA accumulation + accurate expf/tanhf.
//
// Restates the TF1 graph built by mCreateSession (bin/DeepMod_scripts/myMultiBiRNN.py:30-61)
// for the 66 live cell-steps (fw steps 0..10, bw steps 20..10, 3 layers each):
//     g = [inp, h] @ kernel + bias ; i,j,f,o = split(g)
//     c' = c*sigmoid(f+1) + sigmoid(i)*tanh(j) ; h' = tanh(c')*sigmoid(o)
//     logits = [h_fw2@10, h_bw2@10] @ Variable + Variable_1 ; softmax ; argmax
// This kernel is the <=1e-4 path (BASELINE configs[1]); throughput is the business of
// dm_lstm_tc.cu.  One CTA owns 64 windows and both directions; the hidden and cell state
// of all three layers stay in shared memory for the whole recurrence, the weights
// (1.6 MB, L2-resident) stream through L1.
//
// Thread (wg, ug) of the 8 x 25 grid owns windows wg*8..wg*8+7 and units {ug, ug+25, ug+50,
// ug+75} with all four gates: 128 accumulators, the cell update is thread-local.
#include "dm_common.cuh"

namespace {

constexpr int FW = 64;           // windows per CTA
constexpr int FT = 200;          // threads per CTA (8 window groups x 25 unit groups)
constexpr int WT = 8;            // windows per thread

struct __align__(16) Smem32 {
  float h[3][DM_HIDDEN][FW];     // hidden state, [layer][unit][window]
  float c[3][DM_HIDDEN][FW];     // cell state
  float xs[8][FW];               // features of the current step, [feature][window]
  float lg[2][FW];               // logits
  int frow[FW];                  // first feature row of each window
};

// acc[i][j*4+g] += sum_k act[k][wg*8+i] * W[k][j*100 + ug*4 + g]
__device__ __forceinline__ void gemm_part(float (&acc)[WT][16], const float* __restrict__ Wg,
                                          const float* __restrict__ act, int nrows) {
#pragma unroll 2
  for (int k = 0; k < nrows; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(act + k * FW);
    const float4 a1 = *reinterpret_cast<const float4*>(act + k * FW + 4);
    const float4* wr = reinterpret_cast<const float4*>(Wg + (size_t)k * DM_GATES);
    const float4 w0 = __ldg(wr), w1 = __ldg(wr + 25), w2 = __ldg(wr + 50), w3 = __ldg(wr + 75);
    const float a[WT] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                         w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
    for (int i = 0; i < WT; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(FT, 1)
k_lstm_fp32(const float* __restrict__ feat, const int32_t* __restrict__ win_frow, dm_dev_weights w,
            float* __restrict__ p1_out, uint8_t* __restrict__ pred_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem32& s = *reinterpret_cast<Smem32*>(smem_raw);
  const int tid = threadIdx.x;
  const int wg = tid / 25, ug = tid - wg * 25;
  const int64_t w0 = (int64_t)blockIdx.x * FW;

  if (tid < FW) s.frow[tid] = win_frow[w0 + tid];

  for (int dir = 0; dir < 2; ++dir) {
    // zero state (MultiRNNCellZeroState) and stage the first step's features
    for (int i = tid; i < 3 * DM_HIDDEN * FW; i += FT) {
      (&s.h[0][0][0])[i] = 0.f;
      (&s.c[0][0][0])[i] = 0.f;
    }
    __syncthreads();
    for (int i = tid; i < 8 * FW; i += FT) {
      const int win = i >> 3, f = i & 7;
      const int row = s.frow[win] + (dir == 0 ? 0 : DM_WINDOW - 1);
      s.xs[f][win] = feat[(int64_t)row * DM_FEAT_STRIDE + f];
    }
    __syncthreads();

    for (int step = 0; step < DM_LIVE_STEPS; ++step) {
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        float acc[WT][16];
#pragma unroll
        for (int i = 0; i < WT; ++i)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
        const float* Wl = w.w32[dir][l] + ug * 4;
        int n_in;
        if (l == 0) { gemm_part(acc, Wl, &s.xs[0][wg * WT], 8); n_in = 8; }
        else        { gemm_part(acc, Wl, &s.h[l - 1][0][wg * WT], DM_HIDDEN); n_in = DM_HIDDEN; }
        if (step > 0)    // h_prev == 0 at the first step of a window
          gemm_part(acc, Wl + (size_t)n_in * DM_GATES, &s.h[l][0][wg * WT], DM_HIDDEN);
        __syncthreads();     // every read of h[l] (old) and xs is done
        const float* bl = w.b32[dir][l];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int u = j * 25 + ug;
          const float4 b = __ldg(reinterpret_cast<const float4*>(bl) + j * 25 + ug);
          float* cp = &s.c[l][u][wg * WT];
          float* hp = &s.h[l][u][wg * WT];
          float cn[WT], hn[WT];
#pragma unroll
          for (int i = 0; i < WT; ++i) {
            const float gi = acc[i][j * 4 + 0] + b.x;
            const float gj = acc[i][j * 4 + 1] + b.y;
            const float gf = acc[i][j * 4 + 2] + b.z;
            const float go = acc[i][j * 4 + 3] + b.w;
            cn[i] = cp[i] * sigmoid_acc(gf + 1.0f) + sigmoid_acc(gi) * tanhf(gj);
            hn[i] = tanhf(cn[i]) * sigmoid_acc(go);
          }
#pragma unroll
          for (int i = 0; i < WT; i += 4) {
            *reinterpret_cast<float4*>(cp + i) = make_float4(cn[i], cn[i + 1], cn[i + 2], cn[i + 3]);
            *reinterpret_cast<float4*>(hp + i) = make_float4(hn[i], hn[i + 1], hn[i + 2], hn[i + 3]);
          }
        }
        if (l == 0 && step + 1 < DM_LIVE_STEPS) {
          const int t = dir == 0 ? step + 1 : DM_WINDOW - 2 - step;
          for (int i = tid; i < 8 * FW; i += FT) {
            const int win = i >> 3, f = i & 7;
            s.xs[f][win] = feat[(int64_t)(s.frow[win] + t) * DM_FEAT_STRIDE + f];
          }
        }
        __syncthreads();     // h[l] (new) and xs (next step) visible
      }
    }
    // classifier partial sums of this direction: rows dir*100 .. dir*100+99 of `Variable`
    if (tid < 2 * FW) {
      const int win = tid & (FW - 1), cls = tid >> 6;
      float acc = 0.f;
      for (int u = 0; u < DM_HIDDEN; ++u)
        acc = fmaf(s.h[2][u][win], __ldg(&w.cls_w[(dir * DM_HIDDEN + u) * 2 + cls]), acc);
      if (dir == 0) s.lg[cls][win] = acc; else s.lg[cls][win] += acc;
    }
    __syncthreads();
  }
  if (tid < FW) {
    const float l0 = s.lg[0][tid] + __ldg(&w.cls_b[0]);
    const float l1 = s.lg[1][tid] + __ldg(&w.cls_b[1]);
    const float m = fmaxf(l0, l1);                       // tf.nn.softmax, myMultiBiRNN.py:59
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    const float inv = 1.0f / (e0 + e1);
    const float q0 = e0 * inv, q1 = e1 * inv;
    if (p1_out) p1_out[w0 + tid] = q1;
    if (pred_out) pred_out[w0 + tid] = q1 > q0 ? 1 : 0;  // tf.argmax: first maximum wins
  }
}

}  // namespace

int dm_launch_lstm_fp32(dm_ctx* ctx, const float* feat, const int32_t* win_frow, int64_t n_windows,
                        float* p1, uint8_t* pred) {
  const int64_t n_pad = dm_pad_windows(n_windows);
  if (n_pad == 0) return DM_OK;
  if (!ctx->fp32_attr_set) {
    DM_CUDA(ctx, cudaFuncSetAttribute(k_lstm_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(Smem32)));
    ctx->fp32_attr_set = true;
  }
  k_lstm_fp32<<<(unsigned)(n_pad / FW), FT, sizeof(Smem32), ctx->stream>>>(feat, win_frow, ctx->w, p1, pred);
  ctx->launches += 1;
  DM_CUDA(ctx, cudaGetLastError());
  return DM_OK;
}
