// Tensor-core path of the BiLSTM (tcgen05 / TMEM / bulk-TMA; fp16 or bf16 operands, fp32 accumulation), sm_100a only.
//
// Same graph as dm_lstm_fp32.cu (bin/DeepMod_scripts/myMultiBiRNN.py:30-61, 66 live
// cell-steps), restructured for the B200:
//
//  * one CTA owns 128 windows (= the 128 TMEM lanes / UMMA M) for both directions;
//  * every cell-step is ONE logical GEMM  gates[128,400] = A[128,K] * W[K,400]  issued as
//    5 N-chunks of 80 gate columns (20 units x {i,j,f,o}) into a TMEM ring of 5 slots (one per
//    chunk: static addresses in the unrolled chunk loops), so the tensor pipe fills the slots of
//    the next cell-step while the epilogue warps drain those of this one;
//  * A is never materialised per step: it is the concatenation of the resident 16-bit hidden
//    tiles (K-major core-matrix columns of 128 rows x 16 B).  Each hidden tile carries 4
//    extra K slots (1, 1, mean_lo, stdv_lo): the bias rides in the GEMM as a hi/lo
//    pair against the constant ones, the signal features as hi/lo pairs;
//  * the three layers are walked in wavefront order (t+l = const), which makes consecutive
//    cell-steps independent: MMA of step g+1 overlaps the epilogue of step g.  h0/h1 are
//    double-buffered on the parity of t for that;
//  * weights (0.5 folded into the sigmoid gates and into every row against a hidden state, forget
//    bias folded into the bias) stream from L2 through a 5-stage shared-memory ring with cp.async.bulk + mbarriers,
//    pre-arranged on the host in the exact UMMA canonical (no-swizzle, K-major) layout;
//  * the epilogue (20 warps) reads 16 TMEM columns = 4 units per thread and chunk, packs the gate
//    pre-activations of TWO units into f16x2 and applies sigmoid(x) = 0.5*tanh(x/2)+0.5 and tanh
//    with one tanh.approx.f16x2 per pair (which is still two MUFU.TANH.F16 in SASS: the MUFU pipe retires
//    16.5 results per clock and SM whatever the format, and it is what binds this kernel), does the cell update in
//    packed half2 FMAs (cell state in packed fp16 registers) and writes 2h straight into the next
//    A operand (the factor 0.5 of h = 0.5*(tanh(c)*tanh(o/2) + tanh(c)) is folded into the weights);
//  * operands are fp16 (DM_F16: h needs no conversion at all) or bf16 (DM_BF16), fp32 accumulate.
#include "dm_common.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

namespace {

constexpr int TC_CTRL_WARPS = 4;        // warps 20..23: weight producer, MMA issuer, TMEM alloc, pair relay
constexpr int TC_EPI_WARPS = 20;        // warps 0..19: 4 lane quarters x 5 column groups
// The control warps carry the HIGHEST warp ids on purpose: the SM's issue arbiter favours high warp
// ids, and a starved MMA issuer / weight producer stalls all 20 epilogue warps.
constexpr int W_PROD = TC_EPI_WARPS, W_MMA = TC_EPI_WARPS + 1, W_ALLOC = TC_EPI_WARPS + 2, W_RELAY = TC_EPI_WARPS + 3;
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = (TC_CTRL_WARPS + TC_EPI_WARPS) * 32;
constexpr int TC_CHUNK_N = 80;          // gate columns per MMA and TMEM slot
constexpr int TC_NCHUNK = 5;
#ifndef TC_TSLOTS_N
#define TC_TSLOTS_N 5      // accumulator ring depth (slots of 80 TMEM columns).  5 = one slot per N-chunk of a cell-step: slot
                           // indices, TMEM addresses and barrier addresses become compile-time constants in the unrolled
                           // chunk loops (less uniform-datapath bookkeeping), at the price of one slot of run-ahead:
                           // 5 is 1.7 % faster than 6 (profiles/r02_variants_sweep2.txt)
#endif
constexpr int TC_TSLOTS = TC_TSLOTS_N;
constexpr bool TC_STATIC_SLOTS = TC_TSLOTS == 5;          // == TC_NCHUNK (defined below)
constexpr int TC_ACOL = 2048;           // A core column: 128 rows x 16 B
constexpr int TC_HCOLS = 13;            // 100 units + 4 extra K slots
constexpr int TC_HTILE = TC_HCOLS * TC_ACOL;
constexpr int TC_STEPS_PER_DIR = 33;
constexpr int TS_BYTES = 8 * 8192;       // debug timeline behind the operand dump
// first of the four hidden units whose gates sit in accumulator columns 16 s .. 16 s + 15 of chunk j
__host__ __device__ constexpr int tc_unit0(int j, int s);
#ifndef TC_PREFETCH
#define TC_PREFETCH 2      // 1: the accumulator load of chunk j+1 is issued before the stores of chunk j
                           // 2: also across cell-steps wherever the next step's MMAs cannot depend on this one
#endif
#ifndef TC_T0SKIP
#define TC_T0SKIP 1        // 1: no forget-gate tanh at t == 0 (c_prev == 0); 0: no special case (one branch less per chunk)
#endif
#ifndef TC_UNIWARP
#define TC_UNIWARP 1       // warp index through a shuffle (provably warp-uniform for the compiler)
#endif
#ifndef TC_GATE32
#define TC_GATE32 0        // 1: the four gate tanh are taken in f32 straight from the accumulator registers and their RESULTS
                           //    are packed to f16x2 (8 MUFU.TANH + 4 F2FP per unit pair instead of 4 F2FP + 8 MUFU.TANH.F16 + 4 PRMT):
                           //    the gate inputs keep their fp32 precision (mean |dp1| 1.6e-4 instead of 2.8e-4, flip rate 1.2e-4
                           //    instead of 2.4e-4) and short kernels run at the same speed, but the build spills more (160 B) and
                           //    draws more power: in a long run the part settles at 1695 instead of 1747 MHz under its 1000 W cap
                           //    and delivers 96.5 instead of 99.3 Mbases/s (profiles/r02_sustained_ab.txt).  Off: throughput first.
#endif
#ifndef TC_FMAF
#define TC_FMAF 0          // 1: the forget gate's sigmoid is evaluated on the FMA / ALU pipes in packed fp16 (exponent arithmetic
                           //    + two small polynomials, |error| <= 1.0e-3) instead of on the MUFU pipe, which binds this kernel:
                           //    16 instead of 20 MUFU per unit quad.  The f-gate weight columns are then scaled by log2(e)
                           //    instead of 0.5.  tools/numerics_study.py: no measurable change of |dp1| (gates i and f tolerate it).
                           //    MEASURED (profiles/r02_variants_sweep3.txt, r02_fma_offload_ncu.txt): the MUFU pipe drops from 86 % to
                           //    67 % busy and the kernel is NOT faster (-3 % with TC_PARK, -10 % without: spills): 35 % more
                           //    instructions put the issue slots at 67 %, with six warps per scheduler that is the new limit.  Off.
#endif
#ifndef TC_FMAI
#define TC_FMAI 0          // 1: the same for the input gate's sigmoid (12 MUFU per unit quad; measure before use: the FMA pipe
                           //    then carries about as much as the MUFU pipe)
#endif
#ifndef TC_SETMAXNREG
#define TC_SETMAXNREG 0    // 1: the control warpgroup (warps 20..23) shrinks to 40 registers per thread and the five epilogue
                           //    warpgroups grow to 88 (640 x 88 + 128 x 40 = 768 x 80, the launch allocation)
#endif
#ifndef TC_PARK
#define TC_PARK 0          // 1: only the cell state of the layer being stepped lives in registers (10 per thread); the other two
                           //    layers' states are parked in the 112 TMEM columns the 5-slot accumulator ring leaves free
                           //    (tcgen05.st / tcgen05.ld once per cell-step).  20 registers back: no spills in the chunk loops,
                           //    but the swap at every step start costs more than the spills did (-2.7 %).  Off.
#endif
#ifndef TC_EARLYLD
#define TC_EARLYLD 0       // when the next chunk's accumulator load is issued (its 16 registers are free as soon as the
                           // gate columns are packed to f16x2): 1 = right after the packing (waits for its barrier
                           // there), 2 = there if a non-blocking probe finds it ready, else after the cell update,
                           // 0 = always after the cell update
#endif
#ifndef TC_T0TRIM
#define TC_T0TRIM 1        // 1: at t == 0 (h_prev == 0) the MMAs over the own hidden tile are not issued, and a direction
                           //    re-initialisation zeroes only the three core columns those steps still read
#endif
#ifndef TC_UNITMAP
#define TC_UNITMAP 0       // 1: column group s owns units 20 s .. 20 s + 19 (4 per chunk): consecutive chunks complete 16-byte
                           //    row segments of the hidden tile, which go out as conflict-free 16-byte stores
                           // 0: chunk j owns units 20 j .. 20 j + 19 (4 per column group): 8-byte stores, 2-way bank conflict
                           // (measured: 1 is bit-identical and 1.9 % SLOWER despite 12 instead of 20 store wavefronts per
                           //  warp and step - kept as a switch, off)
#endif
#ifndef TC_INTERLEAVE
#define TC_INTERLEAVE 1    // 1: the last two cell-steps of a direction are interleaved with the first two of the next one
                           //    (next direction or next tile), so that NO cell-step depends on its predecessor
#endif
#ifndef TC_SKEW
#define TC_SKEW 0          // SM clocks by which consecutive column groups start a direction later (breaks lockstep)
#endif

__host__ __device__ constexpr int tc_unit0(int j, int s) { return TC_UNITMAP ? 20 * s + 4 * j : 20 * j + 4 * s; }

// Geometry of the weight stream: one stage = the B operand of one N-chunk (all K core columns).
// PAIR = two CTAs of a cluster drive ONE tcgen05.mma.cta_group::2 (M = 256, 128 windows each): every
// CTA stages only its half of the gate columns (40 of the chunk's 80), so the per-SM L2 stream halves
// and the ring holds a whole cell-step ahead of the tensor pipe.
template <bool PAIR> struct Geo {
  static constexpr int BCOL = (PAIR ? TC_CHUNK_N / 2 : TC_CHUNK_N) * 16;   // bytes of one B core column per CTA
  static constexpr int STAGE = 26 * BCOL;                                  // 16640 / 33280
  static constexpr int NSTAGE = PAIR ? 5 : 2;
  static constexpr int NCTA = PAIR ? 2 : 1;
};
constexpr int TC_RING = 5 * 26 * 640;    // 83200 >= 2 * 33280

constexpr int OFF_X = 0;                                   // 2 x-columns
constexpr int OFF_H0 = OFF_X + 2 * TC_ACOL;                // h0[2]
constexpr int OFF_H1 = OFF_H0 + 2 * TC_HTILE;              // h1[2]
constexpr int OFF_H2 = OFF_H1 + 2 * TC_HTILE;              // h2
constexpr int OFF_W = OFF_H2 + TC_HTILE;                   // weight ring
constexpr int OFF_BAR = OFF_W + TC_RING;
constexpr int BAR_FULL = 0, BAR_EMPTY = 5, BAR_TFULL = 10, BAR_TEMPTY = 16, BAR_HDONE = 22, N_BARS = 24;
constexpr int OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr int OFF_PART = OFF_TMEM + 16;                    // float part[5][128]
constexpr int OFF_FROW = OFF_PART + 5 * 128 * 4;           // int frow[128]
constexpr int OFF_CLS = OFF_FROW + (TC_INTERLEAVE ? 2 : 1) * 128 * 4;   // float cls_d[2][100]  (frow is per tile parity when interleaved)
constexpr int TC_SMEM = OFF_CLS + 200 * 4;
static_assert(TC_SMEM <= 227 * 1024, "shared memory budget");

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// try_wait suspends the thread in hardware (up to the hint, in ns) instead of spinning on the issue port
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}"
      ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
// non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_alloc(uint32_t smem_dst) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_dst) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr) {    // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M = 128
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld2(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tc_st2(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(v[0]), "r"(v[1]) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// same, with the loaded registers as in/out operands: no consumer of v can be scheduled above the wait
__device__ __forceinline__ void tc_wait_ld16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :: "memory");
}
// ---- CTA-pair (cta_group::2) variants ----
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}"
      ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {     // both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_alloc2(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_dst) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// one lane of the (converged) warp; the same lane every time for the full mask
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory"); }

// UMMA shared-memory descriptor, K-major, no swizzle: 8-row core matrices of 128 contiguous
// bytes; `lbo` = byte distance between the two K-adjacent core matrices of one K=16 step,
// `sbo` = byte distance between 8-row groups.  (cute::UMMA::SmemDescriptor, version 1.)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D = f32, A = B = bf16 (format 1) or fp16 (format 0), both K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n, bool f16 = false) {
  return (1u << 4) | (f16 ? 0u : (1u << 7) | (1u << 10)) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float tanh32_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two results per PTX instruction - but two MUFU.TANH.F16 + one PRMT in SASS (there is no packed MUFU on sm_100)
__device__ __forceinline__ __half2 tanh2_mufu(__half2 x) {
  uint32_t y;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(*reinterpret_cast<const uint32_t*>(&x)));
  return *reinterpret_cast<__half2*>(&y);
}
// sigmoid(x) for two values at once, WITHOUT the MUFU pipe.  u = x * log2(e) (the GEMM delivers it: weight scaling).
//   a = min(|u|, 14.5);  r = round(a) by the mantissa trick (1039 - a has an ulp of 1);  e = 2^-a = 2^-(a - r) * 2^-r:
//   a quadratic in the fraction times an exponent built from the low bits of (1039 - a);  1 / (1 + e) - 1/2 as a quartic in
//   e on [0, 1];  the sign of u goes back on.  12 FMA-pipe + 5 ALU-pipe instructions for two results; max error 1.0e-3.
__device__ __forceinline__ __half2 sigmoid2_fma(__half2 u) {
  const __half2 kM = __floats2half2_rn(1039.f, 1039.f), kA = __floats2half2_rn(14.5f, 14.5f);
  const __half2 a = __hmin2(__habs2(u), kA);
  const __half2 tm = __hsub2(kM, a);
  const __half2 fr = __hadd2(a, __hsub2(tm, kM));                    // a - round(a), in [-0.5, 0.5]
  // 2^-fr  (fp16-rounded least-squares coefficients on Chebyshev nodes)
  __half2 p = __hfma2(__floats2half2_rn(0.24267578f, 0.24267578f), fr, __floats2half2_rn(-0.7036133f, -0.7036133f));
  p = __hfma2(p, fr, __floats2half2_rn(1.f, 1.f));
  const uint32_t sbits = (*reinterpret_cast<const uint32_t*>(&tm) << 10) & 0x7C007C00u;      // 2^-round(a) (0 when round(a) = 15)
  const __half2 e = __hmul2(p, *reinterpret_cast<const __half2*>(&sbits));
  __half2 q = __hfma2(__floats2half2_rn(0.15686035f, 0.15686035f), e, __floats2half2_rn(-0.54248047f, -0.54248047f));
  q = __hfma2(q, e, __floats2half2_rn(0.8720703f, 0.8720703f));
  q = __hfma2(q, e, __floats2half2_rn(-0.9863281f, -0.9863281f));
  q = __hfma2(q, e, __floats2half2_rn(0.49975586f, 0.49975586f));   // 1 / (1 + e) - 1/2 >= 0
  const uint32_t sq = (*reinterpret_cast<const uint32_t*>(&u) & 0x80008000u) | *reinterpret_cast<const uint32_t*>(&q);
  return __hadd2(__floats2half2_rn(0.5f, 0.5f), *reinterpret_cast<const __half2*>(&sq));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---- TC_INTERLEAVE: the global order of cell-steps ----
// A direction instance (tile, direction) is the sequence S[0..32] of (d, l) in wavefront order.  Within it only S[1]
// depends on S[0] and S[32] on S[31] at distance one.  Order over instances i = 0..N-1:
//   S_0[0..30] | S_1[0] S_0[31] S_1[1] S_0[32] S_1[2..30] | S_2[0] S_1[31] S_2[1] S_1[32] S_2[2..30] | ... | S_{N-1}[31] S_{N-1}[32]
// Every step then reads what steps two or more back have written (except the 2nd and the last step of a CTA's run).
__device__ __forceinline__ void il_decode(int G, int N, int& inst, int& sidx) {
  if (G < 31) { inst = 0; sidx = G; return; }
  const int q = (G - 31) / 33, off = (G - 31) - q * 33, i = q + 1;
  if (i >= N) { inst = N - 1; sidx = 31 + off; return; }
  if (off == 0) { inst = i; sidx = 0; }
  else if (off == 1) { inst = i - 1; sidx = 31; }
  else if (off == 2) { inst = i; sidx = 1; }
  else if (off == 3) { inst = i - 1; sidx = 32; }
  else { inst = i; sidx = off - 2; }
}
__device__ __forceinline__ void il_step(int sidx, int& d, int& l) {
  if (sidx == 0) { d = 0; l = 0; }
  else if (sidx < 3) { d = 1; l = sidx - 1; }
  else if (sidx < 30) { d = 2 + (sidx - 3) / 3; l = (sidx - 3) % 3; }
  else if (sidx == 30) { d = 11; l = 1; }
  else if (sidx == 31) { d = 11; l = 2; }
  else { d = 12; l = 2; }
}

// feature row of time index tau (0..10 in processing order) of a window
__device__ __forceinline__ int tau_row(int dir, int tau) { return dir == 0 ? tau : DM_WINDOW - 1 - tau; }

// ---- the kernel -----------------------------------------------------------------------------
// Roles (warp ids): 0..19 epilogue; 20 weight producer; 21 MMA issuer (leader CTA) / accumulator-
// drained relay (peer CTA); 22 TMEM allocator / hidden-state-written relay (peer); 23 stage-landed
// relay (peer).  The relays exist because a cluster-scope release-arrive costs ~1000 cycles: the
// peer's 20 epilogue warps arrive on cheap CTA-local barriers and ONE thread forwards each phase.
template <bool PAIR, bool DBG, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_lstm_tc(const uint16_t* __restrict__ feat_tc, const int32_t* __restrict__ win_frow, dm_dev_weights w,
          float* __restrict__ p1_out, uint8_t* __restrict__ pred_out, int n_tiles, int max_steps,
          unsigned char* __restrict__ dbg) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  using G = Geo<PAIR>;
  // DBG instantiation only (dm_debug_tc_windows): development switches ride in the upper bits of max_steps
  //   0x100 epilogue skips the gate math, 0x200 no tcgen05.mma is issued, 0x400 no weight loads
  // and a timeline of CTA 0 is written behind the operand dump: [role][step][slot] SM clocks.
  const bool dbg_nomath = DBG && (max_steps & 0x100) != 0, dbg_nomma = DBG && (max_steps & 0x200) != 0,
             dbg_noload = DBG && (max_steps & 0x400) != 0;
  max_steps = DBG ? (max_steps & 0xFF) : 2 * TC_STEPS_PER_DIR;
  unsigned long long* ts = (DBG && dbg != nullptr && blockIdx.x == 0) ? reinterpret_cast<unsigned long long*>(dbg + OFF_W) : nullptr;
#define TS(idx) do { if (DBG && ts) ts[(idx)] = clock64(); } while (0)
  // the warp index goes through a shuffle so that the compiler KNOWS it is warp-uniform: role branches are then
  // uniform and per-warp quantities (column group, descriptors, barrier addresses) can live in uniform registers
  const int tid = threadIdx.x, warp = TC_UNIWARP ? __shfl_sync(0xffffffffu, tid >> 5, 0) : tid >> 5, lane = tid & 31;
  const uint32_t crank = PAIR ? cluster_rank() : 0u;      // 0 = leader (issues the MMAs of the pair)
  const bool leader = crank == 0;
  // persistent: this CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (a pair walks tile pairs);
  // barriers, TMEM and the weight ring live across tiles, only the hidden state is re-initialised
  const int n_iter = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  int* s_frow = reinterpret_cast<int*>(smem + OFF_FROW);
  float* s_part = reinterpret_cast<float*>(smem + OFF_PART);
  float* s_cls = reinterpret_cast<float*>(smem + OFF_CLS);
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  // ---- one-time setup ----
  if (tid < DM_TILE_M) s_frow[tid] = win_frow[(int64_t)blockIdx.x * DM_TILE_M + tid];
  if (tid < 200) s_cls[tid] = w.cls_d[tid];
  if (tid == 0) {
    // leader barriers also collect one forwarded arrival per phase from the peer CTA
    const uint32_t extra = (PAIR && leader) ? 1 : 0;
    for (int i = 0; i < G::NSTAGE; ++i) {
      mbar_init(bar0 + 8 * (BAR_FULL + i), 1 + extra);
      mbar_init(bar0 + 8 * (BAR_EMPTY + i), 1);
    }
    for (int i = 0; i < TC_TSLOTS; ++i) {
      mbar_init(bar0 + 8 * (BAR_TFULL + i), 1);
      mbar_init(bar0 + 8 * (BAR_TEMPTY + i), TC_EPI_WARPS + extra);
    }
    mbar_init(bar0 + 8 * (BAR_HDONE + 0), TC_EPI_WARPS + extra);
    mbar_init(bar0 + 8 * (BAR_HDONE + 1), TC_EPI_WARPS + extra);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ALLOC) { if (PAIR) tc_alloc2(sbase + OFF_TMEM); else tc_alloc(sbase + OFF_TMEM); }
  __syncthreads();     // frow visible to the epilogue's init below

  // epilogue-side coordinates (meaningful for warps < 20)
  const int q = warp & 3;                    // TMEM lane quarter this warp may touch
  const int sgrp = warp >> 2;                // column group 0..4 (4 units per chunk)
  const int row = q * 32 + lane;             // window within the tile
  const uint32_t row_off = (uint32_t)((row >> 3) * 128 + (row & 7) * 16);

  // direction init: zero the hidden tiles, stage x(0), x(1) and the step-0 extras of h0[1]
  auto dir_init = [&](int dir, bool full) {
    const int et = tid;
    // the global loads go first so that their latency hides behind the zeroing and the barrier
    uint4 v = make_uint4(0, 0, 0, 0);
    if (et < 256)
      v = *reinterpret_cast<const uint4*>(feat_tc + ((int64_t)s_frow[et & 127] + tau_row(dir, et >> 7)) * 16);
    else if (et < 384) {
      const uint2 e = *reinterpret_cast<const uint2*>(feat_tc + ((int64_t)s_frow[et - 256] + tau_row(dir, 0)) * 16 + 8);
      v.x = e.x; v.y = e.y;
    }
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (full || !TC_T0TRIM) {
      // (the very first time also for what no step writes before it is multiplied by a zero weight: unused K slots)
      for (int i = et; i < (5 * TC_HTILE) / 16; i += TC_EPI_THREADS)
        *reinterpret_cast<uint4*>(smem + OFF_H0 + i * 16) = z;
    } else if (et < 384) {
      // every tile is rewritten by this direction's own steps before it is read, except what the t = 0 steps read of
      // h(-1): units 96..99 next to layer 0's step-0 extras (h0[1] column 12) and the column that shares a K = 16 MMA
      // with the last column of the tile below (h1[1] column 0, h2 column 0)
      const int c = et >> 7;
      const uint32_t col = c == 0 ? OFF_H0 + TC_HTILE + 12 * TC_ACOL : c == 1 ? OFF_H1 + TC_HTILE : OFF_H2;
      *reinterpret_cast<uint4*>(smem + col + (et & 127) * 16) = z;
    }
    epi_bar();
    if (et < 256) {
      const int r = et & 127, tau = et >> 7;
      *reinterpret_cast<uint4*>(smem + OFF_X + tau * TC_ACOL + (r >> 3) * 128 + (r & 7) * 16) = v;
    } else if (et < 384) {
      const int r = et - 256;
      *reinterpret_cast<uint2*>(smem + OFF_H0 + TC_HTILE + 12 * TC_ACOL + (r >> 3) * 128 + (r & 7) * 16 + 8) =
          make_uint2(v.x, v.y);
    }
    fence_async_smem();
    epi_bar();
    if (TC_SKEW > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)sgrp * TC_SKEW) { }
    }
  };
  if (warp < TC_EPI_WARPS) {
    for (int i = tid; i < 5 * 128; i += TC_EPI_THREADS) s_part[i] = 0.f;
    dir_init(0, true);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();       // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
#if TC_SETMAXNREG
  if (warp >= TC_EPI_WARPS) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
#endif

// walks the 66 cell-steps in wavefront order; BODY sees dir, d, l, t, g
#if TC_INTERLEAVE
#define FOR_EACH_STEP(...)                                    \
  for (int GG = 0; GG < 2 * TC_STEPS_PER_DIR * n_iter; ++GG) { \
    int inst, sidx, d, l;                                     \
    il_decode(GG, 2 * n_iter, inst, sidx);                    \
    il_step(sidx, d, l);                                      \
    const int dir = inst & 1, t = d - l, g = GG, G0 = 0;      \
    (void)G0;                                                 \
    { __VA_ARGS__ } }
#else
#define FOR_EACH_STEP(...)                                    \
  for (int it = 0; it < n_iter; ++it) {                       \
    int g = 0;                                                \
    const int G0 = it * 2 * TC_STEPS_PER_DIR;                 \
    (void)G0;                                                 \
    for (int dir = 0; dir < 2; ++dir)                         \
      for (int d = 0; d < 13; ++d)                            \
        _Pragma("unroll") for (int l = 0; l < 3; ++l) {       \
          const int t = d - l;                                \
          if (t < 0 || t > 10) continue;                      \
          if (g < max_steps) { __VA_ARGS__ }                  \
          ++g;                                                \
        } }
#endif

  if (warp == W_PROD) {
    // ================= weight producer: one stage = one N-chunk of this CTA's gate columns =================
    if (lane == 0 && !dbg_noload) {
      uint32_t slot = 0, use = 0;
      FOR_EACH_STEP({
        const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(PAIR ? w.wtc2[F16][dir][l] : w.wtc[F16][dir][l]);
        const int ncols = l == 0 ? 14 : 26;
        const uint32_t bytes = ncols * G::BCOL;
        for (int j = 0; j < TC_NCHUNK; ++j) {
          if (use > 0) mbar_wait(bar0 + 8 * (BAR_EMPTY + slot), (use - 1) & 1);
          mbar_expect_tx(bar0 + 8 * (BAR_FULL + slot), bytes);
          // image: [chunk j][cta r][K core column][n][8]
          bulk_g2s(sbase + OFF_W + slot * G::STAGE, wsrc + (size_t)((j * G::NCTA + crank) * ncols) * G::BCOL, bytes,
                   bar0 + 8 * (BAR_FULL + slot));
          if (++slot == G::NSTAGE) { slot = 0; ++use; }
        }
      })
    }
  } else if (PAIR && !leader && warp >= W_MMA) {
    // ================= peer-side relays: forward one arrival per phase to the leader =================
    if (lane == 0) {
      if (warp == W_RELAY) {                    // "my half of the stage has landed"
        if (!dbg_noload) {
          uint32_t slot = 0, use = 0;
          FOR_EACH_STEP({
            for (int j = 0; j < TC_NCHUNK; ++j) {
              mbar_wait(bar0 + 8 * (BAR_FULL + slot), use & 1);
              mbar_arrive_remote(bar0 + 8 * (BAR_FULL + slot), 0);
              if (++slot == G::NSTAGE) { slot = 0; ++use; }
            }
          })
        }
      } else if (warp == W_MMA) {               // "my epilogue has drained accumulator slot"
        uint32_t tslot = 0, tuse = 0;
        FOR_EACH_STEP({
          if (TC_STATIC_SLOTS) tslot = 0;
          for (int j = 0; j < TC_NCHUNK; ++j) {
            mbar_wait(bar0 + 8 * (BAR_TEMPTY + tslot), tuse & 1);
            mbar_arrive_remote(bar0 + 8 * (BAR_TEMPTY + tslot), 0);
            if (++tslot == TC_TSLOTS) { tslot = 0; ++tuse; }
          }
        })
      } else {                                  // "my epilogue has written the hidden state of step g"
        FOR_EACH_STEP({
          const int G = G0 + g;
          mbar_wait(bar0 + 8 * (BAR_HDONE + (G & 1)), (G >> 1) & 1);
          mbar_arrive_remote(bar0 + 8 * (BAR_HDONE + (G & 1)), 0);
        })
      }
    }
  } else if ((warp == W_MMA || warp == W_RELAY) && leader) {
    // ================= MMA issuers =================
    // Issuing is single-threaded and latency-bound (two mbarrier waits + 13 tcgen05.mma + two commits per
    // N-chunk), so TWO warps take alternate chunks; chunks own disjoint accumulator slots and weight
    // stages, and each issuer's commits cover exactly its own MMAs.  The whole warp walks the loop
    // converged and one elected lane issues: descriptors then live in uniform registers (a loop under
    // `if (lane == 0)` makes the compiler broadcast every operand of every MMA through a waterfall loop).
    {
      const uint32_t mine = warp == W_MMA ? 0u : 1u;
      constexpr uint32_t idesc = umma_idesc(PAIR ? 256 : 128, TC_CHUNK_N, F16);
      constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);         // SBO = 128 B, descriptor version 1
      constexpr uint32_t b_step = (2 * G::BCOL) >> 4;                // two K core columns per MMA
      uint32_t slot = 0, use = 0, tslot = 0, tuse = 0, c = 0;
      FOR_EACH_STEP({
        if (TC_STATIC_SLOTS) tslot = 0;                       // (it is 0 already: said so for constant folding)
        if (G::NSTAGE == TC_NCHUNK) slot = 0;
        if (mine == 0 && lane == 0) TS(g * 8 + 0);
        // inputs of this cell-step were written by the epilogues of steps <= g-2, or g-1 in the
        // fill/drain corners of the wavefront (and across the direction switch)
        const int G = G0 + g;
        if (G >= 2) mbar_wait(bar0 + 8 * (BAR_HDONE + (G & 1)), ((G - 2) >> 1) & 1);
#if TC_INTERLEAVE
        const bool wait1 = G == 1 || G == 2 * TC_STEPS_PER_DIR * n_iter - 1;     // the only distance-one dependencies left
#else
        const bool wait1 = d == 0 || (d == 1 && l == 0) || d == 12;     // d == 0: first step after a (re-)initialisation
#endif
        if (wait1 && G >= 1) mbar_wait(bar0 + 8 * (BAR_HDONE + ((G - 1) & 1)), ((G - 1) >> 1) & 1);
        tc_fence_after();
        if (mine == 0 && lane == 0) TS(g * 8 + 1);
        // A operand = ascending chain of core columns [n0 columns at base0 | rest at base1]; one
        // descriptor (low word) per K=16 step, shared by the five N-chunks of the step
        uint32_t base0, base1;
        const int n0 = l == 0 ? 1 : TC_HCOLS;
        // t == 0: h(-1) == 0, the K steps over the own hidden tile drop out.  Layers 1, 2 keep the first 7 (the 7th
        // pairs the last column of the tile below with own column 0, which is zero); layer 0 keeps ONE: the x
        // column paired with own column 12 (bias carriers and the low feature halves)
        const bool t0 = TC_T0TRIM && t == 0;
        const int nk16 = l == 0 ? (t0 ? 1 : 7) : (t0 ? 7 : 13);
        if (l == 0)      { base0 = sbase + OFF_X + (t & 1) * TC_ACOL;   base1 = sbase + OFF_H0 + ((t + 1) & 1) * TC_HTILE; }
        else if (l == 1) { base0 = sbase + OFF_H0 + (t & 1) * TC_HTILE; base1 = sbase + OFF_H1 + ((t + 1) & 1) * TC_HTILE; }
        else             { base0 = sbase + OFF_H1 + (t & 1) * TC_HTILE; base1 = sbase + OFF_H2; }
        uint32_t a_lo[13];
        _Pragma("unroll")
        for (int i = 0; i < 13; ++i) {
          if (i >= nk16) break;
          const int v0 = 2 * i, v1 = v0 + 1;
          const uint32_t a0 = v0 < n0 ? base0 + v0 * TC_ACOL : base1 + (v0 - n0) * TC_ACOL;
          const uint32_t a1 = v1 < n0 ? base0 + v1 * TC_ACOL : base1 + (v1 - n0) * TC_ACOL;
          a_lo[i] = ((a0 >> 4) & 0x3FFFu) | (((a1 - a0) >> 4) << 16);
        }
        if (t0 && l == 0) a_lo[0] = ((base0 >> 4) & 0x3FFFu) | (((base1 + 12 * TC_ACOL - base0) >> 4) << 16);
        for (int j = 0; j < TC_NCHUNK; ++j, ++c) {
          if ((c & 1u) == mine) {
            if (tuse > 0) mbar_wait(bar0 + 8 * (BAR_TEMPTY + tslot), (tuse - 1) & 1);
            if (!dbg_noload) mbar_wait(bar0 + 8 * (BAR_FULL + slot), use & 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + tslot * TC_CHUNK_N;
            // (layer 0 at t == 0: weight K columns 0 and 13 of the stage)
            const uint32_t b_lo = (((sbase + OFF_W + slot * G::STAGE) >> 4) & 0x3FFFu) |
                                  ((uint32_t)(((t0 && l == 0) ? 13 * G::BCOL : G::BCOL) >> 4) << 16);
            if (!dbg_nomma) {
              _Pragma("unroll")
              for (int i = 0; i < 13; ++i) {
                if (i >= nk16) break;
                const uint64_t ad = ((uint64_t)desc_hi << 32) | a_lo[i];
                const uint64_t bd = ((uint64_t)desc_hi << 32) | (b_lo + i * b_step);
                if (elect_one()) {
                  if (PAIR) tc_mma2(d_tmem, ad, bd, idesc, i > 0 ? 1u : 0u);
                  else tc_mma(d_tmem, ad, bd, idesc, i > 0 ? 1u : 0u);
                }
              }
            }
            // weight stage reusable / accumulator chunk ready once these MMAs retire (in both CTAs of a pair)
            if (elect_one()) {
              if (PAIR) { tc_commit2(bar0 + 8 * (BAR_EMPTY + slot)); tc_commit2(bar0 + 8 * (BAR_TFULL + tslot)); }
              else      { tc_commit(bar0 + 8 * (BAR_EMPTY + slot));  tc_commit(bar0 + 8 * (BAR_TFULL + tslot)); }
            }
            if (lane == 0 && (mine == 0 || j == 1 || j == 3)) TS(g * 8 + 2 + j);
          }
          if (++slot == G::NSTAGE) { slot = 0; ++use; }
          if (++tslot == TC_TSLOTS) { tslot = 0; ++tuse; }
        }
      })
    }
  } else if (warp < TC_EPI_WARPS) {
    // ================= epilogue: gates -> (c, h) =================
#if TC_PARK
    static_assert(TC_STATIC_SLOTS && TC_INTERLEAVE && TC_T0SKIP, "TC_PARK needs the 5-slot ring (free TMEM columns), the interleaved order and the t = 0 shortcut");
    __half2 cst[1][TC_NCHUNK][2];       // the layer in registers; (park_cur, park_slot) say which one and where the others are
    int park_cur = 0, park_slot1 = 0, park_slot2 = 1, park_slot0 = 0;
    const uint32_t t_park = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(TC_TSLOTS * TC_CHUNK_N + sgrp * 20);
#define CST(l) cst[0]
#else
    __half2 cst[3][TC_NCHUNK][2];
#define CST(l) cst[l]
#endif
    uint32_t tslot = 0, tuse = 0;
    const int ts0 = 1024 + (warp ? 2048 : 0);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)sgrp * 16;
    uint32_t v[16];           // accumulator chunk: 4 units x (i, j, f, o)
    bool have = false;        // v already holds the in-flight load of the next chunk (TC_PREFETCH == 2)
#if TC_INTERLEAVE
#define FROW(r) s_frow[(it & 1) * 128 + (r)]
    const int n_inst = 2 * n_iter, n_steps = 2 * TC_STEPS_PER_DIR * n_iter;
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int j = 0; j < TC_NCHUNK; ++j) CST(l)[j][0] = CST(l)[j][1] = __floats2half2_rn(0.f, 0.f);
    // one cell-step; the layer is a compile-time constant (the cell state of a layer lives in registers)
    auto step = [&](auto LC, const int inst, const int sidx, const int d, const int g) {
      constexpr int l = decltype(LC)::value;
      const int it = inst >> 1, dir = inst & 1, t = d - l;
      const int64_t win0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * DM_TILE_M;
      const bool stamp = false;
      float cls_acc = 0.f;
      {
        {
          // S[29] = (10, 2): stage what the NEXT direction instance's first steps read (x(0), x(1), the step-0 extras
          // next to zeroed units 96..99 in h0[1] column 12).  Their last readers (steps S[27], S[28]) have retired,
          // and this step's "done" arrival is the one the issuers wait for before that instance's first MMA.
          const bool stage = sidx == 29 && inst + 1 < n_inst && tid < 384;
          uint4 sv = make_uint4(0, 0, 0, 0);
          int sfr = 0;
          if (stage) {
            const int ndir = (inst + 1) & 1, nit = (inst + 1) >> 1, r = tid & 127;
            sfr = ndir == 0 ? win_frow[((int64_t)blockIdx.x + (int64_t)nit * gridDim.x) * DM_TILE_M + r] : FROW(r);
            if (tid < 256) {
              sv = *reinterpret_cast<const uint4*>(feat_tc + ((int64_t)sfr + tau_row(ndir, tid >> 7)) * 16);
            } else {
              const uint2 e = *reinterpret_cast<const uint2*>(feat_tc + ((int64_t)sfr + tau_row(ndir, 0)) * 16 + 8);
              sv.z = e.x; sv.w = e.y;
            }
          }
#else
#define FROW(r) s_frow[(r)]
    for (int it = 0; it < n_iter; ++it) {
    const int64_t win0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * DM_TILE_M;
    const int G0 = it * 2 * TC_STEPS_PER_DIR;
    const bool stamp = it == 0 && lane == 0 && (warp == 0 || warp == TC_EPI_WARPS - 1);
    int g = 0;
    for (int dir = 0; dir < 2; ++dir) {
#pragma unroll
      for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int j = 0; j < TC_NCHUNK; ++j) CST(l)[j][0] = CST(l)[j][1] = __floats2half2_rn(0.f, 0.f);
      float cls_acc = 0.f;
      for (int d = 0; d < 13; ++d) {
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          const int t = d - l;
          if (t < 0 || t > 10) continue;
          if (g >= max_steps) { ++g; continue; }
#endif
          if (TC_STATIC_SLOTS) tslot = 0;           // (it is 0 already: said so for constant folding in the unrolled chunks)
#if TC_PARK
          if (park_cur != l) {
            // bring layer l's cell state into the registers, park the one that was there in the slot l leaves
            const int ps = l == 0 ? park_slot0 : l == 1 ? park_slot1 : park_slot2;
            const uint32_t pa = t_park + (uint32_t)ps * 10;
            uint32_t* cr = reinterpret_cast<uint32_t*>(&cst[0][0][0]);
            uint32_t nw[10];
            if (t != 0) {                           // (at t = 0 the state of layer l is not read: c(-1) = 0)
              tc_ld8(pa, nw);
              tc_ld2(pa + 8, nw + 8);
              tc_wait_ld();
            }
            tc_st8(pa, cr);
            tc_st2(pa + 8, cr + 8);
            tc_wait_st();
            if (t != 0) {
#pragma unroll
              for (int i = 0; i < 10; ++i) cr[i] = nw[i];
            }
            if (park_cur == 0) park_slot0 = ps; else if (park_cur == 1) park_slot1 = ps; else park_slot2 = ps;
            park_cur = l;
          }
#endif
          // prefetch what this step's epilogue must stage for later steps
          uint4 xnext = make_uint4(0, 0, 0, 0);
          uint2 lows = make_uint2(0, 0);
          if (l == 0) {
            if (sgrp == 0 && t + 2 <= 10)
              xnext = *reinterpret_cast<const uint4*>(feat_tc + ((int64_t)FROW(row) + tau_row(dir, t + 2)) * 16);
            if (sgrp == 4)   // (1, 1, mean_lo, stdv_lo) of the next time index ride in h0's extras
              lows = *reinterpret_cast<const uint2*>(feat_tc + ((int64_t)FROW(row) + tau_row(dir, t + 1 <= 10 ? t + 1 : 10)) * 16 + 8);
          } else if (l == 1) {
            lows = make_uint2(F16 ? 0x3C003C00u : 0x3F803F80u, 0u);     // (1, 1, 0, 0): bias carriers for layer 2
          }
          const uint32_t htile = l == 0 ? OFF_H0 + (t & 1) * TC_HTILE : l == 1 ? OFF_H1 + (t & 1) * TC_HTILE : OFF_H2;
          // h2 is single-buffered and its old value is an operand of ALL five chunks of this very
          // step: layer 2 keeps its new h in registers until the last chunk's MMAs have retired
          uint32_t hkeep[TC_NCHUNK][2];
          // ... which they normally have by the time this step starts (the issuers run ahead): one probe of the last
          // chunk's barrier decides whether the stores can go out chunk by chunk like in the other layers
          bool direct = true;
          if (l == 2) {
            // (chunks 3 and 4 are the last ones of the two issuer threads, whose MMAs retire in order per thread)
            uint32_t s3 = tslot + (TC_NCHUNK - 2), u3 = tuse, s4 = tslot + (TC_NCHUNK - 1), u4 = tuse;
            if (s3 >= TC_TSLOTS) { s3 -= TC_TSLOTS; ++u3; }
            if (s4 >= TC_TSLOTS) { s4 -= TC_TSLOTS; ++u4; }
            const bool done = mbar_test(bar0 + 8 * (BAR_TFULL + s3), u3 & 1) && mbar_test(bar0 + 8 * (BAR_TFULL + s4), u4 & 1);
            direct = __shfl_sync(0xffffffffu, (int)done, 0) != 0;
          }
          // may the first chunk of the NEXT cell-step be loaded before this step is reported done?  Only if its
          // MMAs cannot wait for this step's hidden state (the wavefront corners and direction ends do)
#if TC_INTERLEAVE
          const bool cross = TC_PREFETCH == 2 && g != 0 && g < n_steps - 2;
#else
          const bool cross = TC_PREFETCH == 2 && !(d == 0 || (d == 11 && l == 2) || d == 12) && g + 1 < max_steps;
#endif
          if (TC_PREFETCH && !have) {
            mbar_wait(bar0 + 8 * (BAR_TFULL + tslot), tuse & 1);
            tc_fence_after();
            tc_ld16(t_lane + tslot * TC_CHUNK_N, v);
          }
          // 2h of chunk jj (units tc_unit0(jj, sgrp) ..+3) as two packed 16-bit pairs: store into the next A operand.
          // With TC_UNITMAP two consecutive chunks fill one 16-byte row segment (8 units): the first half waits in two
          // registers and the pair leaves as ONE 16-byte store per lane (32 lanes x 16 B contiguous: no bank conflict).
          // Even column groups pair chunks (0,1), (2,3); odd ones start mid-segment and pair (1,2), (3,4).
          uint32_t ph0 = 0, ph1 = 0;
          auto emit = [&](int jj, const uint32_t h01, const uint32_t h23) {
            const int u0 = tc_unit0(jj, sgrp);
            // core column u0/8, byte (u0%8)*2 of the row's 16 B
            unsigned char* dst = smem + htile + (u0 >> 3) * TC_ACOL + row_off + (u0 & 7) * 2;
            const bool odd = (sgrp & 1) != 0;
            const bool pair_lo = TC_UNITMAP && (odd ? (jj == 1 || jj == 3) : (jj == 0 || jj == 2));
            const bool pair_hi = TC_UNITMAP && (odd ? (jj == 2 || jj == 4) : (jj == 1 || jj == 3));
            if (TC_INTERLEAVE && l == 2 && t == 10) {
              // h2(10) only feeds the classifier; storing it would race with the next instance's zeroing of h2 column 0
            } else if (l == 2 && !direct) {
              hkeep[jj][0] = h01; hkeep[jj][1] = h23;
              if (jj == TC_NCHUNK - 1) {
#pragma unroll
                for (int j2 = 0; j2 < TC_NCHUNK; ++j2) {
                  const int uu = tc_unit0(j2, sgrp);
                  *reinterpret_cast<uint2*>(smem + htile + (uu >> 3) * TC_ACOL + row_off + (uu & 7) * 2) =
                      make_uint2(hkeep[j2][0], hkeep[j2][1]);
                }
              }
            } else if (jj == TC_NCHUNK - 1 && sgrp == 4) {
              *reinterpret_cast<uint4*>(dst) = make_uint4(h01, h23, lows.x, lows.y);      // units 96..99 + the extras
            } else if (pair_lo) {
              ph0 = h01; ph1 = h23;
            } else if (pair_hi) {
              *reinterpret_cast<uint4*>(dst - 8) = make_uint4(ph0, ph1, h01, h23);
            } else {
              *reinterpret_cast<uint2*>(dst) = make_uint2(h01, h23);
            }
          };
          const __half2 half2_half = __floats2half2_rn(0.5f, 0.5f);
#pragma unroll
          for (int j = 0; j < TC_NCHUNK; ++j) {
            if (stamp) TS(ts0 + g * 16 + 3 * j);
            if (!TC_PREFETCH) {
              mbar_wait(bar0 + 8 * (BAR_TFULL + tslot), tuse & 1);
              tc_fence_after();
              if (stamp) TS(ts0 + g * 16 + 3 * j + 1);
              tc_ld16(t_lane + tslot * TC_CHUNK_N, v);
            }
            tc_wait_ld16(v);       // .sync.aligned: the whole warp's loads have landed here, no __syncwarp needed
            tc_fence_before();
            if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_TEMPTY + tslot));      // CTA-local; the peer's relay forwards
            if (++tslot == TC_TSLOTS) { tslot = 0; ++tuse; }
            if (stamp) TS(ts0 + g * 16 + 3 * j + 2);
            // gate pre-activations of units (0,1) and (2,3) as f16x2 (low half = the even unit): one cvt per pair
            // and gate, and from here on v is dead.  (TC_GATE32: the tanh of the four gates is taken in f32 from the
            // accumulator registers and the RESULTS are packed, which saves the PRMT after every pair of MUFU.TANH.F16.)
            __half2 gi[2], gj[2], gf[2], go[2];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              if (TC_GATE32) {
                if (TC_FMAI) gi[p] = __floats2half2_rn(__uint_as_float(v[8 * p + 0]), __uint_as_float(v[8 * p + 4]));
                else gi[p] = __floats2half2_rn(tanh32_mufu(__uint_as_float(v[8 * p + 0])), tanh32_mufu(__uint_as_float(v[8 * p + 4])));
                gj[p] = __floats2half2_rn(tanh32_mufu(__uint_as_float(v[8 * p + 1])), tanh32_mufu(__uint_as_float(v[8 * p + 5])));
                if (TC_FMAF) gf[p] = __floats2half2_rn(__uint_as_float(v[8 * p + 2]), __uint_as_float(v[8 * p + 6]));
                else if (!(TC_T0SKIP && t == 0))
                  gf[p] = __floats2half2_rn(tanh32_mufu(__uint_as_float(v[8 * p + 2])), tanh32_mufu(__uint_as_float(v[8 * p + 6])));
                go[p] = __floats2half2_rn(tanh32_mufu(__uint_as_float(v[8 * p + 3])), tanh32_mufu(__uint_as_float(v[8 * p + 7])));
              } else {
                gi[p] = __floats2half2_rn(__uint_as_float(v[8 * p + 0]), __uint_as_float(v[8 * p + 4]));
                gj[p] = __floats2half2_rn(__uint_as_float(v[8 * p + 1]), __uint_as_float(v[8 * p + 5]));
                gf[p] = __floats2half2_rn(__uint_as_float(v[8 * p + 2]), __uint_as_float(v[8 * p + 6]));
                go[p] = __floats2half2_rn(__uint_as_float(v[8 * p + 3]), __uint_as_float(v[8 * p + 7]));
              }
            }
            // the next chunk's accumulator load overlaps this chunk's cell update ((tslot, tuse) now name the next chunk)
            const bool nxt = TC_PREFETCH && (j + 1 < TC_NCHUNK || cross);
            bool loaded = false;
            if (nxt && TC_EARLYLD) {
              bool rdy = true;
              if (TC_EARLYLD == 2) rdy = __all_sync(0xffffffffu, mbar_test(bar0 + 8 * (BAR_TFULL + tslot), tuse & 1));
              else mbar_wait(bar0 + 8 * (BAR_TFULL + tslot), tuse & 1);
              if (rdy) {
                tc_fence_after();
                tc_ld16(t_lane + tslot * TC_CHUNK_N, v);
                loaded = true;
              }
            }
            // cell update of chunk j:  c' = c * sig(f) + sig(i) * tanh(j),  2h = tanh(c') * tanh(o/2) + tanh(c')
            uint32_t hp[2] = {0u, 0u};
            if (!dbg_nomath) {
#pragma unroll
              for (int p = 0; p < 2; ++p) {
                const __half2 tj = TC_GATE32 ? gj[p] : tanh2_mufu(gj[p]), to = TC_GATE32 ? go[p] : tanh2_mufu(go[p]);
                const __half2 si = TC_FMAI ? sigmoid2_fma(gi[p])
                                           : __hfma2(TC_GATE32 ? gi[p] : tanh2_mufu(gi[p]), half2_half, half2_half);
                const __half2 y = __hmul2(si, tj);
                __half2 cn;
                if (TC_T0SKIP && t == 0) {            // c_prev == 0: the forget gate cannot matter
                  cn = y;
                } else {
                  const __half2 cp = CST(l)[j][p];
                  if (TC_FMAF) cn = __hfma2(cp, sigmoid2_fma(gf[p]), y);
                  else cn = __hfma2(__hfma2(cp, TC_GATE32 ? gf[p] : tanh2_mufu(gf[p]), cp), half2_half, y);
                }
                CST(l)[j][p] = cn;
                const __half2 tc = tanh2_mufu(cn);
                const __half2 h2 = __hfma2(tc, to, tc);
                if (l == 2 && t == 10) {
                  const float2 hf = __half22float2(h2);
                  const float* cw = s_cls + dir * DM_HIDDEN + tc_unit0(j, sgrp) + 2 * p;
                  cls_acc += hf.x * cw[0] + hf.y * cw[1];
                }
                if (F16) {
                  hp[p] = *reinterpret_cast<const uint32_t*>(&h2);
                } else {
                  const float2 hf = __half22float2(h2);
                  hp[p] = pack_bf16(hf.x, hf.y);
                }
              }
            }
            if (nxt && !loaded) {
              mbar_wait(bar0 + 8 * (BAR_TFULL + tslot), tuse & 1);
              tc_fence_after();
              tc_ld16(t_lane + tslot * TC_CHUNK_N, v);
            }
            emit(j, hp[0], hp[1]);
          }
          if (l == 0 && sgrp == 0 && t + 2 <= 10)
            *reinterpret_cast<uint4*>(smem + OFF_X + (t & 1) * TC_ACOL + row_off) = xnext;
          if (l == 2 && t == 10) s_part[sgrp * 128 + row] += cls_acc;
#if TC_INTERLEAVE
          if (stage) {
            const int r = tid & 127;
            if (tid < 256) {
              *reinterpret_cast<uint4*>(smem + OFF_X + (tid >> 7) * TC_ACOL + r * 16) = sv;
              if (tid < 128 && ((inst + 1) & 1) == 0) s_frow[(((inst + 1) >> 1) & 1) * 128 + r] = sfr;
            } else {
              *reinterpret_cast<uint4*>(smem + OFF_H0 + TC_HTILE + 12 * TC_ACOL + r * 16) = sv;
            }
          }
          // the column that shares a K = 16 MMA with the last column of the tile below must read zero at t = 0:
          // h1[1] column 0 before S[2] = (1,1), h2 column 0 before S[5] = (2,2); their last readers were the
          // previous instance's S[31] / S[32], which sit just before S[1] / S[2] of this one in the order
          if (sidx == 1 && tid < 128) *reinterpret_cast<uint4*>(smem + OFF_H1 + TC_HTILE + tid * 16) = make_uint4(0, 0, 0, 0);
          if (sidx == 2 && tid < 128) *reinterpret_cast<uint4*>(smem + OFF_H2 + tid * 16) = make_uint4(0, 0, 0, 0);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_HDONE + (g & 1)));         // CTA-local; the peer's relay forwards
          have = cross;
          if (sidx == 32 && dir == 1) {
            // both directions' classifier partials of this tile are in: softmax over two classes
            epi_bar();
            if (sgrp == 0) {
              float dl = w.cls_db;
#pragma unroll
              for (int s5 = 0; s5 < 5; ++s5) { dl += 0.5f * s_part[s5 * 128 + row]; s_part[s5 * 128 + row] = 0.f; }   // partials hold 2h
              if (p1_out) p1_out[win0 + row] = 1.0f / (1.0f + __expf(-dl));
              if (pred_out) pred_out[win0 + row] = dl > 0.f ? 1 : 0;
            }
          }
        }
      }
    };
    for (int GG = 0; GG < n_steps; ++GG) {
      int inst, sidx, d, l;
      il_decode(GG, n_inst, inst, sidx);
      il_step(sidx, d, l);
      if (l == 0) step(std::integral_constant<int, 0>{}, inst, sidx, d, GG);
      else if (l == 1) step(std::integral_constant<int, 1>{}, inst, sidx, d, GG);
      else step(std::integral_constant<int, 2>{}, inst, sidx, d, GG);
    }
#else
          if (d == 12) {
            // last cell-step of a direction: everything the next direction / next tile needs is staged
            // BEFORE this step is reported done, because the issuers' next step waits on exactly that
            epi_bar();          // every warp is done with this direction
            if (dir == 0) {
              if (max_steps > TC_STEPS_PER_DIR) dir_init(1, false);
            } else {
              if (sgrp == 0) {
                float dl = w.cls_db;
#pragma unroll
                for (int s = 0; s < 5; ++s) dl += 0.5f * s_part[s * 128 + row];      // the partials hold 2h
                // softmax over two classes: p1 = 1/(1+exp(l0-l1)); argmax picks class 1 iff l1 > l0
                if (p1_out) p1_out[win0 + row] = 1.0f / (1.0f + __expf(-dl));
                if (pred_out) pred_out[win0 + row] = dl > 0.f ? 1 : 0;
              }
              if (it + 1 < n_iter) {
                epi_bar();        // partial logits consumed
                for (int i = tid; i < 5 * 128; i += TC_EPI_THREADS) s_part[i] = 0.f;
                if (tid < DM_TILE_M) s_frow[tid] = win_frow[win0 + (int64_t)gridDim.x * DM_TILE_M + tid];
                epi_bar();        // next tile's feature rows visible
                dir_init(0, false);
              }
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_HDONE + (g & 1)));         // CTA-local; the peer's relay forwards
          if (stamp) TS(ts0 + g * 16 + 15);
          have = cross;
          ++g;
        }
      }
    }
    (void)G0;
    }
#endif
#undef FROW
  }
#undef FOR_EACH_STEP
#undef TS
  tc_fence_before();
  __syncthreads();
  if (DBG && dbg != nullptr && blockIdx.x == 0)
    for (int i = tid; i < OFF_W / 16; i += TC_THREADS)
      reinterpret_cast<uint4*>(dbg)[i] = *reinterpret_cast<const uint4*>(smem + i * 16);
  if (PAIR) cluster_sync_all();       // the peer may still be reading this CTA's operands / signalling its barriers
  if (warp == W_ALLOC) { tc_fence_after(); if (PAIR) tc_dealloc2(tmem_base); else tc_dealloc(tmem_base); }
}

// ---- descriptor / TMEM / bulk-copy self-test: D[128,n] = A[128,k] * B[n,k]^T -----------------
// A's core columns are placed at irregular (ascending) offsets on purpose, the way the BiLSTM
// kernel strings hidden tiles together, so the per-MMA leading-byte-offset is exercised.
__global__ void __launch_bounds__(128, 1)
k_umma_selftest(const unsigned char* __restrict__ a_img, const unsigned char* __restrict__ b_img, int n, int kcols,
                float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A column kc lives at kc*2048, plus a 4 KB hole after every 13th column
  auto acol = [](int kc) { return (uint32_t)(kc * TC_ACOL + (kc / 13) * 4096); };
  const uint32_t b_off = acol(kcols) + 4096;
  for (int kc = 0; kc < kcols; ++kc)
    *reinterpret_cast<uint4*>(smem + acol(kc) + tid * 16) = *reinterpret_cast<const uint4*>(a_img + (size_t)kc * TC_ACOL + tid * 16);
  fence_async_smem();
  const uint32_t bfull = smem_u32(&bars[0]), bdone = smem_u32(&bars[1]);
  if (tid == 0) {
    mbar_init(bfull, 1);
    mbar_init(bdone, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tc_alloc(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t bcol = (uint32_t)n * 16;
  if (tid == 0) {
    mbar_expect_tx(bfull, bcol * kcols);
    bulk_g2s(sbase + b_off, b_img, bcol * kcols, bfull);
    mbar_wait(bfull, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc(128, n);
    for (int k16 = 0; k16 < kcols / 2; ++k16) {
      const uint32_t a0 = sbase + acol(2 * k16), a1 = sbase + acol(2 * k16 + 1);
      tc_mma(tmem_base, umma_desc(a0, a1 - a0, 128), umma_desc(sbase + b_off + 2 * k16 * bcol, bcol, 128), idesc,
             k16 > 0 ? 1u : 0u);
    }
    tc_commit(bdone);
  }
  mbar_wait(bdone, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t v[16];
    tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc_wait_ld();
    for (int i = 0; i < 16; ++i) out[(size_t)(warp * 32 + lane) * n + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tc_dealloc(tmem_base); }
}

uint16_t h_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
float h_bf16f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// 16-bit operand formats of the tensor-core path (host side): round to nearest even, and back
uint16_t h_rn16(float f, bool f16) { return f16 ? __half_as_ushort(__float2half_rn(f)) : h_bf16(f); }
float h_rn16f(uint16_t b, bool f16) { return f16 ? __half2float(__ushort_as_half(b)) : h_bf16f(b); }

}  // namespace

// Host-side image of one (direction, layer) for the weight ring: [chunk j][core column kc][n 80][8 k]
// bf16, n = 4*unit_in_chunk + gate.  K order (virtual operand columns):
//   layer 0: col 0 = (A, C, G, T, mean_hi, stdv_hi, len_hi, len_lo); cols 1..13 = h0 tile
//            (k 0..99 hidden, 100/101 bias hi/lo against the ones, 102 mean_lo, 103 stdv_lo)
//   layer 1,2: cols 0..12 = tile of the layer below (100/101 carry this layer's bias),
//            cols 13..25 = own hidden tile (extras unused).
// The sigmoid gates (i, f, o) are pre-scaled by 0.5 because the epilogue evaluates
// sigmoid(x) = 0.5*tanh(x/2)+0.5; forget_bias = 1.0 (BasicLSTMCell) is folded into the bias.  Every row that
// multiplies a hidden state carries another 0.5: the epilogue stores 2h = tanh(c)*tanh(o/2) + tanh(c).
// (Both factors are powers of two: the rounded weights are exactly the scaled roundings.)
// pair = true: [chunk j][cta r][core column kc][n 40][8 k], CTA r owning gate columns n = 40 r .. 40 r + 39
// of the chunk (tcgen05.mma.cta_group::2 takes the first half of B's N rows from the even CTA).
// f16: fp16 instead of bf16 words.
void dm_tc_pack_weights(const float* kernel, const float* bias, int layer, bool pair, bool f16, std::vector<uint16_t>& img) {
  const int ncols = layer == 0 ? 14 : 26;
  img.assign((size_t)TC_NCHUNK * ncols * TC_CHUNK_N * 8, 0);
  for (int j = 0; j < TC_NCHUNK; ++j)
    for (int kc = 0; kc < ncols; ++kc)
      for (int n = 0; n < TC_CHUNK_N; ++n) {
        const int unit = tc_unit0(j, n / 16) + (n / 4) % 4, gate = n % 4, col = gate * DM_HIDDEN + unit;
        // (TC_FMAF: the forget gate's pre-activation is wanted as x * log2(e), not x / 2)
        const float scale = gate == 1 ? 1.0f : ((gate == 2 && TC_FMAF) || (gate == 0 && TC_FMAI)) ? 1.4426950408889634f : 0.5f;
        const float bsc = (bias[col] + (gate == 2 ? 1.0f : 0.0f)) * scale;
        const uint16_t bhi = h_rn16(bsc, f16), blo = h_rn16(bsc - h_rn16f(bhi, f16), f16);
        for (int e = 0; e < 8; ++e) {
          uint16_t val = 0;
          auto wref = [&](int r) { return h_rn16(kernel[(size_t)r * DM_GATES + col] * scale, f16); };          // input rows
          auto href = [&](int r) { return h_rn16(kernel[(size_t)r * DM_GATES + col] * scale * 0.5f, f16); };   // rows against 2h
          if (layer == 0) {
            if (kc == 0) {
              static const int xr[8] = {0, 1, 2, 3, 4, 5, 6, 6};
              val = wref(xr[e]);
            } else {
              const int kk = (kc - 1) * 8 + e;
              if (kk < DM_HIDDEN) val = href(DM_FNUM + kk);
              else if (kk == 100) val = bhi;
              else if (kk == 101) val = blo;
              else if (kk == 102) val = wref(4);
              else val = wref(5);
            }
          } else {
            if (kc < TC_HCOLS) {
              const int kk = kc * 8 + e;
              if (kk < DM_HIDDEN) val = href(kk);
              else if (kk == 100) val = bhi;
              else if (kk == 101) val = blo;
            } else {
              const int kk = (kc - TC_HCOLS) * 8 + e;
              if (kk < DM_HIDDEN) val = href(DM_HIDDEN + kk);
            }
          }
          if (pair) img[((((size_t)j * 2 + n / 40) * ncols + kc) * 40 + n % 40) * 8 + e] = val;
          else img[(((size_t)j * ncols + kc) * TC_CHUNK_N + n) * 8 + e] = val;
        }
      }
}

using tc_kernel_t = void (*)(const uint16_t*, const int32_t*, dm_dev_weights, float*, uint8_t*, int, int, unsigned char*);
static tc_kernel_t tc_kernel(bool pair, bool dbg, bool f16) {
  static const tc_kernel_t tab[8] = {
      k_lstm_tc<false, false, false>, k_lstm_tc<false, false, true>, k_lstm_tc<false, true, false>, k_lstm_tc<false, true, true>,
      k_lstm_tc<true, false, false>,  k_lstm_tc<true, false, true>,  k_lstm_tc<true, true, false>,  k_lstm_tc<true, true, true>};
  return tab[(pair ? 4 : 0) + (dbg ? 2 : 0) + (f16 ? 1 : 0)];
}

static int tc_prepare(dm_ctx* ctx) {
  if (!ctx->tc_attr_set) {
    for (int i = 0; i < 8; ++i)
      DM_CUDA(ctx, cudaFuncSetAttribute(reinterpret_cast<const void*>(tc_kernel((i & 4) != 0, (i & 2) != 0, (i & 1) != 0)),
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    DM_CUDA(ctx, cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ctx->tc_attr_set = true;
  }
  return DM_OK;
}

static int tc_launch(dm_ctx* ctx, const uint16_t* feat_tc, const int32_t* win_frow, int64_t n_pad, float* p1,
                     uint8_t* pred, int max_steps, unsigned char* dbg) {
  const unsigned tiles = (unsigned)(n_pad / DM_TILE_M);       // n_pad is a multiple of 256: an even number of tiles
  // persistent CTAs: one per SM (an even number, so that pairs stay whole); the debug entry runs one tile per CTA
  unsigned grid = (unsigned)(ctx->sm_count & ~1);
  if (dbg != nullptr || max_steps != 2 * TC_STEPS_PER_DIR || grid > tiles) grid = tiles;
  const bool debug = dbg != nullptr || max_steps != 2 * TC_STEPS_PER_DIR;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ctx->tc_pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DM_CUDA(ctx, cudaLaunchKernelEx(&cfg, tc_kernel(ctx->tc_pair, debug, ctx->tc_f16), feat_tc, win_frow, ctx->w, p1, pred,
                                  (int)tiles, max_steps, dbg));
  ctx->launches += 1;
  DM_CUDA(ctx, cudaGetLastError());
  return DM_OK;
}

int dm_launch_lstm_tc(dm_ctx* ctx, const uint16_t* feat_tc, const int32_t* win_frow, int64_t n_windows,
                      float* p1, uint8_t* pred) {
  const int64_t n_pad = dm_pad_windows(n_windows);
  if (n_pad == 0) return DM_OK;
  int rc = tc_prepare(ctx);
  if (rc != DM_OK) return rc;
  return tc_launch(ctx, feat_tc, win_frow, n_pad, p1, pred, 2 * TC_STEPS_PER_DIR, nullptr);
}

// debug: run the first `max_steps` cell-steps and return tile 0's shared-memory operand region
// (x columns + the five hidden tiles, OFF_W bytes)
int dm_tc_debug(dm_ctx* ctx, const uint16_t* feat_tc, const int32_t* win_frow, int64_t n_windows, float* p1,
                uint8_t* pred, int max_steps, unsigned char* dump_host, int64_t dump_cap) {
  const int64_t n_pad = dm_pad_windows(n_windows);
  if (n_pad == 0) return DM_OK;
  int rc = tc_prepare(ctx);
  if (rc != DM_OK) return rc;
  unsigned char* dbg = nullptr;
  DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&dbg), OFF_W + TS_BYTES));
  cudaMemsetAsync(dbg, 0, OFF_W + TS_BYTES, ctx->stream);
  cudaEventRecord(ctx->ev1, ctx->stream);
  rc = tc_launch(ctx, feat_tc, win_frow, n_pad, p1, pred, max_steps, dbg);
  cudaEventRecord(ctx->ev2, ctx->stream);
  cudaError_t e = rc == DM_OK ? cudaStreamSynchronize(ctx->stream) : cudaErrorUnknown;
  if (e == cudaSuccess) cudaEventElapsedTime(&ctx->lstm_ms, ctx->ev1, ctx->ev2);
  if (e == cudaSuccess && dump_host)
    e = cudaMemcpy(dump_host, dbg, (size_t)std::min<int64_t>(dump_cap, OFF_W + TS_BYTES), cudaMemcpyDeviceToHost);
  cudaFree(dbg);
  if (rc != DM_OK) return rc;
  if (e != cudaSuccess) { dm_set_error(ctx, std::string("dm_tc_debug: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  return DM_OK;
}

int dm_tc_selftest(dm_ctx* ctx, int n, int k, float* max_err) {
  if (n < 16 || n > 256 || n % 16 || k < 16 || k > 416 || k % 16) {
    dm_set_error(ctx, "dm_selftest_umma: need n in [16,256] and k in [16,416], both multiples of 16");
    return DM_ERR_ARG;
  }
  int rc = tc_prepare(ctx);
  if (rc != DM_OK) return rc;
  const int kcols = k / 8;
  std::vector<float> A((size_t)128 * k), B((size_t)n * k);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
  for (auto& x : A) x = h_bf16f(h_bf16(rnd()));
  for (auto& x : B) x = h_bf16f(h_bf16(rnd() * 4.0f));
  std::vector<uint16_t> ai((size_t)kcols * 128 * 8), bi((size_t)kcols * n * 8);
  for (int kc = 0; kc < kcols; ++kc)
    for (int e = 0; e < 8; ++e) {
      for (int r = 0; r < 128; ++r) ai[((size_t)kc * 128 + r) * 8 + e] = h_bf16(A[(size_t)r * k + kc * 8 + e]);
      for (int c = 0; c < n; ++c) bi[((size_t)kc * n + c) * 8 + e] = h_bf16(B[(size_t)c * k + kc * 8 + e]);
    }
  unsigned char *a_d = nullptr, *b_d = nullptr;
  float* o_d = nullptr;
  DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&a_d), ai.size() * 2));
  DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b_d), bi.size() * 2));
  DM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&o_d), sizeof(float) * 128 * n));
  DM_CUDA(ctx, cudaMemcpy(a_d, ai.data(), ai.size() * 2, cudaMemcpyHostToDevice));
  DM_CUDA(ctx, cudaMemcpy(b_d, bi.data(), bi.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)kcols * TC_ACOL + (kcols / 13 + 2) * 4096 + (size_t)kcols * n * 16;
  if (smem > 200 * 1024) { dm_set_error(ctx, "dm_selftest_umma: n*k too large for one CTA"); return DM_ERR_ARG; }
  k_umma_selftest<<<1, 128, smem, ctx->stream>>>(a_d, b_d, n, kcols, o_d);
  ctx->launches += 1;
  std::vector<float> out((size_t)128 * n);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy(out.data(), o_d, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(a_d); cudaFree(b_d); cudaFree(o_d);
  if (e != cudaSuccess) { dm_set_error(ctx, std::string("dm_selftest_umma: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  double worst = 0.0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < n; ++c) {
      double ref = 0.0;
      for (int kk = 0; kk < k; ++kk) ref += (double)A[(size_t)r * k + kk] * (double)B[(size_t)c * k + kk];
      worst = std::max(worst, std::fabs(ref - (double)out[(size_t)r * n + c]));
    }
  *max_err = (float)worst;
  return DM_OK;
}
