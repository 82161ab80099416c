// tcgen05 MLP engines (mlp_engine = 2: fp16x3 split operands, the default; mlp_engine = 0: 3xTF32): the NeuMesh
// geometry / colour MLPs on the 5th-generation tensor cores.
//
// Why split operands.  The sdf feeds sigmoid(s * sdf) with s ~ 50-300 and a discrete re-sampling cascade; single-pass TF32
// (10-bit mantissa) or BF16 operands miss the 1e-4 / 1e-5 parity bar by 1-3 orders of magnitude, while the split
// x = hi + lo (both TF32), D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi with fp32 accumulation in TMEM is fp32-accurate
// (SURVEY.md section 7.3); the fp16 variant (hi = fp16(x), lo = fp16(x - hi), weights pre-scaled by 2^8) is as accurate
// at twice the MMA rate and half the operand bytes.  Every algorithmic MAC is therefore issued three times: tensor-pipe
// utilisation is quoted against ISSUED MMAs and the 3x factor is stated wherever a FLOP/s figure appears.
//
// One persistent CTA per SM (clusters of 2 share every weight slab), 128 rows (points) per tile, warp-specialised:
//   warps 0-15  epilogue : TMEM -> registers (tcgen05.ld), bias + activation, hi/lo split, next layer's A slabs -> smem;
//                          quadrant = w & 3 (TMEM lanes), chunk parity = (w >> 2) & 1, column half = w >> 3
//   warps 16-23 builder  : neighbour gather + blend + positional encoding -> first-layer A slabs (2 threads per row)
//   warp  24    MMA      : one lane issues tcgen05.mma (M = 128, N = 256; kind::f16 K = 16 or kind::tf32 K = 8) and the
//                          tcgen05.commit that release ring slots / publish the accumulator
//   warp  25    loader   : weight slabs L2 -> smem with cp.async.bulk (TMA 1-D bulk copy, cluster multicast) + complete_tx
// Layer l's 128x256 fp32 accumulator lives in TMEM columns [256*(l&1), +256); while the epilogue drains it 16 columns
// at a time into K-slabs of layer l+1, the MMA warp is already accumulating layer l+1 into the other half.
// Schedules that were built and measured slower on B200 (CTA pairs with cta_group::2, two tiles per CTA, shuffle exchange
// for the tangent rows, packed f32x2 epilogue arithmetic): profiles/r2_mlp_schedule_experiments.txt, DESIGN.md 4.2.
//
// Operand layout (no-swizzle, K-major "interleave" canonical layout): a K-slab of 16 columns is stored as
// [k/4][row][k%4] fp32, i.e. 8x16-byte core matrices with SBO = 128 B (next 8 rows) and LBO = rows*16 B (next 4 k).
// Weights are pre-packed in exactly this image (hi slab then lo slab) so a slab is ONE contiguous 32 KB bulk copy.
#include <cuda_fp16.h>

#include <cstdlib>
#include <vector>

#include "field_build.cuh"

namespace nmb {

namespace tc {

constexpr int ROWS = 128;                  // tile rows (TMEM lanes)
constexpr int SLAB_K = 16;                 // K columns per pipeline slab
constexpr int A_HALF = ROWS * SLAB_K * 4;  // 8 KB: one hi (or lo) A slab
constexpr int A_SLOT = 2 * A_HALF;         // 16 KB
constexpr int B_HALF = MLP_W * SLAB_K * 4; // 16 KB
constexpr int B_SLOT = 2 * B_HALF;         // 32 KB
constexpr int NA0 = 2;                     // first-layer A ring (builder -> MMA)
constexpr int NA1 = 4;                     // hidden-layer A ring (epilogue -> MMA)
constexpr int NA = NA0 + NA1;
constexpr int NB = 3;                      // B ring slots (loader -> MMA)
constexpr int CLUSTER = 2;                 // CTAs per cluster sharing every weight slab through TMA multicast
// NOTE two separate A rings: an mbarrier parity wait is only meaningful while the waiter is at most one phase
// ahead of the barrier.  The builder runs a whole tile ahead of the epilogue, so the two producer groups must not
// share one ring (a shared ring deadlocks as soon as a CTA processes a second tile).
constexpr int N_EPI = 512;                 // 16 warps: quadrant = w & 3 (TMEM lanes), group g = (w >> 2) & 1 drains the chunks
                                           // j with j % 2 == g, half hh = w >> 3 takes columns [8 hh, 8 hh + 8) of a chunk
constexpr int N_BUILD = 256;               // 2 threads per row: half h builds columns [8h, 8h+8) of every first-layer slab
constexpr int THREADS = N_EPI + N_BUILD + 64;
constexpr int WARP_BUILD = N_EPI / 32, WARP_MMA = (N_EPI + N_BUILD) / 32;
constexpr int SIG_BUF = 64 * 16;           // floats per exp(100 z) exchange buffer (64 value rows x 16 columns)
constexpr int CONST_FLOATS = (MAX_LAYERS + 3) * MLP_W;   // biases of every hidden layer + up to 3 output rows

// fp16x3 operand variant (mlp_engine = 2): the same slabs with fp16 hi / lo images - half the bytes, and a 16-column
// slab is ONE kind::f16 K step (K = 16) instead of two tf32 ones.  The halved slots buy rings twice as deep.
constexpr int A_HALF16 = ROWS * SLAB_K * 2;   // 4 KB
constexpr int A_SLOT16 = 2 * A_HALF16;        // 8 KB
constexpr int B_HALF16 = MLP_W * SLAB_K * 2;  // 8 KB
constexpr int B_SLOT16 = 2 * B_HALF16;        // 16 KB
constexpr float F16_W_SCALE = 256.f;          // weights are packed as 2^8 W (keeps their lo parts out of the subnormals)
// Synchronisation step of the fp16 engine: GR16 = 2 slabs (32 K-columns) share ONE ring slot, i.e. one full / empty
// barrier pair and one tcgen05.commit pair.  Measured in round 2 (profiles/r2_mlp_schedule_experiments.txt): the single
// MMA-issuing thread pays ~560 clk of fixed cost per ring step (two barrier polls, tcgen05.fence, descriptors, two
// commits - one of them cluster-multicast) against ~50 clk per MMA it issues, and that fixed cost - not the tensor core -
// bounded the engine; one step per 32 columns halves it.  Ring depths below are in steps.
constexpr int GR16 = 2;
constexpr int NA0_16 = 2, NA1_16 = 4, NA_16 = NA0_16 + NA1_16, NB_16 = 3;

// SIG: the kernel instantiation exchanges exp(100 z) between value and tangent rows (MODE 1); the others spend those
// 16 KB on one more first-layer ring step (the MMA thread waits for the builder ~10 % of its time with two)
template <bool F16, bool SIG = true>
struct SmemLayoutT {
  static constexpr int NA0_ = F16 ? (SIG ? NA0_16 : NA0_16 + 1) : NA0;
  static constexpr int a_off = 0;
  static constexpr int b_off = F16 ? (NA0_ + NA1_16) * GR16 * A_SLOT16 : NA * A_SLOT;
  static constexpr int sig_off = b_off + (F16 ? NB_16 * GR16 * B_SLOT16 : NB * B_SLOT);   // [group][parity] buffers
  static constexpr int part_off = sig_off + ((SIG || !F16) ? 4 * SIG_BUF * 4 : 0);   // last-layer partial sums: [3 helpers][3][128]
  static constexpr int const_off = part_off + 9 * ROWS * 4;    // biases + output weights
  static constexpr int bar_off = const_off + CONST_FLOATS * 4;
  static constexpr int total = bar_off + (F16 ? 512 : 256);   // 2 (NA + NB) + 4 mbarriers + the TMEM base address
};
using SmemLayout = SmemLayoutT<false>;
static_assert(SmemLayoutT<true, true>::total <= 232448 && SmemLayoutT<true, false>::total <= 232448,
              "shared memory budget (fp16 variant)");
static_assert((2 * (NA_16 + 1 + NB_16) + 4) * 8 + 4 <= 512 && (2 * (NA + NB) + 4) * 8 + 4 <= 256, "mbarrier block");
static_assert(SmemLayout::total <= 232448, "shared memory budget");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// single-thread roles (MMA issuer, loader): back off between polls so the spin does not steal issue slots from the
// epilogue / builder warps that share the scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t done;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(40);
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Multicast variant: the bytes land at the same CTA-relative offset in every CTA of `mask`, and complete_tx is
// signalled on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64))
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
// kind::f16 (A and B fp16, fp32 accumulate), same shape: a_format = b_format = 0 (F16)
constexpr uint32_t IDESC_F16 = (1u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(IDESC_F16), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// the registers are tied to the wait ("+r") so that no consumer can be scheduled ahead of it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// write 16 consecutive K-columns of row `r` into an A slot: hi half then lo half, [k/4][row][k%4]
__device__ __forceinline__ void store_a_row(char* a_slot, int r, const float (&v)[16]) {
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    float4 hi, lo;
    hi.x = tf32_rna(v[kc * 4 + 0]);
    hi.y = tf32_rna(v[kc * 4 + 1]);
    hi.z = tf32_rna(v[kc * 4 + 2]);
    hi.w = tf32_rna(v[kc * 4 + 3]);
    lo.x = tf32_rna(v[kc * 4 + 0] - hi.x);
    lo.y = tf32_rna(v[kc * 4 + 1] - hi.y);
    lo.z = tf32_rna(v[kc * 4 + 2] - hi.z);
    lo.w = tf32_rna(v[kc * 4 + 3] - hi.w);
    *reinterpret_cast<float4*>(a_slot + kc * (ROWS * 16) + r * 16) = hi;
    *reinterpret_cast<float4*>(a_slot + A_HALF + kc * (ROWS * 16) + r * 16) = lo;
  }
}

// write 8 consecutive K-columns [8h, 8h+8) of row `r` (two 4-column chunks) into an A slot
__device__ __forceinline__ void store_a_half(char* a_slot, int r, int h, const float (&v)[8]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int kc = 2 * h + c;
    float4 hi, lo;
    hi.x = tf32_rna(v[c * 4 + 0]);
    hi.y = tf32_rna(v[c * 4 + 1]);
    hi.z = tf32_rna(v[c * 4 + 2]);
    hi.w = tf32_rna(v[c * 4 + 3]);
    lo.x = tf32_rna(v[c * 4 + 0] - hi.x);
    lo.y = tf32_rna(v[c * 4 + 1] - hi.y);
    lo.z = tf32_rna(v[c * 4 + 2] - hi.z);
    lo.w = tf32_rna(v[c * 4 + 3] - hi.w);
    *reinterpret_cast<float4*>(a_slot + kc * (ROWS * 16) + r * 16) = hi;
    *reinterpret_cast<float4*>(a_slot + A_HALF + kc * (ROWS * 16) + r * 16) = lo;
  }
}

// fp16 variant of store_a_half: the 8 columns [8h, 8h+8) of row `r` are ONE 16-byte chunk ([k/8][row][k%8] halves);
// hi = fp16(x), lo = fp16(x - hi) (lo may be subnormal: absolute error <= 2^-25, see tools/split_precision_study.py)
__device__ __forceinline__ void store_a_half_f16(char* a_slot, int r, int h, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const __half2 hp = __floats2half2_rn(v[2 * c], v[2 * c + 1]);   // .x (low 16 bits) = even column
    const float2 hf = __half22float2(hp);
    const __half2 lp = __floats2half2_rn(v[2 * c] - hf.x, v[2 * c + 1] - hf.y);
    hi[c] = *reinterpret_cast<const uint32_t*>(&hp);
    lo[c] = *reinterpret_cast<const uint32_t*>(&lp);
  }
  *reinterpret_cast<uint4*>(a_slot + h * (ROWS * 16) + r * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(a_slot + A_HALF16 + h * (ROWS * 16) + r * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct Params {
  FieldLayout lay;
  FieldIn in;
  FieldTables tab;
  const float* w;          // packed slabs
  const float* bias;       // [n_layers][256]
  const float* w_out;      // [n_out][256]
  const float* b_out;
  int64_t slab_off[MAX_LAYERS];  // floats
  int n_slabs[MAX_LAYERS];
  int n_layers;
  int slabs_per_tile;
  int64_t P;
  float* out0;
  float* out1;
  unsigned long long* dbg;   // optional [8] cycle counters (NMB_TC_PROFILE=1): where the pipeline waits
};

// wait that accounts its cycles into `acc` when profiling
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool prof, unsigned long long& acc) {
  if (!prof) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}
__device__ __forceinline__ void mbar_wait_bt(uint32_t bar, uint32_t parity, bool prof, unsigned long long& acc) {
  if (!prof) {
    mbar_wait_backoff(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait_backoff(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}

}  // namespace tc

// MODE 0: geometry, 128 points / tile.  MODE 1: geometry + tangent rows (rows 64..127 carry d/d(ds) of rows 0..63).
// MODE 2: colour, 128 points / tile.
// F16 = false: 3xTF32 operands (mlp_engine 0).  F16 = true: fp16x3 operands (mlp_engine 2, see A_HALF16 above).
template <int MODE, bool F16 = false>
__global__ void __cluster_dims__(tc::CLUSTER, 1, 1) __launch_bounds__(tc::THREADS, 1)
mlp_tc_kernel(const tc::Params prm) {
  using namespace tc;
  using SmemLayout = SmemLayoutT<F16, MODE == 1>;
  constexpr int GR = F16 ? GR16 : 1;   // 16-column slabs per ring slot (synchronisation step)
  constexpr int A_HALF = F16 ? A_HALF16 : tc::A_HALF, A_SUB = F16 ? A_SLOT16 : tc::A_SLOT, A_SLOT = GR * A_SUB;
  constexpr int B_HALF = F16 ? B_HALF16 : tc::B_HALF, B_SUB = F16 ? B_SLOT16 : tc::B_SLOT, B_SLOT = GR * B_SUB;
  constexpr int NA0 = SmemLayout::NA0_, NA1 = F16 ? NA1_16 : tc::NA1, NA = NA0 + NA1, NB = F16 ? NB_16 : tc::NB;
  extern __shared__ __align__(1024) char smem[];
  char* a_ring = smem + SmemLayout::a_off;
  char* b_ring = smem + SmemLayout::b_off;
  float* sig = reinterpret_cast<float*>(smem + SmemLayout::sig_off);
  float* part = reinterpret_cast<float*>(smem + SmemLayout::part_off);
  float* cst = reinterpret_cast<float*>(smem + SmemLayout::const_off);   // [n_layers][256] biases, then output rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SmemLayout::bar_off);
  // A_FULL / A_EMPTY: slots [0, NA0) belong to the first-layer ring, [NA0, NA) to the hidden-layer ring
  constexpr int A_FULL = 0, A_EMPTY = NA, B_FULL = 2 * NA, B_EMPTY = 2 * NA + NB, D_FULL = 2 * NA + 2 * NB,
                D_EMPTY = D_FULL + 2, N_BARS = D_EMPTY + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + N_BARS);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  constexpr int PTS = (MODE == 1) ? 64 : 128;
  constexpr int N_CHUNK = MLP_W / SLAB_K;
  const int64_t n_tiles_real = (prm.P + PTS - 1) / PTS;
  // every CTA runs the SAME number of tiles (padding with all-invalid tiles): the CTAs of a cluster consume the
  // weight-slab stream in lock step, so none may stop early
  const int64_t n_tiles = ((n_tiles_real + gridDim.x - 1) / gridDim.x) * gridDim.x;
  const FieldLayout& L = prm.lay;
  const int NL = prm.n_layers;

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) {
      // two half-row producers per row and slab (builders for slots < NA0, epilogue otherwise)
      mbar_init(bar(A_FULL + i), 256 * GR);
      mbar_init(bar(A_EMPTY + i), 1);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(bar(B_FULL + i), 1);
      mbar_init(bar(B_EMPTY + i), CLUSTER);   // released by the MMA warps of every CTA in the cluster
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(D_FULL + i), 1);
      mbar_init(bar(D_EMPTY + i), N_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const int n_out = (MODE == 2) ? 3 : 1;
    for (int i = tid; i < prm.n_layers * MLP_W; i += THREADS) cst[i] = prm.bias[i];
    for (int i = tid; i < n_out * MLP_W; i += THREADS) cst[prm.n_layers * MLP_W + i] = prm.w_out[i];
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of every CTA in the cluster are initialised before any multicast can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < WARP_BUILD) {
    // =========================================== epilogue ===========================================
    const int grp = (warp >> 2) & 1;           // which half of the chunks this warp drains
    const int hh = warp >> 3;                  // which 8 columns of a chunk
    const int r = (warp & 3) * 32 + (tid & 31);            // row == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float* sig_g = sig + grp * 2 * SIG_BUF;
    uint32_t g = 0;                            // global layer counter of this CTA
    uint32_t it = 0;                           // tile iteration of this CTA
    const bool prof = prm.dbg != nullptr;
    unsigned long long t_dfull = 0, t_aempty = 0;
    const long long t_begin = clock64();
    constexpr float K_EXP = 144.26950408889634f;       // 100 * log2(e)
    constexpr float K_LOG = 0.0069314718055994531f;    // ln(2) / 100
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int64_t p = tile * PTS + ((MODE == 1) ? (r & 63) : r);
      const bool valid = p < prm.P;
      // hidden-layer ring counter: slabs of layers 1..NL-1 of all tiles of this CTA, in MMA order
      uint32_t q = it * (uint32_t)(prm.slabs_per_tile - prm.n_slabs[0]);
      for (int l = 0; l < NL; ++l, ++g) {
        const uint32_t buf = g & 1u;
        mbar_wait_t(bar(D_FULL + buf), (g >> 1) & 1u, prof, t_dfull);
        tc_fence_after();
        const float* bl = cst + l * MLP_W + 8 * hh;
        const bool last = (l == NL - 1);
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        // software-pipelined TMEM reads: the load of this warp's NEXT chunk is in flight while the current one is
        // being activated / split / stored
        const uint32_t t_row = tmem_base + lane_base + buf * 256u + (uint32_t)(8 * hh);
        uint32_t raw[8];
        tmem_ld8_issue(t_row + (uint32_t)(grp * SLAB_K), raw);
#pragma unroll 1
        for (int j = grp; j < N_CHUNK; j += 2) {
          float v[8];
          tmem_ld_wait8(raw);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = F16 ? __uint_as_float(raw[i]) * (1.0f / F16_W_SCALE) : __uint_as_float(raw[i]);
          if (j + 2 < N_CHUNK) tmem_ld8_issue(t_row + (uint32_t)((j + 2) * SLAB_K), raw);
          if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i] + bl[j * 16 + i], 0.f);
          } else if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float z = v[i] + bl[j * 16 + i];
              const float y = __log2f(1.0f + fast_exp2(z * K_EXP)) * K_LOG;
              v[i] = z > 0.2f ? z : y;   // 100 z > 20
            }
          } else {
            // MODE 1: value rows (0..63) publish e = exp(100 z); both halves then work in parallel:
            // value: softplus = log(1 + e) / 100, tangent: sigma'(z) * (W t) with sigma' = e / (1 + e)
            float* sb = sig_g + ((j >> 1) & 1) * SIG_BUF;
            float e[8];
            if (r < 64) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[i] = v[i] + bl[j * 16 + i];
                e[i] = fast_exp2(v[i] * K_EXP);
              }
              *reinterpret_cast<float4*>(sb + r * 16 + 8 * hh) = make_float4(e[0], e[1], e[2], e[3]);
              *reinterpret_cast<float4*>(sb + r * 16 + 8 * hh + 4) = make_float4(e[4], e[5], e[6], e[7]);
            }
            if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
            else asm volatile("bar.sync 2, 256;" ::: "memory");
            if (r < 64) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float y = __log2f(1.0f + e[i]) * K_LOG;
                v[i] = (v[i] > 0.2f) ? v[i] : y;
              }
            } else {
              const float4 e0 = *reinterpret_cast<const float4*>(sb + (r - 64) * 16 + 8 * hh);
              const float4 e1 = *reinterpret_cast<const float4*>(sb + (r - 64) * 16 + 8 * hh + 4);
              const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                // 100 z > 20  <=>  e > exp(20)
                const float sg = ee[i] > 485165195.4097903f ? 1.f : __fdividef(ee[i], ee[i] + 1.f);
                v[i] *= sg;
              }
            }
          }
          if (!last) {
            const uint32_t qs = q + (uint32_t)j;          // slab counter; GR consecutive slabs share a ring slot
            const uint32_t gq = qs / GR;
            const uint32_t slot = NA0 + gq % NA1;
            char* a_dst = a_ring + slot * A_SLOT + (qs % GR) * A_SUB;
            mbar_wait_t(bar(A_EMPTY + slot), ((gq / NA1) & 1u) ^ 1u, prof, t_aempty);
            if constexpr (F16) store_a_half_f16(a_dst, r, hh, v);
            else store_a_half(a_dst, r, hh, v);
            fence_proxy_async();
            mbar_arrive(bar(A_FULL + slot));
          } else {
            const float* wo = cst + NL * MLP_W + j * 16 + 8 * hh;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o0 = fmaf(v[i], wo[i], o0);
              if (MODE == 2) {
                o1 = fmaf(v[i], wo[MLP_W + i], o1);
                o2 = fmaf(v[i], wo[2 * MLP_W + i], o2);
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar(D_EMPTY + buf));
        if (!last) {
          q += N_CHUNK;
        } else {
          // combine the four partial dot products of a row (helpers -> smem -> warp-group (grp 0, hh 0))
          const int helper = grp + 2 * hh;   // 0 = finaliser, 1..3 = helpers
          if (helper != 0) {
            float* pp = part + (helper - 1) * 3 * ROWS;
            pp[r] = o0;
            if (MODE == 2) {
              pp[ROWS + r] = o1;
              pp[2 * ROWS + r] = o2;
            }
          }
          asm volatile("bar.sync 3, 512;" ::: "memory");
          if (helper == 0) {
#pragma unroll
            for (int hlp = 0; hlp < 3; ++hlp) {
              const float* pp = part + hlp * 3 * ROWS;
              o0 += pp[r];
              if (MODE == 2) {
                o1 += pp[ROWS + r];
                o2 += pp[2 * ROWS + r];
              }
            }
            if (valid) {
              if (MODE == 2) {
                prm.out0[0 * prm.in.stride + p] = sigmoid_acc(o0 + __ldg(prm.b_out + 0));
                prm.out0[1 * prm.in.stride + p] = sigmoid_acc(o1 + __ldg(prm.b_out + 1));
                prm.out0[2 * prm.in.stride + p] = sigmoid_acc(o2 + __ldg(prm.b_out + 2));
              } else if (MODE == 0 || r < 64) {
                prm.out0[p] = o0 + __ldg(prm.b_out);
              } else if (prm.out1) {
                const int64_t ps = field_src(prm.in, p);
                prm.out1[0 * prm.in.stride + p] = o0 * prm.in.grad[0 * prm.in.stride + ps];
                prm.out1[1 * prm.in.stride + p] = o0 * prm.in.grad[1 * prm.in.stride + ps];
                prm.out1[2 * prm.in.stride + p] = o0 * prm.in.grad[2 * prm.in.stride + ps];
              }
            }
          }
          // helpers may only overwrite `part` after the finaliser has read it
          asm volatile("bar.sync 4, 512;" ::: "memory");
        }
      }
    }
    if (prof && (tid & 31) == 0) {
      atomicAdd(prm.dbg + 0, t_dfull);
      atomicAdd(prm.dbg + 1, t_aempty);
      atomicAdd(prm.dbg + 2, (unsigned long long)(clock64() - t_begin));
    }
  } else if (warp < WARP_MMA) {
    // =========================================== builder ============================================
    // two threads per row: half h owns features {8g + 4h + i : g < 4, i < 4} and writes columns [8h, 8h+8) of
    // every first-layer slab (see tc_first_layer_map for the column order)
    const int tb = tid - N_EPI;
    const int r = tb & (ROWS - 1);
    const int h = tb >> 7;
    uint32_t it = 0;
    const int off_feat = (MODE == 2) ? L.off_ft : L.off_fg;   // multiple of 16
    const int Lf = (MODE == 2) ? L.Lft : L.Lfg;
    const int Fdim = (MODE == 2) ? L.Fc : L.Fg;   // code width = n_fb blocks of FEAT columns
    const int n_fb = Fdim / FEAT;
    const float* __restrict__ table = (MODE == 2) ? prm.tab.fc : prm.tab.fg;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int64_t p = tile * PTS + ((MODE == 1) ? (r & 63) : r);
      const bool valid = p < prm.P;
      const bool tangent = (MODE == 1) && (r >= 64);
      uint32_t q = it * (uint32_t)prm.n_slabs[0];   // first-layer ring counter
      auto emit = [&](const float (&v)[8]) {
        const uint32_t gq = q / GR;                  // q counts slabs; GR consecutive slabs share a ring slot
        const uint32_t slot = gq % NA0;
        char* a_dst = a_ring + slot * A_SLOT + (q % GR) * A_SUB;
        if (q % GR == 0) mbar_wait(bar(A_EMPTY + slot), ((gq / NA0) & 1u) ^ 1u);
        if constexpr (F16) store_a_half_f16(a_dst, r, h, v);
        else store_a_half(a_dst, r, h, v);
        fence_proxy_async();
        mbar_arrive(bar(A_FULL + slot));
        ++q;
      };
      // ---- gather + blend of this half's 16 features of code block fb (registers).  Block 0 is issued before any
      //      ring wait so that its latency overlaps the previous tile ----
      float feat[16];
      float ds = 0.f;
      // where this point's neighbour data lives (geometry modes: optionally indirected, see FieldIn::index)
      const int64_t ps = (valid && MODE != 2) ? field_src(prm.in, p) : p;
      if (valid) ds = prm.in.ds[ps];
      auto gather = [&](int fb) {
#pragma unroll
        for (int i = 0; i < 16; ++i) feat[i] = 0.f;
        if (valid && !tangent) {
#pragma unroll
          for (int k = 0; k < KNN_K; ++k) {
            const int32_t sl = prm.in.slot[k * prm.in.stride + ps];
            const float w = prm.in.w[k * prm.in.stride + ps];
            const float* row = table + (int64_t)sl * Fdim + fb * FEAT + 4 * h;
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(row + 8 * g4));
              feat[g4 * 4 + 0] = __fadd_rn(feat[g4 * 4 + 0], __fmul_rn(a.x, w));
              feat[g4 * 4 + 1] = __fadd_rn(feat[g4 * 4 + 1], __fmul_rn(a.y, w));
              feat[g4 * 4 + 2] = __fadd_rn(feat[g4 * 4 + 2], __fmul_rn(a.z, w));
              feat[g4 * 4 + 3] = __fadd_rn(feat[g4 * 4 + 3], __fmul_rn(a.w, w));
            }
          }
        }
      };
      gather(0);
      // ---- head block: columns [0, off_feat): PE(ds) [, nabla, PE(view)], zero padded ----
      {
        float head[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) head[i] = 0.f;
        auto st = [&](int col, float v) { head[col] = v; };
        if (valid) {
          if (tangent) {
            store_scalar_pe_tangent(ds, 0, L.Ld, st);
          } else {
            store_scalar_pe(ds, 0, L.Ld, st);
            if (MODE == 2) {
              float dx, dy, dz;
              load_dir(prm.in, p, dx, dy, dz);
              store_vec3_pe(dx, dy, dz, L.off_view, L.Lv, st);
              if (L.use_nabla) {
                st(L.off_nabla + 0, prm.in.nabla[0 * prm.in.stride + p]);
                st(L.off_nabla + 1, prm.in.nabla[1 * prm.in.stride + p]);
                st(L.off_nabla + 2, prm.in.nabla[2 * prm.in.stride + p]);
              }
            }
          }
        }
        for (int s = 0; s < off_feat / SLAB_K; ++s) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = head[s * 16 + 8 * h + i];
          emit(v);
        }
      }
      for (int fb = 0; fb < n_fb; ++fb) {
        if (fb > 0) gather(fb);
        // ---- raw features: slab s holds groups g = 2s, 2s+1: columns [8h, 8h+8) = feat[g = 2s][0..3], feat[2s+1][0..3] ----
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = feat[s2 * 8 + i];
          emit(v);
        }
        // ---- bands: slab (b, g): columns [8h, 8h+8) = [sin(2^b f[g][0..3]), cos(2^b f[g][0..3])] ----
        float fr = 1.f;
        for (int b = 0; b < Lf; ++b) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float sn = 0.f, cs = 0.f;
              if (valid && !tangent) sincosf(feat[g4 * 4 + i] * fr, &sn, &cs);
              v[i] = sn;
              v[4 + i] = cs;
            }
            emit(v);
          }
          fr *= 2.f;
        }
      }
      if constexpr (GR > 1) {   // the first layer is padded with zero slabs (zero weights) to a whole number of ring steps
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        while (q < (it + 1u) * (uint32_t)prm.n_slabs[0]) emit(zero);
      }
    }
  } else if (warp == WARP_MMA) {
    // =========================================== MMA issuer =========================================
    if ((tid & 31) == 0) {
      uint32_t g = 0, q = 0, q0 = 0, q1 = 0;   // q: B ring; q0 / q1: first-layer / hidden-layer A rings
      const uint32_t a0 = smem_u32(a_ring), b0 = smem_u32(b_ring);
      const bool prof = prm.dbg != nullptr;
      unsigned long long t_dempty = 0, t_a0 = 0, t_a1 = 0, t_b = 0;
      const long long t_begin = clock64();
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < NL; ++l, ++g) {
          const uint32_t buf = g & 1u;
          mbar_wait_bt(bar(D_EMPTY + buf), ((g >> 1) & 1u) ^ 1u, prof, t_dempty);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * 256u;
          const int ns = prm.n_slabs[l];
          for (int j = 0; j < ns; j += GR, ++q) {   // q, q0, q1 count ring steps of GR slabs
            uint32_t sa, pa;
            if (l == 0) {
              sa = q0 % NA0;
              pa = (q0 / NA0) & 1u;
              ++q0;
            } else {
              sa = NA0 + q1 % NA1;
              pa = (q1 / NA1) & 1u;
              ++q1;
            }
            const uint32_t sb = q % NB;
            mbar_wait_bt(bar(A_FULL + sa), pa, prof, l == 0 ? t_a0 : t_a1);
            mbar_wait_bt(bar(B_FULL + sb), (q / NB) & 1u, prof, t_b);
            tc_fence_after();
            const uint32_t a_addr = a0 + sa * A_SLOT, b_addr = b0 + sb * B_SLOT;
            if constexpr (F16) {
#pragma unroll
              for (int sub = 0; sub < GR; ++sub) {
                // one K = 16 step per slab: two 8-column chunks; chunk stride: A 128 rows * 16 B, B 256 rows * 16 B
                const uint64_t a_hi = make_desc(a_addr + sub * A_SUB, ROWS * 16, 128);
                const uint64_t a_lo = make_desc(a_addr + sub * A_SUB + A_HALF, ROWS * 16, 128);
                const uint64_t b_hi = make_desc(b_addr + sub * B_SUB, MLP_W * 16, 128);
                const uint64_t b_lo = make_desc(b_addr + sub * B_SUB + B_HALF, MLP_W * 16, 128);
                mma_f16(d_tmem, a_lo, b_hi, (j | sub) ? 1u : 0u);   // small terms first
                mma_f16(d_tmem, a_hi, b_lo, 1u);
                mma_f16(d_tmem, a_hi, b_hi, 1u);
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                // two 4-column chunks per K=8 step; chunk stride: A 128 rows * 16 B, B 256 rows * 16 B
                const uint64_t a_hi = make_desc(a_addr + ks * 2 * (ROWS * 16), ROWS * 16, 128);
                const uint64_t a_lo = make_desc(a_addr + A_HALF + ks * 2 * (ROWS * 16), ROWS * 16, 128);
                const uint64_t b_hi = make_desc(b_addr + ks * 2 * (MLP_W * 16), MLP_W * 16, 128);
                const uint64_t b_lo = make_desc(b_addr + B_HALF + ks * 2 * (MLP_W * 16), MLP_W * 16, 128);
                mma_tf32(d_tmem, a_lo, b_hi, (j | ks) ? 1u : 0u);   // small terms first
                mma_tf32(d_tmem, a_hi, b_lo, 1u);
                mma_tf32(d_tmem, a_hi, b_hi, 1u);
              }
            }
            mma_commit(bar(A_EMPTY + sa));
            mma_commit_mc(bar(B_EMPTY + sb), (uint16_t)((1u << CLUSTER) - 1u));
          }
          mma_commit(bar(D_FULL + buf));
        }
      }
      if (prof) {
        atomicAdd(prm.dbg + 3, t_dempty);
        atomicAdd(prm.dbg + 4, t_a0);
        atomicAdd(prm.dbg + 5, t_a1);
        atomicAdd(prm.dbg + 6, t_b);
        atomicAdd(prm.dbg + 7, (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    // =========================================== weight loader ======================================
    if ((tid & 31) == 0) {
      uint32_t q = 0;
      const uint32_t b0 = smem_u32(b_ring);
      const uint32_t crank = cluster_ctarank();
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < NL; ++l) {
          const float* src = prm.w + prm.slab_off[l];
          for (int j = 0; j < prm.n_slabs[l]; j += GR, ++q) {   // GR consecutive slabs are contiguous in the image
            const uint32_t sb = q % NB;
            mbar_wait_backoff(bar(B_EMPTY + sb), ((q / NB) & 1u) ^ 1u);
            mbar_expect_tx(bar(B_FULL + sb), B_SLOT);
            // this CTA fetches 1/CLUSTER of the slab from L2 and multicasts it to every CTA of the cluster
            constexpr uint32_t PART = B_SLOT / CLUSTER;
            bulk_load_mc(b0 + sb * B_SLOT + crank * PART, src + (int64_t)j * (B_SUB / 4) + crank * (PART / 4), PART,
                         bar(B_FULL + sb), (uint16_t)((1u << CLUSTER) - 1u));
          }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA leaves while a peer may still multicast into its shared memory / barriers
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
// packing: W^T [K][256] fp32 (already weight-norm folded, OUR column order) -> per-slab hi/lo tf32 images
// ------------------------------------------------------------------------------------------------------------
__global__ void pack_tc_kernel(const float* __restrict__ wt /*[K_src][256]*/, const int32_t* __restrict__ kmap /*[K]*/,
                               int K, float* __restrict__ dst) {
  // one thread per (k, n)
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)K * MLP_W) return;
  const int n = (int)(t % MLP_W);
  const int k = (int)(t / MLP_W);
  const int ksrc = kmap[k];
  const float x = ksrc >= 0 ? wt[(int64_t)ksrc * MLP_W + n] : 0.f;
  const float hi = tc::tf32_rna(x);
  const float lo = tc::tf32_rna(x - hi);
  const int slab = k / tc::SLAB_K, kk = k % tc::SLAB_K;
  float* base = dst + (int64_t)slab * (tc::B_SLOT / 4);
  const int off = (kk / 4) * (MLP_W * 4) + n * 4 + (kk % 4);
  base[off] = hi;
  base[tc::B_HALF / 4 + off] = lo;
}

// fp16x3 variant: per slab [hi | lo] x [k/8][256][k%8] halves of 2^8 W (hi = fp16(x), lo = fp16(x - hi))
__global__ void pack_tc16_kernel(const float* __restrict__ wt /*[K_src][256]*/, const int32_t* __restrict__ kmap /*[K]*/,
                                 int K, float* __restrict__ dst) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)K * MLP_W) return;
  const int n = (int)(t % MLP_W);
  const int k = (int)(t / MLP_W);
  const int ksrc = kmap[k];
  const float x = (ksrc >= 0 ? wt[(int64_t)ksrc * MLP_W + n] : 0.f) * tc::F16_W_SCALE;
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  const int slab = k / tc::SLAB_K, kk = k % tc::SLAB_K;
  __half* base = reinterpret_cast<__half*>(dst + (int64_t)slab * (tc::B_SLOT16 / 4));
  const int off = (kk / 8) * (MLP_W * 8) + n * 8 + (kk % 8);
  base[off] = hi;
  base[tc::B_HALF16 / 2 + off] = lo;
}

// TC first-layer column k -> FFMA first-layer column (both in "our" orders; see FieldLayout).
// Builder half h (0/1) owns features F(g, h, i) = 8 g + 4 h + i (g < 4, i < 4) and columns [8h, 8h+8) of each slab:
//   raw slab s (2):       col 8h + 4u + i  = feature F(2s + u, h, i)            (u < 2)
//   band slab (b, g):     col 8h + i       = sin(2^b F(g, h, i)),  col 8h + 4 + i = cos(2^b F(g, h, i))
static std::vector<int32_t> tc_first_layer_map(const FieldLayout& L, bool color) {
  const int off = color ? L.off_ft : L.off_fg;
  const int Lf = color ? L.Lft : L.Lfg;
  std::vector<int32_t> m;
  for (int k = 0; k < off; ++k) m.push_back(k);                 // head block: identical order
  const int Fdim = color ? L.Fc : L.Fg;
  auto F = [](int g, int h, int i) { return 8 * g + 4 * h + i; };
  for (int fb = 0; fb < Fdim / FEAT; ++fb) {   // one run of (2 raw + 4 Lf band) slabs per 32-column code block
    const int f0 = fb * FEAT;
    for (int s = 0; s < 2; ++s)
      for (int h = 0; h < 2; ++h)
        for (int u = 0; u < 2; ++u)
          for (int i = 0; i < 4; ++i) m.push_back(off + f0 + F(2 * s + u, h, i));
    for (int b = 0; b < Lf; ++b)
      for (int g = 0; g < 4; ++g)
        for (int h = 0; h < 2; ++h) {
          for (int i = 0; i < 4; ++i) m.push_back(off + (1 + 2 * b) * Fdim + f0 + F(g, h, i));  // sin block
          for (int i = 0; i < 4; ++i) m.push_back(off + (2 + 2 * b) * Fdim + f0 + F(g, h, i));  // cos block
        }
  }
  return m;
}

static int pack_one(const MlpFfma& src, const FieldLayout& L, bool color, bool f16, MlpTc* dst, cudaStream_t stream) {
  int64_t total = 0;
  dst->total_slabs = 0;
  const int slot_floats = (f16 ? tc::B_SLOT16 : tc::B_SLOT) / 4;
  for (int l = 0; l < src.n_layers; ++l) {
    dst->n_slabs[l] = src.K[l] / tc::SLAB_K;
    if (f16) dst->n_slabs[l] = (int)align_up((int64_t)dst->n_slabs[l], (int64_t)tc::GR16);   // whole ring steps (zero slabs)
    dst->slab_off[l] = total;
    total += (int64_t)dst->n_slabs[l] * slot_floats;
    dst->total_slabs += dst->n_slabs[l];
  }
  NMB_CUDA_OK(dst->w.alloc(total));
  NMB_CUDA_OK(cudaMemsetAsync(dst->w.p, 0, (size_t)total * sizeof(float), stream));   // padding slabs stay zero
  for (int l = 0; l < src.n_layers; ++l) {
    std::vector<int32_t> kmap;
    if (l == 0) {
      kmap = tc_first_layer_map(L, color);
      NMB_CHECK((int)kmap.size() == src.K[0], "first-layer column map size mismatch");
    } else {
      kmap.resize(MLP_W);
      for (int i = 0; i < MLP_W; ++i) kmap[i] = i;
    }
    DevBuf<int32_t>& km = dst->kmap[l];   // kept in the field; pageable upload = staged before the call returns
    NMB_CUDA_OK(km.alloc((int64_t)kmap.size()));
    NMB_CUDA_OK(cudaMemcpyAsync(km.p, kmap.data(), kmap.size() * 4, cudaMemcpyHostToDevice, stream));
    const int64_t n = (int64_t)src.K[l] * MLP_W;
    if (f16) {
      pack_tc16_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(src.w.p + src.w_off[l], km.p, src.K[l],
                                                                      dst->w.p + dst->slab_off[l]);
    } else {
      pack_tc_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(src.w.p + src.w_off[l], km.p, src.K[l],
                                                                    dst->w.p + dst->slab_off[l]);
    }
    NMB_LAUNCH_OK();
  }
  return 0;
}

int pack_mlp_tc(const nmb_field_desc*, const FieldLayout& lay, nmb_field* f, cudaStream_t stream) {
  const bool f16 = f->engine == 2;   // geo_t / col_t hold the images of the engine the field was created for
  int rc = pack_one(f->geo_f, lay, false, f16, &f->geo_t, stream);
  if (rc) return rc;
  return pack_one(f->col_f, lay, true, f16, &f->col_t, stream);
}

template <int MODE, bool F16>
static int launch_tc(const nmb_field* f, const MlpFfma& fm, const MlpTc& tm, const FieldIn& in, int64_t P, float* out0,
                     float* out1, cudaStream_t stream) {
  if (P <= 0) return 0;
  NMB_CHECK(f->lay.off_fg <= 64 && f->lay.off_ft <= 64, "head block wider than 64 columns");
  tc::Params prm;
  prm.lay = f->lay;
  prm.in = in;
  prm.tab = FieldTables{f->fg.p, in.color_table ? in.color_table : f->fc.p};
  prm.w = tm.w.p;
  prm.bias = fm.b.p;
  prm.w_out = fm.w_out.p;
  prm.b_out = fm.b_out.p;
  prm.slabs_per_tile = 0;
  for (int i = 0; i < MAX_LAYERS; ++i) {
    prm.slab_off[i] = tm.slab_off[i];
    prm.n_slabs[i] = i < fm.n_layers ? tm.n_slabs[i] : 0;
    prm.slabs_per_tile += prm.n_slabs[i];
  }
  prm.n_layers = fm.n_layers;
  prm.P = P;
  prm.out0 = out0;
  prm.out1 = out1;
  prm.dbg = nullptr;
  static const bool want_prof = getenv("NMB_TC_PROFILE") != nullptr;
  unsigned long long* dbg_dev = nullptr;   // diagnostics only: allocated per launch, the launch is synchronous then
  if (want_prof) {
    NMB_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&dbg_dev), 8 * sizeof(unsigned long long)));
    NMB_CUDA_OK(cudaMemsetAsync(dbg_dev, 0, 8 * sizeof(unsigned long long), stream));
    prm.dbg = dbg_dev;
  }
  constexpr int PTS = (MODE == 1) ? 64 : 128;
  const size_t smem = tc::SmemLayoutT<F16, MODE == 1>::total;
  static DeviceOnce attr_once;
  NMB_CUDA_OK(attr_once.run([&] {
    return cudaFuncSetAttribute(mlp_tc_kernel<MODE, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }));
  const int64_t tiles = ceil_div(P, PTS);
  int64_t grid = tiles < (int64_t)sm_count() ? tiles : (int64_t)sm_count();
  grid = align_up(grid, tc::CLUSTER);
  if (grid > sm_count()) grid -= tc::CLUSTER;
  if (grid < tc::CLUSTER) grid = tc::CLUSTER;
  ProfScope prof(MODE == 2 ? PROF_COLOR : (MODE == 1 ? PROF_GEO_JVP : PROF_GEO), P, stream);
  mlp_tc_kernel<MODE, F16><<<(unsigned)grid, tc::THREADS, smem, stream>>>(prm);
  NMB_LAUNCH_OK();
  if (want_prof) {
    unsigned long long h[8];
    NMB_CUDA_OK(cudaMemcpyAsync(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost, stream));
    NMB_CUDA_OK(cudaStreamSynchronize(stream));
    const double ne = 16.0 * grid, nm = 1.0 * grid;  // 16 epilogue warps and 1 MMA thread per CTA report
    fprintf(stderr,
            "[tc-prof] mode %d P %lld grid %lld | epilogue warp avg cycles: total %.0f wait D_FULL %.0f wait A_EMPTY %.0f | "
            "MMA thread: total %.0f wait D_EMPTY %.0f wait A_FULL(L0) %.0f wait A_FULL(hidden) %.0f wait B_FULL %.0f\n",
            MODE, (long long)P, (long long)grid, h[2] / ne, h[0] / ne, h[1] / ne, h[7] / nm, h[3] / nm, h[4] / nm,
            h[5] / nm, h[6] / nm);
    cudaFree(dbg_dev);
  }
  return 0;
}

int launch_geo_tc(const nmb_field* f, const FieldIn& in, int64_t P, float* sdf, float* nabla, cudaStream_t stream) {
  if (f->engine == 2) {
    if (nabla) return launch_tc<1, true>(f, f->geo_f, f->geo_t, in, P, sdf, nabla, stream);
    return launch_tc<0, true>(f, f->geo_f, f->geo_t, in, P, sdf, nullptr, stream);
  }
  if (nabla) return launch_tc<1, false>(f, f->geo_f, f->geo_t, in, P, sdf, nabla, stream);
  return launch_tc<0, false>(f, f->geo_f, f->geo_t, in, P, sdf, nullptr, stream);
}

int launch_color_tc(const nmb_field* f, const FieldIn& in, int64_t P, float* rgb, cudaStream_t stream) {
  if (f->engine == 2) return launch_tc<2, true>(f, f->col_f, f->col_t, in, P, rgb, nullptr, stream);
  return launch_tc<2, false>(f, f->col_f, f->col_t, in, P, rgb, nullptr, stream);
}

}  // namespace nmb
