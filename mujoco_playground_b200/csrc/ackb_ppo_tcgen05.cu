// ackb_ppo_tcgen05.cu -- PPO minibatch gradient on the Blackwell tensor-core path (tcgen05.mma kind::tf32, accumulators in TMEM).
//
// Same contract as ppo_grad_kernel_tc (ackb_ppo.cu; include/ackb_ppo.h: ackb_ppo_minibatch_grad): for a minibatch of the rollout,
// forward of the two 79 -> 64 -> 64 tanh MLPs of SB3's MlpPolicy, PPO loss derivatives per sample, backward, and the weight
// gradients summed over the minibatch (stable_baselines3 PPO.train(), called through model.learn, src/rl/train.py:175-179).
//
// Mapping: one persistent CTA of 512 threads per SM walks its tiles of 128 samples (= UMMA M) ONCE; each tile goes through the policy
// net and then the value net (the observation rows are gathered from HBM once).  Thread = (sample row, quarter of the 64 columns); a
// thread's sample row is also its TMEM lane, so everything a thread computes stays in its lane:
//   * operands whose reduction index is the FEATURE (forward and backward-data GEMMs) are read by the tensor core straight from
//     TMEM: the gathered observation row goes from registers into TMEM (tcgen05.st), the epilogues write H1 / dZ2 back IN PLACE over
//     the accumulator they just consumed.
//   * operands whose reduction index is the SAMPLE (weight-gradient GEMMs) are the same arrays stored [sample][feature] in shared
//     memory in SWIZZLE_128B_BASE32B, the MN-major layout the tensor core accepts for 32-bit operands (umma.cuh; probed in
//     tools/microbench/umma_probe.cu: MN-major tf32 with any other layout type returns zeros).  The weight-gradient GEMMs are
//     computed TRANSPOSED (D = activation^T x delta, M = 128 rows of which 64 / 65 / 80 are real, N = 64), so every accumulator is 64
//     TMEM columns and both nets' accumulators stay resident over all tiles of the CTA.
//   * the weights (W1, W2, W2^T of one net, K-major SWIZZLE_128B, TF32-rounded) do not fit next to one tile for both nets, so a
//     small kernel writes their shared-memory images to global memory once per launch and the CTA streams them: as soon as the MMA
//     group that read a weight buffer has completed, one thread issues the bulk copy (cp.async.bulk, mbarrier completion) of the
//     OTHER net's matrix into it -- it lands several phases before its first use.
//   * six GEMM groups per tile and net, issued by ONE thread (74 tcgen05.mma):
//       (1) Z1  [128 x 64] = X  W1^T         A = X   (TMEM),            B = W1   K-major            -> TMEM ZA
//       (2) Z2  [128 x 64] = H1 W2^T         A = H1  (TMEM, over Z1),   B = W2                      -> TMEM ZB
//       (3) dH1 [128 x 64] = dZ2 W2          A = dZ2 (TMEM, over Z2),   B = W2^T                    -> TMEM ZA
//       (4) dW2^T [65 x 64] += [H1|1]^T dZ2  A = [H1 | ones] MN-major,  B = dZ2 MN-major            -> TMEM GW2[net]
//       (5) dW1^T [80 x 64] += X^T dZ1       A = X MN-major,            B = dZ1 MN-major            -> TMEM GW1[net]
//       (6) dW3^T [64 x 3]  += H2^T dOut     A = H2 MN-major,           B = [1 | dOut] MN-major, N = 16 -> TMEM GW3[net]
//     Bias gradients come out of the same GEMMs (column 79 of X and column 0 of the ones block are 1); the loss derivatives dOut
//     of a sample are written into columns 1, 2 of its row of the ones block.  Rows of an A operand beyond its real extent address
//     whatever follows the buffer (inside the allocation) and produce accumulator rows that are never read.
//   * epilogues (bias + tanh.approx, tanh', Gaussian / value heads, clipped-surrogate derivatives) on all 512 threads.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>
#include <utility>

#include "ackb.h"
#include "ackb_ppo_common.cuh"
#include "umma.cuh"

namespace ackb_ppo {
namespace {

using namespace umma;

constexpr int TM = 128;          // samples per tile
constexpr int NT5 = 512;
constexpr uint32_t BLK = 16384;  // one BASE32B block of 128 rows x 32 columns

// shared-memory map (bytes from a 1024-aligned base)
constexpr uint32_t O_DZ = 0;                  // [2 blocks] dZ2, then dZ1 [sample][64]
constexpr uint32_t O_H1 = O_DZ + 2 * BLK;     // [2 blocks] H1 [sample][64]
constexpr uint32_t O_ONE = O_H1 + 2 * BLK;    // [1 block]  column 0 = 1, columns 1, 2 = dOut of the current net (must follow H1: row 64 of (4))
constexpr uint32_t O_H2 = O_ONE + BLK;        // [2 blocks] H2 [sample][64]
constexpr uint32_t O_X = O_H2 + 2 * BLK;      // [3 blocks] observation tile [sample][80], column 79 = 1 for real samples
constexpr uint32_t O_W1 = O_X + 3 * BLK;      // [3 blocks of 64 rows] K-major SW128  W1 of the net in flight
constexpr uint32_t W1_BYTES = 3 * 8192, W2_BYTES = 2 * 8192;
constexpr uint32_t O_W2 = O_W1 + W1_BYTES;    // [2 blocks of 64 rows] W2   [out][in]
constexpr uint32_t O_W2T = O_W2 + W2_BYTES;   // [2 blocks of 64 rows] W2^T [in][out]
constexpr uint32_t IMG_BYTES = W1_BYTES + 2 * W2_BYTES;     // weight images of one net in global memory: W1 | W2 | W2^T
constexpr uint32_t O_F = O_W2T + W2_BYTES;    // floats from here
constexpr int F_B1 = 0, F_B2 = 128, F_W3 = 256, F_B3 = 448, F_LS = 452, F_SC = 456, F_PART = F_SC + TM * 5, F_END = F_PART + TM * 8;
constexpr uint32_t O_BAR = O_F + F_END * 4;          // mbarriers: MMA completion, 3 weight buffers; TMEM base
constexpr uint32_t SMEM_BYTES = O_BAR + 56 + 1024;   // + alignment slack
static_assert(SMEM_BYTES <= 232448, "shared memory budget of one SM (227 KB)");

// TMEM columns: X (A operand of (1) of both nets); ZA = Z1 -> H1 (in place) -> dH1; ZB = Z2 -> dZ2 (in place); accumulators per net
constexpr uint32_t C_X = 0, C_ZA = 80, C_ZB = 144, C_GW1 = 208, C_GW2 = 336, C_GW3 = 464;

// round to TF32 (10 mantissa bits, nearest, ties away from zero = cvt.rna.tf32.f32) with two integer instructions: the cvt runs on the
// quarter-rate conversion pipe, which the epilogues (6 conversions per activation element and net) would otherwise saturate
__device__ __forceinline__ float tf32r(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// read-only global load that the compiler may NOT sink towards its use (volatile asm keeps its place among the other volatile
// statements): the prefetch of the next tile must be issued a whole tile ahead of its consumer
__device__ __forceinline__ float ldg_early(const float* p) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_early4(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ int64_t ldg_early_s64(const int64_t* p) {
  int64_t v;
  asm volatile("ld.global.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
// byte offset of feature f (a multiple of 4: one 16-byte vector) of a row inside a BASE32B array of 32-column blocks
__device__ __forceinline__ uint32_t b32_feat(uint32_t rowoff, int rx, int f) {
  return (uint32_t)(f >> 5) * BLK + rowoff + (uint32_t)((((((f & 31) >> 3) ^ rx) & 3) << 5) + (f & 7) * 4);
}

// 16 consecutive floats of a 64-byte aligned shared-memory vector (4 x LDS.128; the compiler cannot prove the alignment itself)
__device__ __forceinline__ void lds16(const float* p, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 x = reinterpret_cast<const float4*>(p)[i];
    v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
  }
}

// This thread's 16 consecutive columns (col0 a multiple of 4) of its own row of a BASE32B array, as four 16-byte vectors.  Lanes are
// consecutive rows; rows r and r + 4 (neighbouring 4-row atoms) map to the same banks, so a quarter-warp would hit every bank twice.
// Rows of odd atoms (swap = 1) therefore take the two halves of each 32-byte chunk in the opposite order: the same instruction then
// touches the other 16 banks.
__device__ __forceinline__ void sts_row16(unsigned char* base, uint32_t rowoff, int rx, int swap, int col0, const float (&v)[16]) {
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const int o = c4 ^ 1;
    const float4 x = make_float4(swap ? v[4 * o] : v[4 * c4], swap ? v[4 * o + 1] : v[4 * c4 + 1], swap ? v[4 * o + 2] : v[4 * c4 + 2],
                                 swap ? v[4 * o + 3] : v[4 * c4 + 3]);
    *reinterpret_cast<float4*>(base + b32_feat(rowoff, rx, col0 + 4 * (c4 ^ swap))) = x;
  }
}
__device__ __forceinline__ void lds_row16(const unsigned char* base, uint32_t rowoff, int rx, int swap, int col0, float (&v)[16]) {
  float4 t[4];
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) t[c4] = *reinterpret_cast<const float4*>(base + b32_feat(rowoff, rx, col0 + 4 * (c4 ^ swap)));
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const float4 x = swap ? t[c4 ^ 1] : t[c4];
    v[4 * c4] = x.x; v[4 * c4 + 1] = x.y; v[4 * c4 + 2] = x.z; v[4 * c4 + 3] = x.w;
  }
}

#ifdef ACKB_T5_PROFILE
__device__ long long g_t5_prof[32];
#define T5_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long c_ = clock64(); g_t5_prof[i] += c_ - t5_last; t5_last = c_; } } while (0)
#else
#define T5_MARK(i) do { } while (0)
#endif

// shared-memory images of both nets' weights: [net][W1 | W2 | W2^T], TF32-rounded, K-major SWIZZLE_128B blocks of 64 rows
// (the same launch zeroes the gradient vector and the diagnostics, which the gradient kernel accumulates into)
// Blocks 16 .. of the grid (if any) compute the advantage statistics of the minibatch (adv_stats_block) at the same time.
constexpr int IMG_BLOCKS = 16;       // 2 nets x 8 slices of the elements, 256 threads each
__global__ void __launch_bounds__(256) ppo_t5_weight_images(const float* __restrict__ P, int D, unsigned char* __restrict__ img, float* __restrict__ grads,
                                                            float* __restrict__ diag, const float* __restrict__ adv, const int64_t* __restrict__ idx,
                                                            int mb, float* __restrict__ adv_out, double* __restrict__ adv_ws, int diag_keep) {
  asm volatile("griddepcontrol.launch_dependents;");      // the gradient kernel's set-up and first gather overlap this launch
  if ((int)blockIdx.x >= IMG_BLOCKS) {
    adv_stats_block(adv, idx, mb, adv_out, adv_ws, (int)blockIdx.x - IMG_BLOCKS, (int)gridDim.x - IMG_BLOCKS);
    return;
  }
  const int net = blockIdx.x & 1, slice = blockIdx.x >> 1;
  const Offsets o = offsets(D);
  const int tid = slice * blockDim.x + threadIdx.x, nth = (IMG_BLOCKS / 2) * blockDim.x;
  for (int i = net * nth + tid; i < o.total; i += 2 * nth) grads[i] = 0.0f;
  if (blockIdx.x == 0 && threadIdx.x < 5 && !diag_keep) diag[threadIdx.x] = 0.0f;
  unsigned char* im = img + (size_t)net * IMG_BYTES;
  const float* W1g = P + (net ? o.W1v : o.W1p);
  const float* W2g = P + (net ? o.W2v : o.W2p);
  for (int i = tid; i < H * 96; i += nth) {
    const int r = i / 96, k = i % 96;
    *reinterpret_cast<float*>(im + (k >> 5) * 8192 + sw128_off(r, k & 31)) = (k < D) ? tf32r(W1g[r * D + k]) : 0.0f;
  }
  for (int i = tid; i < H * H; i += nth) {
    const int r = i >> 6, k = i & 63;
    const float w = tf32r(W2g[i]);
    *reinterpret_cast<float*>(im + W1_BYTES + (k >> 5) * 8192 + sw128_off(r, k & 31)) = w;
    *reinterpret_cast<float*>(im + W1_BYTES + W2_BYTES + (r >> 5) * 8192 + sw128_off(k, r & 31)) = w;
  }
}

__global__ void __launch_bounds__(NT5, 1) ppo_grad_kernel_tcgen05(PpoArgs a, const unsigned char* __restrict__ wimg) {
#ifdef ACKB_T5_PROFILE
  long long t5_last = clock64();
#endif
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned base, computed as an OFFSET from the __shared__ symbol: rounding the pointer through an integer makes every
  // access below a generic LD / ST instead of LDS / STS
  unsigned char* const sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* const F = reinterpret_cast<float*>(sm + O_F);
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(sm + O_BAR);          // MMA groups
  uint64_t* const mbw = reinterpret_cast<uint64_t*>(sm + O_BAR + 8);       // [3] weight buffers W1, W2, W2^T
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(sm + O_BAR + 32);
  uint64_t* const mbar_bg = reinterpret_cast<uint64_t*>(sm + O_BAR + 40);     // (4) and (6): two issuing lanes commit to it
  uint64_t* const mbar_bg1 = reinterpret_cast<uint64_t*>(sm + O_BAR + 48);    // (5) of the policy net behind (1) of the value net
  const uint32_t sb = smem_u32(sm);

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int q4 = warp & 3, part = warp >> 2;        // TMEM lane quarter of this warp, column quarter of this thread
  const int row = 32 * q4 + lane;                   // sample row of the tile = TMEM lane
  const uint32_t rowoff = (uint32_t)((row >> 2) * 512 + (row & 3) * 128);   // BASE32B: 4-row atoms, 32-byte chunks XOR (row & 3)
  const int rx = row & 3, swp = (row >> 2) & 1;
  const int D = a.D;
  const Offsets o = offsets(D);
  const float* P = a.params;
  const int ntiles = (a.mb + TM - 1) / TM;

  // ---- one-time setup: TMEM, mbarriers, ones block, small vectors of both nets
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (t == 32) {
    mbar_init(mbar, 1);
    mbar_init(mbar_bg, 2);
    mbar_init(mbar_bg1, 1);
    for (int i = 0; i < 3; ++i) mbar_init(mbw + i, 1);
    mbar_fence_init();
  }
  for (int i = t; i < TM * 32; i += NT5) *reinterpret_cast<float*>(sm + O_ONE + b32_off(i >> 5, i & 31)) = ((i & 31) == 0) ? 1.0f : 0.0f;
  for (int i = t; i < 3 * H; i += NT5) F[F_W3 + i] = (i < 2 * H) ? P[o.Wa + i] : P[o.Wv + (i - 2 * H)];
  if (t < 128) {
    const int net = t >> 6, r = t & 63;
    F[F_B1 + t] = P[(net ? o.b1v : o.b1p) + r];
    F[F_B2 + t] = P[(net ? o.b2v : o.b2p) + r];
  }
  if (t < 2) { F[F_B3 + t] = P[o.ba + t]; F[F_LS + t] = P[o.ls + t]; }
  if (t == 2) F[F_B3 + 2] = P[o.bv];
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = *tmem_slot;
  const uint32_t tlane = tb + ((uint32_t)(32 * q4) << 16);     // this warp's TMEM lane quarter

  // weight streaming (elected lane of warp 0): buffer `which` (0 W1, 1 W2, 2 W2^T) <- image of `net`
  uint32_t wph = 0;                                            // parity of the next completion of each weight barrier (bit per buffer), tracked by every thread
  auto load_w = [&](int which, int net) {
    const uint32_t bytes = which == 0 ? W1_BYTES : W2_BYTES;
    const uint32_t dst = sb + (which == 0 ? O_W1 : (which == 1 ? O_W2 : O_W2T));
    const unsigned char* src = wimg + (size_t)net * IMG_BYTES + (which == 0 ? 0u : (which == 1 ? W1_BYTES : W1_BYTES + W2_BYTES));
    mbar_expect_tx(mbw + which, bytes);
    bulk_g2s(dst, src, bytes, mbw + which);
  };
  const float inv_mb = 1.0f / (float)a.mb;
  const float ls0 = F[F_LS], ls1 = F[F_LS + 1];
  const float iv0 = expf(-2.0f * ls0), iv1 = expf(-2.0f * ls1);

  // per-thread sums over all tiles of this CTA (threads with part == 0: one per sample row)
  float gb3[3] = {0.f, 0.f, 0.f}, gls0 = 0.f, gls1 = 0.f, d_pg = 0.f, d_vl = 0.f, d_kl = 0.f, d_cf = 0.f;
  uint32_t phase = 0;
  float* G = a.grads;

  // issue a group of MMAs once every thread's shared-memory / TMEM accesses of the previous phase are done, wait for it.  The group
  // is issued by ONE elected lane of warp 0 inside a warp-uniform branch (elect.sync: ptxas then emits the UTCHMMA sequence straight,
  // a threadIdx-divergent branch makes it wrap every instruction in a lane loop).
  // ST: this phase wrote TMEM (tcgen05.st);  px_fence: a shared-memory operand written by the threads is read by this group (generic
  // -> async proxy fence; a later fence also covers the earlier writes of the thread, so groups that read only TMEM and the
  // bulk-copied weights skip it)
  // `issue` = the MMAs the next epilogue needs (issued by the elected lane of warp 0, completion awaited here); `bg1` / `bg2` = MMAs off
  // the critical chain, issued at the same time by the elected lanes of warps 1 and 2 (other schedulers; the accumulators of the three
  // issuers are disjoint, so their relative order in the tensor pipe does not matter) -- an issuing lane is busy until the pipe has
  // taken its MMAs, and the whole CTA waits for the slowest warp at the next group.  Each background lambda commits to its own
  // barrier; the operands of background MMAs stay untouched until that barrier has been waited for.
  uint32_t phase_bg = 0, phase_bg1 = 0;
  auto wait_bg = [&]() {
    mbar_wait(mbar_bg, phase_bg);
    phase_bg ^= 1u;
    fence_after_sync();
  };
  auto wait_bg1 = [&]() {
    mbar_wait(mbar_bg1, phase_bg1);
    phase_bg1 ^= 1u;
    fence_after_sync();
  };
  auto run_mma = [&](auto st, bool px_fence, auto grp, auto&& issue, auto&& bg1, auto&& bg2) {
    if (decltype(st)::value) tmem_st_wait();
    T5_MARK(15);
    if (px_fence) fence_proxy_async();
    fence_before_sync();
    T5_MARK(12);
    __syncthreads();
    T5_MARK(13);
    if (warp < 3) {
      if (elect_one()) {
        fence_after_sync();
        if (warp == 0) { issue(); commit(mbar); }
        else if (warp == 1) bg1();
        else bg2();
      }
      __syncwarp();
    }
    T5_MARK(16 + decltype(grp)::value);
    mbar_wait(mbar, phase);
    phase ^= 1u;
    fence_after_sync();
    T5_MARK(24 + decltype(grp)::value);
  };
  auto none = [] {};
  using Yes = std::true_type;
  using No = std::false_type;
  // descriptor low words of the operands (K step 0); the CTA owns all 512 TMEM columns, so its TMEM base is lane 0, column 0 and the
  // accumulator / operand TMEM addresses below are plain constants for the issuing thread
  if (tb != 0u) __trap();
  const uint32_t lo_w1 = kmajor_lo(sb + O_W1), lo_w2 = kmajor_lo(sb + O_W2), lo_w2t = kmajor_lo(sb + O_W2T);
  const uint32_t lo_dz = mn32_lo(sb + O_DZ, BLK), lo_h1 = mn32_lo(sb + O_H1, BLK), lo_h2 = mn32_lo(sb + O_H2, BLK), lo_one = mn32_lo(sb + O_ONE, BLK),
                 lo_x = mn32_lo(sb + O_X, BLK);

  // Gather of a tile's observation rows into registers, one tile ahead of its use.  Four threads share a row (thread t: row t >> 2,
  // 16-byte vectors (t & 3) + 4 j): one load instruction of a warp covers 8 rows x 64 contiguous bytes, a quarter of the L1 tag
  // look-ups of a one-row-part-per-thread mapping.  16-byte loads when the row pitch allows (pitch 80: DRAM sectors are whole), else
  // 4-byte loads.  Threads t < 128 also fetch the five scalars of row t.
  const bool vec = (a.pitch % 4 == 0) && a.pitch >= KP && ((reinterpret_cast<uintptr_t>(a.obs) & 15) == 0);
  const int PT = a.pitch;
  const int gr = t >> 2, gc = t & 3;
  const uint32_t g_rowoff = (uint32_t)((gr >> 2) * 512 + (gr & 3) * 128);
  float px[20], psn[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  int64_t prow, prow_sc;          // global rows of the tile fetched NEXT (loaded one fetch earlier: no dependent-load stall)
  int64_t qrow, qrow_sc;          // ... and of the tile after it: its rows are pulled into L2 a whole tile before they are loaded
  auto row_index = [&](int tile, int r) -> int64_t {
    const int sbase = tile * TM, ns = tile < ntiles ? min(TM, a.mb - sbase) : 0;
    // (a plain load would be sunk by the compiler to its first use a tile later, and the observation loads would wait for it)
    return (r < ns) ? (a.idx ? ldg_early_s64(a.idx + sbase + r) : (int64_t)(sbase + r)) : (int64_t)-1;
  };
  // The loads of a tile are issued in several places of the previous tile's body.  Their addresses come from registers computed
  // ONCE per tile (set_sources) from the row numbers, which were loaded a tile earlier: an address that still depended on a load in
  // flight would wait on a scoreboard it shares with the observation loads issued just before, i.e. for a full DRAM round trip per
  // group of loads.
  const float* gsrc = a.obs;
  bool glive = false;
  int64_t gsc = -1;
  auto set_sources = [&]() {
    glive = prow >= 0;
    gsrc = a.obs + (glive ? prow : 0) * (int64_t)PT + 4 * gc;
    gsc = prow_sc;
  };
  auto fetch_part = [&](int j) {      // columns 16 j + 4 gc .. + 3 of row gr
    const float* src = gsrc + 16 * j;
    if (vec) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (glive) v = ldg_early4(src);
      px[4 * j] = v.x; px[4 * j + 1] = v.y; px[4 * j + 2] = v.z; px[4 * j + 3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) px[4 * j + i] = (glive && 16 * j + 4 * gc + i < D) ? ldg_early(src + i) : 0.0f;
    }
  };
  auto fetch_scalars = [&]() {
    if (t < TM && gsc >= 0) {
      psn[0] = ldg_early(&a.act[gsc * 2]); psn[1] = ldg_early(&a.act[gsc * 2 + 1]); psn[2] = ldg_early(&a.old_logp[gsc]);
      psn[3] = ldg_early(&a.adv[gsc]); psn[4] = ldg_early(&a.ret[gsc]);
    }
  };
  prow = row_index(blockIdx.x, gr); prow_sc = row_index(blockIdx.x, t);
  set_sources();
#pragma unroll
  for (int j = 0; j < 5; ++j) fetch_part(j);
  fetch_scalars();
  prow = row_index(blockIdx.x + gridDim.x, gr); prow_sc = row_index(blockIdx.x + gridDim.x, t);
  qrow = row_index(blockIdx.x + 2 * gridDim.x, gr); qrow_sc = row_index(blockIdx.x + 2 * gridDim.x, t);
  // L2 prefetch of one observation row (bulk prefetch: handled by the copy engine, no register, no L1 miss-queue entry of the warp).
  // The loads of the row then find it in L2: a third of the DRAM latency, so the L1 miss queue turns over three times faster and the
  // issuing warps are not held up behind it
  auto prefetch_row = [&](int64_t r) {
    if (gc == 0 && r >= 0) {
      const float* p = a.obs + r * (int64_t)PT;
      if (vec) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(KP * 4) : "memory");
      else {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 64));
      }
    }
  };
  prefetch_row(prow);
  // Everything above reads only what earlier launches produced (parameters, rollout arrays, the epoch's permutation) and overlaps
  // the prologue launch (weight images, gradient / diagnostics zeroing, advantage statistics), of which this grid is a programmatic
  // dependent.  From here on its results are needed: every thread waits (the gradient atomics at the end must follow the zeroing).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (warp == 0) {
    if (elect_one()) { load_w(0, 0); load_w(1, 0); load_w(2, 0); }
    __syncwarp();
  }
  const float adv_mean = a.adv_stats[0], adv_istd = 1.0f / (a.adv_stats[1] + 1e-8f);
  T5_MARK(10);

  bool first = true;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int ns = min(TM, a.mb - tile * TM);
    const bool more = tile + (int)gridDim.x < ntiles;      // CTA-uniform
    // ---- this tile's observation rows: registers -> TF32 -> shared memory in the operand layout (A operand of (5)); columns >= D
    // are zero except column 79 = 1 on real samples (bias gradient of layer 1); the row scalars go to shared memory as well
    {
      const float one = (gr < ns) ? 1.0f : 0.0f;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = 16 * j + 4 * gc + i;
          v[i] = (f < D) ? tf32r(px[4 * j + i]) : ((f == KP - 1) ? one : 0.0f);
        }
        *reinterpret_cast<float4*>(sm + O_X + b32_feat(g_rowoff, gr & 3, 16 * j + 4 * gc)) = make_float4(v[0], v[1], v[2], v[3]);
      }
      if (t < TM) {
#pragma unroll
        for (int i = 0; i < 5; ++i) F[F_SC + i * TM + t] = psn[i];
      }
    }
    T5_MARK(21);
    // ---- prefetch of the next tile.  Two constraints place the loads: (i) a burst of all 5 x 16-byte loads per thread overruns the
    // L1 miss queue and blocks the issuing warps for ~5 k cycles, so they are trickled out one or two at a time; (ii)
    // fence.proxy.async (before every MMA group that reads thread-written shared memory) waits for the thread's outstanding global
    // loads, so a load issued shortly before a fenced group puts a DRAM round trip on the critical path.  The groups (1) and (2) read
    // only TMEM and the bulk-copied weights and are not fenced: loads go right here and into the first two phases of each net.
    if (more) {
      set_sources();           // rows of the next tile (numbers loaded during the previous tile)
      fetch_part(0); fetch_part(1);
      prefetch_row(qrow);      // rows of the tile after it -> L2
    }
    __syncthreads();
    T5_MARK(0);
    // ---- this thread's part of its own row (columns 20 part .. + 19): shared memory -> TMEM (A operand of (1) of both nets)
    {
      float v[16], w4[4];
      lds_row16(sm + O_X, rowoff, rx, swp, 20 * part, v);      // (any 4-aligned first column works: the vectors are swapped in pairs)
      {
        const float4 x = *reinterpret_cast<const float4*>(sm + O_X + b32_feat(rowoff, rx, 20 * part + 16));
        w4[0] = x.x; w4[1] = x.y; w4[2] = x.z; w4[3] = x.w;
      }
      tmem_st16(tlane + C_X + 20 * part, v);
      tmem_st4(tlane + C_X + 20 * part + 16, w4);
    }
    T5_MARK(22);
#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
      const bool reload = (net == 0) || more;      // the other net's weights are needed again
      const uint32_t acc0 = first ? 0u : 1u;
      // ---- (1) Z1 = X W1^T; for the value net the group starts with (5) of the policy net, dW1^T += X^T dZ1, which nothing of the
      // value net's chain depends on (its operands X, dZ1 are overwritten only after later groups have completed)
      {
        const uint32_t par = (wph >> 0) & 1u;
        run_mma(Yes{}, net == 1, std::integral_constant<int, 0>{}, [&] {
          mbar_wait(mbw + 0, par);
          const uint32_t id = idesc_tf32(128, 64, 0, 0);
#pragma unroll
          for (int k = 0; k < KP / 8; ++k) mma_tf32_ts_lohi(C_ZA, C_X + 8 * k, lo_w1 + (k >> 2) * 512 + (k & 3) * 2, KMAJOR_HI, id, k > 0);
        }, [&] {
          if (net == 0) return;
          const uint32_t id5 = idesc_tf32(128, 64, 1, 1);      // (5) of the policy net runs next to (1) of the value net
#pragma unroll
          for (int k = 0; k < TM / 8; ++k) mma_tf32_lohi(C_GW1, lo_x + 64 * k, MN32_HI, lo_dz + 64 * k, MN32_HI, id5, k > 0 ? 1u : acc0);
          commit(mbar_bg1);
        }, none);
        wph ^= 1u;
      }
      if (warp == 0 && reload) { if (elect_one()) load_w(0, net ^ 1); __syncwarp(); }
      if (more && net == 1) fetch_part(3);
      T5_MARK(1);
      // ---- H1 = tanh(Z1 + b1): back into TMEM in place (A operand of (2)) and into shared memory (A operand of (4), tanh' later)
      {
        float v[16];
        tmem_ld16(tlane + C_ZA + 16 * part, v);
        float b1[16];
        lds16(&F[F_B1 + 64 * net + 16 * part], b1);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tf32r(tanh_fast(v[j] + b1[j]));
        tmem_st16(tlane + C_ZA + 16 * part, v);
        sts_row16(sm + O_H1, rowoff, rx, swp, 16 * part, v);
      }
      if (more) {
        if (net == 0) fetch_part(2);
        else {
          fetch_part(4); fetch_scalars();
          prow = qrow; prow_sc = qrow_sc;
          qrow = row_index(tile + 3 * (int)gridDim.x, gr); qrow_sc = row_index(tile + 3 * (int)gridDim.x, t);
        }
      }
      T5_MARK(2);
      // ---- (2) Z2 = H1 W2^T
      {
        const uint32_t par = (wph >> 1) & 1u;
        run_mma(Yes{}, false, std::integral_constant<int, 1>{}, [&] {
          mbar_wait(mbw + 1, par);
          const uint32_t id = idesc_tf32(128, 64, 0, 0);
#pragma unroll
          for (int k = 0; k < 8; ++k) mma_tf32_ts_lohi(C_ZB, C_ZA + 8 * k, lo_w2 + (k >> 2) * 512 + (k & 3) * 2, KMAJOR_HI, id, k > 0);
        }, none, none);
        wph ^= 2u;
        if (net == 1) wait_bg1();     // (5) of the policy net must be complete before dZ1 is overwritten (it is, long since)
      }
      if (warp == 0 && reload) { if (elect_one()) load_w(1, net ^ 1); __syncwarp(); }
      T5_MARK(3);
      // ---- heads: H2 = tanh(Z2 + b2) of this thread's 16 columns (registers for tanh', shared memory for (6)), partial dot
      // products with the head weights
      float h2v[16];
      const float* wa = &F[F_W3 + (net ? 2 * H : 0) + 16 * part];
      {
        float v[16], b2[16], w0[16], p0 = 0.f, p1 = 0.f;
        tmem_ld16(tlane + C_ZB + 16 * part, v);
        lds16(&F[F_B2 + 64 * net + 16 * part], b2);
        lds16(wa, w0);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float h2 = tanh_fast(v[j] + b2[j]);
          h2v[j] = h2;
          p0 = fmaf(h2, w0[j], p0);
          v[j] = tf32r(h2);
        }
        if (net == 0) {
          lds16(wa + H, w0);
#pragma unroll
          for (int j = 0; j < 16; ++j) p1 = fmaf(h2v[j], w0[j], p1);
        }
        sts_row16(sm + O_H2, rowoff, rx, swp, 16 * part, v);
        F[F_PART + (part * 2) * TM + row] = p0; F[F_PART + (part * 2 + 1) * TM + row] = p1;      // [part, head][row]: conflict free
      }
      __syncthreads();
      T5_MARK(4);
      // ---- PPO loss derivatives of this thread's sample: computed by all four threads of the row (no second barrier), summed into the
      // per-launch totals and written to the ones block by the thread with part == 0
      float d0 = 0.f, d1 = 0.f;
      if (row < ns) {
        const float* PR = &F[F_PART + row];
        const float* SCp = &F[F_SC + row];
        const float SC[5] = {SCp[0], SCp[TM], SCp[2 * TM], SCp[3 * TM], SCp[4 * TM]};
        const float s0 = (PR[0] + PR[2 * TM]) + (PR[4 * TM] + PR[6 * TM]), s1 = (PR[TM] + PR[3 * TM]) + (PR[5 * TM] + PR[7 * TM]);
        if (net == 0) {
          const float m0 = s0 + F[F_B3], m1 = s1 + F[F_B3 + 1];
          const float e0 = SC[0] - m0, e1 = SC[1] - m1;
          const float q0 = e0 * e0 * iv0, q1 = e1 * e1 * iv1;
          const float logp = -0.5f * q0 - ls0 - 0.9189385332046727f - 0.5f * q1 - ls1 - 0.9189385332046727f;
          const float A_ = (SC[3] - adv_mean) * adv_istd;
          const float lr = logp - SC[2];
          const float r = expf(lr);
          const float rc = fminf(fmaxf(r, 1.0f - a.clip), 1.0f + a.clip);
          const float u1 = A_ * r, u2 = A_ * rc;
          const float dlogp = (u1 <= u2) ? -A_ * r : 0.0f;      // d(-min(u1, u2)) / d logp (the clipped branch has zero slope)
          d0 = dlogp * e0 * iv0 * inv_mb; d1 = dlogp * e1 * iv1 * inv_mb;
          if (part == 0) {
            gls0 += dlogp * (q0 - 1.0f) * inv_mb; gls1 += dlogp * (q1 - 1.0f) * inv_mb;
            d_pg += -fminf(u1, u2); d_kl += (r - 1.0f) - lr; d_cf += (fabsf(r - 1.0f) > a.clip) ? 1.0f : 0.0f;
            gb3[0] += d0; gb3[1] += d1;
          }
        } else {
          const float v = s0 + F[F_B3 + 2], R = SC[4];
          d0 = a.vf_coef * 2.0f * (v - R) * inv_mb;
          if (part == 0) { d_vl += (v - R) * (v - R); gb3[2] += d0; }
        }
      }
      if (part == 0) {
        float* ob = reinterpret_cast<float*>(sm + O_ONE + rowoff + (uint32_t)(rx << 5));      // columns 0..7 of this row of the ones block
        ob[1] = tf32r(d0); ob[2] = tf32r(d1);
      }
      T5_MARK(5);
      // ---- dZ2 = (dOut W3) (1 - H2^2): into TMEM in place (A operand of (3)) and shared memory (B operand of (4))
      {
        float v[16], w0[16];
        lds16(wa, w0);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = d0 * w0[j];
        if (net == 0) {
          lds16(wa + H, w0);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(d1, w0[j], v[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tf32r(v[j] * (1.0f - h2v[j] * h2v[j]));
        tmem_st16(tlane + C_ZB + 16 * part, v);
        sts_row16(sm + O_DZ, rowoff, rx, swp, 16 * part, v);
      }
      T5_MARK(6);
      // ---- (3) dH1 = dZ2 W2,  (4) dW2^T += [H1 | 1]^T dZ2,  (6) dW3^T += H2^T [1 | dOut]
      {
        const uint32_t par = (wph >> 2) & 1u;
        run_mma(Yes{}, true, std::integral_constant<int, 2>{}, [&] {
          mbar_wait(mbw + 2, par);
          const uint32_t id3 = idesc_tf32(128, 64, 0, 0);
#pragma unroll
          for (int k = 0; k < 8; ++k) mma_tf32_ts_lohi(C_ZA, C_ZB + 8 * k, lo_w2t + (k >> 2) * 512 + (k & 3) * 2, KMAJOR_HI, id3, k > 0);
        }, [&] {
          const uint32_t id4 = idesc_tf32(128, 64, 1, 1);
          const uint32_t gw2 = C_GW2 + 64 * net;
#pragma unroll
          for (int k = 0; k < TM / 8; ++k) mma_tf32_lohi(gw2, lo_h1 + 64 * k, MN32_HI, lo_dz + 64 * k, MN32_HI, id4, k > 0 ? 1u : acc0);
          commit(mbar_bg);
        }, [&] {
          const uint32_t id6 = idesc_tf32(128, 16, 1, 1);
          const uint32_t gw3 = C_GW3 + 16 * net;
#pragma unroll
          for (int k = 0; k < TM / 8; ++k) mma_tf32_lohi(gw3, lo_h2 + 64 * k, MN32_HI, lo_one + 64 * k, MN32_HI, id6, k > 0 ? 1u : acc0);
          commit(mbar_bg);
        });
        wph ^= 4u;
      }
      if (warp == 0 && reload) { if (elect_one()) load_w(2, net ^ 1); __syncwarp(); }
      T5_MARK(7);
      // ---- dZ1 = dH1 (1 - H1^2), computed while (4) and (6) still run; stored over dZ2 once they have completed
      {
        float v[16], h1[16];
        tmem_ld16(tlane + C_ZA + 16 * part, v);
        lds_row16(sm + O_H1, rowoff, rx, swp, 16 * part, h1);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tf32r(v[j] * (1.0f - h1[j] * h1[j]));
        wait_bg();
        sts_row16(sm + O_DZ, rowoff, rx, swp, 16 * part, v);
      }
      T5_MARK(8);
      // ---- (5) dW1^T += X^T dZ1 of the value net (the policy net's rides in the value net's first group); the next tile overwrites X
      if (net == 1) {
        run_mma(No{}, true, std::integral_constant<int, 3>{}, [&] {
          const uint32_t id5 = idesc_tf32(128, 64, 1, 1);
#pragma unroll
          for (int k = 0; k < TM / 8; ++k) mma_tf32_lohi(C_GW1 + 64, lo_x + 64 * k, MN32_HI, lo_dz + 64 * k, MN32_HI, id5, k > 0 ? 1u : acc0);
        }, none, none);
      }
      T5_MARK(9);
    }
    first = false;
  }

  // ---- weight-gradient accumulators: TMEM -> global.  Accumulator row = input feature (this thread's lane), columns = output unit
#pragma unroll 1
  for (int net = 0; net < 2; ++net) {
    if (q4 < 3) {
      const int k = row;
      float v[16];
      tmem_ld16(tlane + C_GW1 + 64 * net + 16 * part, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = 16 * part + j;
        if (k < D) atomicAdd(&G[(net ? o.W1v : o.W1p) + i * D + k], v[j]);
        else if (k == KP - 1) atomicAdd(&G[(net ? o.b1v : o.b1p) + i], v[j]);
      }
      tmem_ld16(tlane + C_GW2 + 64 * net + 16 * part, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = 16 * part + j;
        if (k < H) atomicAdd(&G[(net ? o.W2v : o.W2p) + i * H + k], v[j]);
        else if (k == H) atomicAdd(&G[(net ? o.b2v : o.b2p) + i], v[j]);
      }
    }
    if (part == 0 && q4 < 2) {      // head weights: row = hidden unit, column 1 + head
      float v[16];
      tmem_ld16(tlane + C_GW3 + 16 * net, v);
      if (net == 0) { atomicAdd(&G[o.Wa + row], v[1]); atomicAdd(&G[o.Wa + H + row], v[2]); }
      else atomicAdd(&G[o.Wv + row], v[1]);
    }
  }
  T5_MARK(11);
  if (part == 0) {           // per-sample sums: one value per thread of warps 0..3
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      gls0 += __shfl_xor_sync(0xffffffffu, gls0, off); gls1 += __shfl_xor_sync(0xffffffffu, gls1, off);
      d_pg += __shfl_xor_sync(0xffffffffu, d_pg, off); d_vl += __shfl_xor_sync(0xffffffffu, d_vl, off);
      d_kl += __shfl_xor_sync(0xffffffffu, d_kl, off); d_cf += __shfl_xor_sync(0xffffffffu, d_cf, off);
      gb3[0] += __shfl_xor_sync(0xffffffffu, gb3[0], off); gb3[1] += __shfl_xor_sync(0xffffffffu, gb3[1], off);
      gb3[2] += __shfl_xor_sync(0xffffffffu, gb3[2], off);
    }
    if (lane == 0) {
      atomicAdd(&G[o.ls], gls0); atomicAdd(&G[o.ls + 1], gls1);
      atomicAdd(&G[o.ba], gb3[0]); atomicAdd(&G[o.ba + 1], gb3[1]); atomicAdd(&G[o.bv], gb3[2]);
      atomicAdd(&a.diag[0], d_pg * inv_mb); atomicAdd(&a.diag[1], d_vl * inv_mb);
      atomicAdd(&a.diag[3], d_kl * inv_mb); atomicAdd(&a.diag[4], d_cf * inv_mb);
    }
  }
  if (blockIdx.x == 0 && t == 0) {
    // entropy of the state-independent Gaussian: sum_j (0.5 + 0.5 log 2 pi + log_std_j); -ent_coef * H enters the loss
    if (a.diag_keep) atomicAdd(&a.diag[2], 2.0f * 1.4189385332046727f + ls0 + ls1);
    else a.diag[2] = 2.0f * 1.4189385332046727f + ls0 + ls1;
    atomicAdd(&G[o.ls], -a.ent_coef); atomicAdd(&G[o.ls + 1], -a.ent_coef);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}


// =================================================================================================================
// Rollout forward (SB3 policy.forward() inside collect_rollouts) on the same path: both MLPs of a 128-row tile, Gaussian sample,
// log-probability and value.  Rows are contiguous here, so a tile is ONE bulk copy (cp.async.bulk, double buffered, issued a tile
// ahead by one thread: no registers, no load instructions in the epilogue warps).  Two MMA groups per tile:
//   (1) [Z1p | Z1v] [128 x 128] = X [W1p ; W1v]^T    A = X (TMEM),  B = both W1 stacked, N = 128
//   (2) Z2p = H1p W2p^T,  Z2v = H1v W2v^T            A = H1 (TMEM, in place of Z1), N = 64 each
// Both nets' weights stay in shared memory (80 KB, staged by the CTA itself).  Same random stream as ppo_act_kernel (ackb_ppo.cu).
// =================================================================================================================
constexpr uint32_t A_XS = 0;                               // [2] staging buffers, row-major [128][pitch <= 80] floats
constexpr uint32_t A_XBYTES = TM * KP * 4;
constexpr uint32_t A_W1 = A_XS + 2 * A_XBYTES;             // [3 blocks of 128 rows] K-major SW128: rows 0..63 policy, 64..127 value
constexpr uint32_t A_W2 = A_W1 + 3 * 16384;                // [2 nets][2 blocks of 64 rows]
constexpr uint32_t A_F = A_W2 + 2 * W2_BYTES;
constexpr int AF_B1 = 0, AF_B2 = 128, AF_W3 = 256, AF_B3 = 448, AF_LS = 452, AF_PART = 456, AF_END = AF_PART + TM * 12;
constexpr uint32_t A_BAR = A_F + AF_END * 4;               // mbarriers: MMA, 2 x tile landed; TMEM base
constexpr uint32_t A_SMEM = A_BAR + 32 + 1024;
constexpr uint32_t CA_X = 0, CA_Z1 = 128, CA_Z2 = 256;

__device__ __forceinline__ void philox_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(NT5, 1) ppo_act_kernel_tcgen05(ActT5Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* const sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* const F = reinterpret_cast<float*>(sm + A_F);
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(sm + A_BAR);
  uint64_t* const mbx = reinterpret_cast<uint64_t*>(sm + A_BAR + 8);        // [2] tile landed
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(sm + A_BAR + 24);
  const uint32_t sb = smem_u32(sm);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int q4 = warp & 3, part = warp >> 2;
  const int row = 32 * q4 + lane;
  const int D = a.D, PT = a.pitch;
  const Offsets o = offsets(D);
  const float* P = a.params;
  const int ntiles = (a.n + TM - 1) / TM;

  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (t == 32) {
    mbar_init(mbar, 1);
    mbar_init(mbx, 1);
    mbar_init(mbx + 1, 1);
    mbar_fence_init();
  }
  // weights -> K-major SWIZZLE_128B operand images, TF32-rounded.  W1: 128 rows (policy, value) x 96 columns in 3 blocks of 16 KB
  for (int i = t; i < 128 * 96; i += NT5) {
    const int r = i / 96, k = i % 96, net = r >> 6;
    const float w = (k < D) ? tf32r(__ldg(P + (net ? o.W1v : o.W1p) + (r & 63) * D + k)) : 0.0f;
    *reinterpret_cast<float*>(sm + A_W1 + (k >> 5) * 16384 + sw128_off(r, k & 31)) = w;
  }
  for (int i = t; i < 2 * H * H; i += NT5) {
    const int net = i >> 12, r = (i >> 6) & 63, k = i & 63;
    *reinterpret_cast<float*>(sm + A_W2 + net * W2_BYTES + (k >> 5) * 8192 + sw128_off(r, k & 31)) = tf32r(__ldg(P + (net ? o.W2v : o.W2p) + r * H + k));
  }
  for (int i = t; i < 3 * H; i += NT5) F[AF_W3 + i] = (i < 2 * H) ? P[o.Wa + i] : P[o.Wv + (i - 2 * H)];
  if (t < 128) {
    const int net = t >> 6, r = t & 63;
    F[AF_B1 + t] = P[(net ? o.b1v : o.b1p) + r];
    F[AF_B2 + t] = P[(net ? o.b2v : o.b2p) + r];
  }
  if (t < 2) { F[AF_B3 + t] = P[o.ba + t]; F[AF_LS + t] = P[o.ls + t]; }
  if (t == 2) F[AF_B3 + 2] = P[o.bv];
  fence_proxy_async();           // the weight images are read by the tensor core (async proxy)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = *tmem_slot;
  if (tb != 0u) __trap();        // the CTA owns all 512 TMEM columns
  const uint32_t tlane = ((uint32_t)(32 * q4) << 16);
  const float ls0 = F[AF_LS], ls1 = F[AF_LS + 1];
  const float sd0 = expf(ls0), sd1 = expf(ls1);
  const uint32_t lo_w1 = kmajor_lo(sb + A_W1), lo_w2 = kmajor_lo(sb + A_W2);

  // one thread fetches tiles: rows are contiguous, a tile is one bulk copy of ns * pitch floats
  auto fetch = [&](int tile, int buf) {
    const int sbase = tile * TM, ns = min(TM, a.n - sbase);
    const uint32_t bytes = (uint32_t)ns * (uint32_t)PT * 4u;
    mbar_expect_tx(mbx + buf, bytes);
    bulk_g2s(sb + A_XS + buf * A_XBYTES, a.obs + (size_t)sbase * PT, bytes, mbx + buf);
  };
  const bool lead = (warp == 0) && elect_one();
  if (lead && (int)blockIdx.x < ntiles) fetch(blockIdx.x, 0);
  uint32_t phase = 0, xph = 0;       // xph: parity bit per staging buffer
  auto run_mma = [&](auto&& issue) {
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        fence_after_sync();
        issue();
        commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    fence_after_sync();
  };

  int buf = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int sbase = tile * TM, ns = min(TM, a.n - sbase);
    // ---- this thread's part of its own row (columns 20 part .. + 19): staging buffer -> TF32 -> TMEM; rows >= ns and columns >= D are 0.
    // The other staging buffer was read by every thread before the first MMA group of the previous tile: refill it now.
    if (lead && tile + (int)gridDim.x < ntiles) fetch(tile + gridDim.x, buf ^ 1);
    mbar_wait(mbx + buf, (xph >> buf) & 1u);
    xph ^= 1u << buf;
    {
      const float* xr = reinterpret_cast<const float*>(sm + A_XS + buf * A_XBYTES) + row * PT + 20 * part;
      float v[16], w4[4];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < ns) x = *reinterpret_cast<const float4*>(xr + 4 * j);
        const int f = 20 * part + 4 * j;
        const float e0 = (f < D) ? tf32r(x.x) : 0.0f, e1 = (f + 1 < D) ? tf32r(x.y) : 0.0f, e2 = (f + 2 < D) ? tf32r(x.z) : 0.0f,
                    e3 = (f + 3 < D) ? tf32r(x.w) : 0.0f;
        if (j < 4) { v[4 * j] = e0; v[4 * j + 1] = e1; v[4 * j + 2] = e2; v[4 * j + 3] = e3; }
        else { w4[0] = e0; w4[1] = e1; w4[2] = e2; w4[3] = e3; }
      }
      tmem_st16(tlane + CA_X + 20 * part, v);
      tmem_st4(tlane + CA_X + 20 * part + 16, w4);
    }
    // ---- (1) [Z1p | Z1v] = X [W1p ; W1v]^T
    run_mma([&] {
      const uint32_t id = idesc_tf32(128, 128, 0, 0);
#pragma unroll
      for (int k = 0; k < KP / 8; ++k) mma_tf32_ts_lohi(CA_Z1, CA_X + 8 * k, lo_w1 + (k >> 2) * 1024 + (k & 3) * 2, KMAJOR_HI, id, k > 0);
    });
    // ---- H1 = tanh(Z1 + b1) of both nets, back into TMEM in place
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      float v[16], b1[16];
      tmem_ld16(tlane + CA_Z1 + 64 * net + 16 * part, v);
      lds16(&F[AF_B1 + 64 * net + 16 * part], b1);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = tf32r(tanh_fast(v[j] + b1[j]));
      tmem_st16(tlane + CA_Z1 + 64 * net + 16 * part, v);
    }
    // ---- (2) Z2p = H1p W2p^T, Z2v = H1v W2v^T
    run_mma([&] {
      const uint32_t id = idesc_tf32(128, 64, 0, 0);
#pragma unroll
      for (int net = 0; net < 2; ++net)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mma_tf32_ts_lohi(CA_Z2 + 64 * net, CA_Z1 + 64 * net + 8 * k, lo_w2 + net * (W2_BYTES >> 4) + (k >> 2) * 512 + (k & 3) * 2, KMAJOR_HI, id, k > 0);
    });
    // ---- heads: partial dot products of this thread's 16 columns of H2 = tanh(Z2 + b2) with the head weights
    {
      float v[16], b2[16], w0[16], p0 = 0.f, p1 = 0.f, pv = 0.f;
      tmem_ld16(tlane + CA_Z2 + 16 * part, v);
      lds16(&F[AF_B2 + 16 * part], b2);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = tanh_fast(v[j] + b2[j]);
      lds16(&F[AF_W3 + 16 * part], w0);
#pragma unroll
      for (int j = 0; j < 16; ++j) p0 = fmaf(v[j], w0[j], p0);
      lds16(&F[AF_W3 + H + 16 * part], w0);
#pragma unroll
      for (int j = 0; j < 16; ++j) p1 = fmaf(v[j], w0[j], p1);
      tmem_ld16(tlane + CA_Z2 + 64 + 16 * part, v);
      lds16(&F[AF_B2 + 64 + 16 * part], b2);
      lds16(&F[AF_W3 + 2 * H + 16 * part], w0);
#pragma unroll
      for (int j = 0; j < 16; ++j) pv = fmaf(tanh_fast(v[j] + b2[j]), w0[j], pv);
      F[AF_PART + (part * 3) * TM + row] = p0; F[AF_PART + (part * 3 + 1) * TM + row] = p1; F[AF_PART + (part * 3 + 2) * TM + row] = pv;
    }
    __syncthreads();
    if (part == 0 && row < ns) {
      const float* PR = &F[AF_PART + row];
      const float m0 = (PR[0] + PR[3 * TM]) + (PR[6 * TM] + PR[9 * TM]) + F[AF_B3], m1 = (PR[TM] + PR[4 * TM]) + (PR[7 * TM] + PR[10 * TM]) + F[AF_B3 + 1];
      const float vv = (PR[2 * TM] + PR[5 * TM]) + (PR[8 * TM] + PR[11 * TM]) + F[AF_B3 + 2];
      const size_t grow = (size_t)(sbase + row);
      a.value[grow] = vv;
      if (a.mean) { a.mean[grow * 2] = m0; a.mean[grow * 2 + 1] = m1; }
      if (a.action) {
        uint32_t r[4];
        philox_10(a.step, (uint32_t)grow, (uint32_t)(grow >> 32), 0x50504F41u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), r);
        // Box-Muller on two uniforms in (0, 1]
        const float u1 = ((float)(r[0] >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(r[1] >> 8) * (1.0f / 16777216.0f);
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincosf(6.283185307179586f * u2, &sn, &cs);
        const float e0 = rad * cs, e1 = rad * sn;
        a.action[grow * 2] = m0 + sd0 * e0; a.action[grow * 2 + 1] = m1 + sd1 * e1;
        if (a.logp) a.logp[grow] = -0.5f * e0 * e0 - ls0 - 0.9189385332046727f - 0.5f * e1 * e1 - ls1 - 0.9189385332046727f;
      }
    }
    // (the partial sums are overwritten two MMA groups later: no barrier needed here)
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

}  // namespace

#ifdef ACKB_T5_PROFILE
extern "C" int ackb_ppo_t5_profile(long long* out16) {     // 32 counters
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out16, g_t5_prof, sizeof(long long) * 32) != cudaSuccess) return -1;
  long long z[32] = {0};
  cudaMemcpyToSymbol(g_t5_prof, z, sizeof z);
  return 0;
}
#endif

int launch_grad_tcgen05(const PpoArgs& a, cudaStream_t stream, float* adv_out, double* adv_ws) {
  if (a.D >= KP) return ACKB_ERR_ARG;     // the bias gradient of layer 1 rides in column KP - 1 of the observation tile
  if (a.mb <= 0) return ACKB_ERR_ARG;
  static bool attr_done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return ACKB_ERR_NO_DEVICE;
  if (dev < 64 && !attr_done[dev]) {
    if (cudaFuncSetAttribute(ppo_grad_kernel_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) return ACKB_ERR_CUDA;
    attr_done[dev] = true;
  }
  // scratch for the weight images: a stream-ordered allocation, so that concurrent calls on different streams never share it and the
  // call can be captured into a CUDA graph (the allocation becomes a node of the graph; a plain cudaMalloc would invalidate capture)
  static bool pool_done[64] = {false};
  if (dev < 64 && !pool_done[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;                  // keep freed blocks cached: the next call re-uses this one
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_done[dev] = true;
  }
  void* img = nullptr;
  if (cudaMallocAsync(&img, 2 * IMG_BYTES, stream) != cudaSuccess || !img) return ACKB_ERR_CUDA;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (a.mb + TM - 1) / TM;
  int stat_blocks = 0;
  if (adv_out && adv_ws) {
    stat_blocks = (a.mb + 1023) / 1024;                 // >= 4 values per thread
    stat_blocks = stat_blocks > 280 ? 280 : (stat_blocks < 1 ? 1 : stat_blocks);
  }
  ppo_t5_weight_images<<<IMG_BLOCKS + stat_blocks, 256, 0, stream>>>(a.params, a.D, static_cast<unsigned char*>(img), a.grads, a.diag, a.adv, a.idx,
                                                                     a.mb, adv_out, adv_ws, a.diag_keep);
  // programmatic dependent launch: the gradient kernel's CTAs are scheduled while the prologue runs and wait (griddepcontrol.wait)
  // after their set-up and the gather of their first tile
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(ntiles < sms ? ntiles : sms)); cfg.blockDim = dim3(NT5); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, ppo_grad_kernel_tcgen05, a, static_cast<const unsigned char*>(img));
  if (e == cudaSuccess) e = cudaGetLastError();
  if (cudaFreeAsync(img, stream) != cudaSuccess) return ACKB_ERR_CUDA;
  return e == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

int launch_act_tcgen05(const ActT5Args& a, cudaStream_t stream) {
  if (a.D <= 0 || a.D > KP || a.n <= 0 || a.pitch < a.D || a.pitch > KP || (a.pitch & 3) || (reinterpret_cast<uintptr_t>(a.obs) & 15)) return ACKB_ERR_ARG;
  static bool attr_done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return ACKB_ERR_NO_DEVICE;
  if (dev < 64 && !attr_done[dev]) {
    if (cudaFuncSetAttribute(ppo_act_kernel_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A_SMEM) != cudaSuccess) return ACKB_ERR_CUDA;
    attr_done[dev] = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (a.n + TM - 1) / TM;
  ppo_act_kernel_tcgen05<<<ntiles < sms ? ntiles : sms, NT5, A_SMEM, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

}  // namespace ackb_ppo
