// Fused front of cascade blocks 3 and 4: conv 7x7 s1 (2 -> C1) + LeakyReLU  ->  conv 5x5 s2 (C1 -> C2) + LeakyReLU
// in ONE persistent tcgen05 kernel, run as CTA pairs (tcgen05.mma.cta_group::2).  The first conv's output (1.15 MB /
// 0.57 MB per pair, the largest activations of the network) never exists in HBM: its tiles go TMEM -> registers ->
// shared memory, laid out exactly as the second conv's A operand, and the second conv's MMAs read them from there.
//
// Tile = 7 x 14 patch of conv-2 pixel groups (one group = 64/C2 output pixels = G1 = 64/C1 conv-1 pixels = one
// 128-byte K row).  Per tile and CTA:
//   1. TMA: one 4-D box {32 el, 8 groups, 38 rows} of the 2-channel block input (overlapping x windows, halo 5;
//      64-byte rows, SWIZZLE_64B), double-buffered and fetched a tile ahead.
//   2. conv 1 as ONE 128-row MMA tile with N = 128: GEMM row (t, g) = conv-1 row PAIR 2t, 2t+1 of group g — Toeplitz
//      expansion in y as well as in x.  Window row j = 0..7 of the pair is the same input plane shifted by j rows with the
//      8-row atoms of the A descriptor two input rows apart (SBO = 1024 B), so A still comes straight from the TMA box;
//      B1[j] holds W1[ky = j - r] for output row r = 0, 1 in columns r*64 .. r*64+63.  16 MMAs of N = 128 replace the 28
//      of N = 64 of the one-row formulation: at N = 64 the tensor core waits on its shared-memory operand fetch (4 KB of A
//      per 32 cycles of math, ncu: l1tex__data_pipe_tc_wavefronts_mem_shared 89 %), at N = 128 the same A bytes feed 64
//      cycles of math.  The accumulators were pre-loaded with the bias (tcgen05.st), every MMA accumulates.
//   3. conv-1 epilogue warps (2..9): D1 -> bf16 -> LeakyReLU on the packed pair, zero outside the image (conv 2's zero
//      padding) -> shared-memory planes [row parity r][t][group], 128B-swizzled by ADDRESS (the UMMA swizzle is purely
//      address-based — tools/umma_offset_test.cu — so operands may start at any 128-byte row): accumulator columns
//      r*64 .. r*64+63 ARE plane r; then they re-arm D1 with the bias.
//   4. conv 2: tap ky = rho + 2a, chunk c reads plane rho at row offset (a*8 + c): "chunk 1 of group w" is "chunk 0
//      of group w+1", so nothing is duplicated.  B2 and B1 stay resident in shared memory, half of the N rows per CTA.
//   5. conv-2 epilogue warps (10..17): D2 -> bf16 -> LeakyReLU -> each thread stores its 64 contiguous bytes of the
//      (haloed NHWC) conv-2 output.
// The MMA warp issues conv 1 of tile k+1 AHEAD of conv 2 of tile k (D1 double-buffered in TMEM), so the tensor pipe
// works through the D1 -> planes hand-off instead of idling on it.  Only the leader CTA of a pair issues MMAs (M = 256);
// its hand-off barriers count arrivals from both CTAs.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cudaTypedefs.h>

#include "conv_bf16.h"
#include "tc_ptx.cuh"

namespace uahn {
namespace {

constexpr int FF_EPI1_WARPS = 8;                     // conv-1 epilogue: 2 per TMEM lane quadrant, one row parity (64 columns) each
constexpr int FF_EPI2_WARPS = 8;                     // conv-2 epilogue: 2 per quadrant, 32 columns each
constexpr int FF_THREADS = 64 + 32 * (FF_EPI1_WARPS + FF_EPI2_WARPS);
constexpr int IN_ROWS = 38;                          // 32 conv-1 rows + 6 (7x7 halo)
// input plane: per (row, group) the first 32 elements of the group's x window (14 / 10 pixels x 2 channels are used):
// 64-byte rows, SWIZZLE_64B — half the shared-memory writes and L2 reads of a 128-byte-row plane
constexpr int IN_ROW_BYTES = 8 * 64;                 // one image row of the plane: 8 groups x 64 B = one 8-row swizzle atom
constexpr int IN_BYTES = IN_ROWS * IN_ROW_BYTES;     // 19 456
constexpr int B1_STAGES = 4, B2_STAGES = 10;         // 8 window rows packed in pairs; 5 taps x 2 chunks
constexpr int B1_STAGE = 128 * 128, B2_STAGE = 64 * 128;   // full stages: N = 128 / 64 rows x 128 B
constexpr int BST1 = B1_STAGE / 2, BST2 = B2_STAGE / 2;    // this CTA's half of the N rows
// conv-2 operand planes: rows t = 0..15 are written.  Taps a = 1, 2 and chunk 1 make the DISCARDED M rows (rr2 >= 14 or
// w2l = 7) read up to 17 group-rows past a plane: that lands in the next plane or, after the last one, in the resident B1
// operand placed right behind the planes — finite bf16 either way, and every GEMM row is independent.
constexpr int PLANE_ROWS = 16;
constexpr int PLANE_BYTES = PLANE_ROWS * 8 * 128;    // 16 384
constexpr int SMEM_BYTES = 1024 + 2 * IN_BYTES + B1_STAGES * BST1 + B2_STAGES * BST2 + 2 * 2 * PLANE_BYTES + 2 * 64 * 4 + 32 * 8;
static_assert(SMEM_BYTES <= 232448, "fused front: shared memory");

struct FusedParams {
  const uint8_t* b1_image;
  const uint8_t* b2_image;
  const float* bias1_x;     // [64]: b1[n % C1]
  const float* bias2_x;     // [64]: b2[n % C2]
  uint8_t* out;             // conv-2 output (haloed NHWC bf16)
  int n_img, TX, TY;
  int H1, W1;               // conv-1 output size (= block input size)
  int Wox2;                 // conv-2 groups per row
  long long out_pitch_n_b;
  int out_pitch_y_b;
  long long out_origin_b;
  unsigned long long magic_tiles, magic_tx;
};

// C1: conv-1 output channels (8 or 16); KS2B: k-steps of conv-2 chunk 1 (2 for block 4, 3 for block 3).
// Launched as clusters of two CTAs that share every UMMA (tcgen05.mma.cta_group::2, M = 256): each CTA keeps its own tile
// stream, input planes, D1 -> planes epilogue and output, but holds only HALF of the resident B operands (N rows
// [rank * N/2, (rank + 1) * N/2)).  Only the leader (cluster rank 0) issues MMAs.
template <int C1, int KS2B>
__global__ void __launch_bounds__(FF_THREADS, 1) conv_fused_front_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                          const __grid_constant__ FusedParams p) {
  constexpr int G1 = 64 / C1;                        // conv-1 pixels per group
  const uint32_t crank = cluster_ctarank();
  const bool cta_leader = crank == 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sIn = smem;                                             // conv-1 input plane, double-buffered
  uint8_t* sPl = sIn + 2 * IN_BYTES;                               // [2 buffers][2 parities][PLANE_BYTES]
  uint8_t* sB1 = sPl + 2 * 2 * PLANE_BYTES;                        // (behind the planes: see PLANE_ROWS)
  uint8_t* sB2 = sB1 + B1_STAGES * BST1;
  float* sBias = reinterpret_cast<float*>(sB2 + B2_STAGES * BST2);  // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_img = p.TX * p.TY;
  const int total_tiles = p.n_img * tiles_per_img;

  // bars: 0-1 in_full[buf], 2 bres, 3-4 d1_full[buf], 5-6 d1_empty[buf], 7 d2_full, 8 d2_empty, 9-10 in_empty[buf],
  //       11-12 planes_full[pb], 13-14 planes_empty[pb]
  constexpr int B_IN_FULL = 0, B_RES = 2, B_D1_FULL = 3, B_D1_EMPTY = 5, B_D2_FULL = 7, B_D2_EMPTY = 8, B_IN_EMPTY = 9,
                B_PL_FULL = 11, B_PL_EMPTY = 13;
  constexpr uint32_t TMEM_COLS = 512;                    // D1: 2 buffers x 128 columns; D2: 64 columns at 256
  constexpr uint32_t D2_COL = 256;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_IN_FULL + i), 1); mbar_init(BAR(B_IN_EMPTY + i), 1); }
      mbar_init(BAR(B_RES), 1);
      for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_D1_FULL + i), 1); mbar_init(BAR(B_D1_EMPTY + i), 2 * FF_EPI1_WARPS); }
      for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_PL_FULL + i), 2 * FF_EPI1_WARPS); mbar_init(BAR(B_PL_EMPTY + i), 1); }
      mbar_init(BAR(B_D2_FULL), 1); mbar_init(BAR(B_D2_EMPTY), 2 * FF_EPI2_WARPS);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 64; i += FF_THREADS) { sBias[i] = p.bias1_x[i]; sBias[64 + i] = p.bias2_x[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Epilogue warps: TMEM lane quadrant q = warp % 4 (hardware rule).  Warps 2..9 run the conv-1 epilogue (quadrant q, row
  // parity `half`: 64 accumulator columns), warps 10..17 the conv-2 epilogue (32 columns each), so the D1 -> planes
  // hand-off that conv 2 waits for never queues behind global stores.  The accumulators are pre-loaded with the bias
  // (tcgen05.st): every MMA accumulates, the epilogues do no bias add, and each drain re-arms its columns.
  const int q = warp & 3, half = ((warp - 2) >> 2) & 1;
  const bool epi2_warp = warp >= 2 + FF_EPI1_WARPS;
  const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t bias_addr = smem_u32(sBias);
  // this thread's lane of D1[buf] columns [half*64, half*64 + 64) <- bias1 (the same 64 values for both row parities)
  auto arm_d1 = [&](int buf) {
#pragma unroll
    for (int c16 = 0; c16 < 4; ++c16) {
      uint32_t b[16];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint4 x = ld_shared_v4(bias_addr + (uint32_t)(c16 * 64 + v * 16));
        b[4 * v] = x.x; b[4 * v + 1] = x.y; b[4 * v + 2] = x.z; b[4 * v + 3] = x.w;
      }
      tmem_st16(t_lane + (uint32_t)(buf * 128 + half * 64 + c16 * 16), b);
    }
  };
  uint32_t biasu[32];                                     // conv-2 epilogue: this warp's 32 columns of bias2
  if (warp >= 2) {
    if (!epi2_warp) {
      arm_d1(0);
      arm_d1(1);
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) biasu[c] = __float_as_uint(sBias[64 + half * 32 + c]);
      tmem_st16(t_lane + D2_COL + (uint32_t)(half * 32), biasu);
      tmem_st16(t_lane + D2_COL + (uint32_t)(half * 32 + 16), biasu + 16);
    }
    tmem_st_wait();
  }
  if (warp == 0 && lane == 0) {
    // resident B operands: this CTA's half of the N rows of every stage
    tma_prefetch_desc(&tmap);
    mbar_arrive_expect_tx(BAR(B_RES), (uint32_t)(B1_STAGES * BST1 + B2_STAGES * BST2));
    for (int s = 0; s < B1_STAGES; ++s)
      bulk_g2s(smem_u32(sB1 + s * BST1), p.b1_image + (size_t)s * B1_STAGE + (size_t)crank * BST1, BST1, BAR(B_RES));
    for (int s = 0; s < B2_STAGES; ++s)
      bulk_g2s(smem_u32(sB2 + s * BST2), p.b2_image + (size_t)s * B2_STAGE + (size_t)crank * BST2, BST2, BAR(B_RES));
    mbar_wait(BAR(B_RES), 0);                  // the leader's MMAs read the peer's half: resident before the cluster sync
  }
  __syncwarp();                                // barrier.cluster is .aligned: every warp arrives converged
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();                 // everything above touched only shared memory, TMEM and weights
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int k = 0; k < my_tiles; ++k) {
        const int tile = blockIdx.x + k * gridDim.x;
        const int img = fast_div(tile, p.magic_tiles), rem = tile - img * tiles_per_img;
        const int ty = fast_div(rem, p.magic_tx), tx = rem - ty * p.TX;
        const int ib = k & 1;
        mbar_wait(BAR(B_IN_EMPTY + ib), ((k >> 1) & 1) ^ 1);  // conv-1 MMAs of tile k-2 have read this buffer
        // both CTAs' planes complete on the leader's barrier, which the leader arms for both
        if (cta_leader) mbar_arrive_expect_tx(BAR(B_IN_FULL + ib), 2 * IN_BYTES);
        tma_load_4d_pair(smem_u32(sIn + ib * IN_BYTES), &tmap, 0, 7 * tx, 28 * ty, img, BAR(B_IN_FULL + ib));
      }
    }
  } else if (warp == 1 && cta_leader) {
    // ===================== MMA issuer (the leader CTA of a pair issues for both) =====================
    // Issue order: conv1(0), then per tile k: conv1(k+1), conv2(k).  conv 2 of tile k has to wait for the conv-1
    // epilogue of tile k (TMEM -> registers -> planes); with conv 1 of the NEXT tile queued in front of it the tensor
    // pipe works through that wait instead of idling (D1 is double-buffered in TMEM for this).
    constexpr uint32_t idesc1 = umma_idesc_bf16(256, 128), idesc2 = umma_idesc_bf16(256, 64);
    // K-major SW128 descriptors: LBO = 1 (unused), version 1, layout 2; SBO = bytes between 8-row atoms
    constexpr uint64_t DESC_BASE = (1ull << 16) | (1ull << 46) | (2ull << 61);
    constexpr uint64_t DESC_SBO1K = DESC_BASE | ((uint64_t)(1024 >> 4) << 32);
    // conv-1 A: K-major SWIZZLE_64B (layout 4), 8-row atoms (one image row of the plane, 512 B) two image rows apart:
    // GEMM row (t, g) of window row j reads image row 2t + j
    constexpr uint64_t DESC_A1 = (1ull << 16) | (1ull << 46) | (4ull << 61) | ((uint64_t)((2 * IN_ROW_BYTES) >> 4) << 32);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t in16 = (smem_u32(sIn) & 0x3FFFFu) >> 4, b1_16 = (smem_u32(sB1) & 0x3FFFFu) >> 4;
    const uint32_t b2_16 = (smem_u32(sB2) & 0x3FFFFu) >> 4, pl16 = (smem_u32(sPl) & 0x3FFFFu) >> 4;
    const bool leader = elect_one();
    auto conv1 = [&](int k) {
      const int buf = k & 1;
      mbar_wait(BAR(B_IN_FULL + buf), (k >> 1) & 1);
      mbar_wait(BAR(B_D1_EMPTY + buf), ((k >> 1) & 1) ^ 1);          // the epilogue has drained and re-armed D1 of tile k-2
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int j = 0; j < 8; ++j)                                   // window row j of the conv-1 row pair
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t alo = in16 + (uint32_t)(buf * (IN_BYTES / 16) + j * (IN_ROW_BYTES / 16) + kk * 2);
            const uint32_t blo = b1_16 + (uint32_t)((j >> 1) * (BST1 / 16) + (j & 1) * 4 + kk * 2);
            tc_mma_bf16_pair(tmem_u + (uint32_t)(buf * 128), DESC_A1 | (uint64_t)alo, DESC_SBO1K | (uint64_t)blo, idesc1, 1u);
          }
        tc_commit_pair(BAR(B_D1_FULL + buf));
        tc_commit_pair(BAR(B_IN_EMPTY + buf));                        // this input buffer may be refilled
      }
      __syncwarp();
    };
    if (my_tiles > 0) conv1(0);
    for (int k = 0; k < my_tiles; ++k) {
      if (k + 1 < my_tiles) conv1(k + 1);
      // ---- conv 2: 5 taps x (4 + KS2B) k-steps on the planes written by the conv-1 epilogue ----
      const int pb = k & 1;
      mbar_wait(BAR(B_PL_FULL + pb), (k >> 1) & 1);
      mbar_wait(BAR(B_D2_EMPTY), (k & 1) ^ 1);                        // D2 of the previous tile has been read
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
          const int rho = ky & 1, a = ky >> 1;
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int kk = 0; kk < (c == 0 ? 4 : KS2B); ++kk) {
              const uint32_t alo = pl16 + (uint32_t)((pb * 2 + rho) * (PLANE_BYTES / 16) + (a * 8 + c) * (128 / 16) + kk * 2);
              const uint32_t blo = b2_16 + (uint32_t)((ky * 2 + c) * (BST2 / 16) + kk * 2);
              tc_mma_bf16_pair(tmem_u + D2_COL, DESC_SBO1K | (uint64_t)alo, DESC_SBO1K | (uint64_t)blo, idesc2, 1u);
            }
        }
        tc_commit_pair(BAR(B_D2_FULL));
        tc_commit_pair(BAR(B_PL_EMPTY + pb));                         // these planes may be rewritten
      }
      __syncwarp();
    }
    tc_fence_before();
  } else if (warp >= 2) {
    // ===================== epilogue warps (2..17) =====================
    // arrivals on the barriers the MMA issuer waits on go to the leader CTA of the pair
    const int m = q * 32 + lane;                              // accumulator row
    // tile coordinates advance incrementally (no divisions in the loop)
    const int step_img = (int)gridDim.x / tiles_per_img, step_rem = (int)gridDim.x - step_img * tiles_per_img;
    int img = (int)blockIdx.x / tiles_per_img, rem = (int)blockIdx.x - img * tiles_per_img;
    const uint32_t mtx = (uint32_t)((65536 + p.TX - 1) / p.TX);                                     // rem < 65536 / TX
    if (!epi2_warp) {
      // ---- conv-1 epilogue: D1[buf] columns [half*64, +64) -> LeakyReLU -> bf16 plane `half` (conv 2's A operand) ----
      const int t = m >> 3, g = m & 7;                        // conv-1 row pair of the tile, group
      // plane row t*8 + g of plane `half`; 16-byte chunk cc lands at ((cc ^ g) << 4): address swizzle with (row & 7) = g
      const uint32_t prow = smem_u32(sPl) + (uint32_t)(half * PLANE_BYTES + (t * 8 + g) * 128);
      for (int k = 0; k < my_tiles; ++k) {
        const int ty = (int)(((uint32_t)rem * mtx) >> 16), tx = rem - ty * p.TX;
        const int buf = k & 1, pb = k & 1;
        const int y0 = 28 * ty - 2, x0 = G1 * 7 * tx - 2;     // first conv-1 row / pixel of the tile region
        // conv 2's zero padding: conv-1 outputs outside the image must be stored as zeros (border tiles only)
        const bool border = y0 < 0 || y0 + 31 >= p.H1 || x0 < 0 || x0 + G1 * 8 - 1 >= p.W1;
        const bool yok = (unsigned)(y0 + 2 * t + half) < (unsigned)p.H1;
        mbar_wait(BAR(B_PL_EMPTY + pb), ((k >> 1) & 1) ^ 1);   // conv 2 of tile k - 2 has read these planes
        mbar_wait(BAR(B_D1_FULL + buf), (k >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tmem_ld32(t_lane + (uint32_t)(buf * 128 + half * 64 + hh * 32), r);
          tmem_ld_wait();
          uint32_t packed[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) packed[e] = pack_lrelu_bf16x2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
          if (border) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {                     // 16-byte chunk 4*hh + c = columns hh*32 + 8c .. +7: one pixel
              const int x = x0 + G1 * g + (hh * 32 + 8 * c) / C1;
              const bool ok = yok && (unsigned)x < (unsigned)p.W1;
#pragma unroll
              for (int e = 0; e < 4; ++e) packed[4 * c + e] = ok ? packed[4 * c + e] : 0u;
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st_shared_v4(prow + (uint32_t)(pb * 2 * PLANE_BYTES) + (uint32_t)(((4 * hh + c) ^ g) << 4), packed[4 * c],
                         packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
        }
        fence_proxy_async();                                  // generic-proxy writes -> visible to the UMMA reads
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(BAR(B_PL_FULL + pb));
        // off the critical path: re-arm D1 with the bias and hand it back to the MMA warp
        arm_d1(buf);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(BAR(B_D1_EMPTY + buf));
        rem += step_rem; img += step_img;
        if (rem >= tiles_per_img) { rem -= tiles_per_img; ++img; }
      }
    } else {
      // ---- conv-2 epilogue: D2 -> LeakyReLU -> bf16 -> global; each thread stores its 64 contiguous bytes ----
      const int rr2 = m >> 3, w2l = m & 7;
      const bool row_ok = rr2 < 14 && w2l < 7;
      const long long thr_off = p.out_origin_b + (long long)rr2 * p.out_pitch_y_b + (long long)w2l * 128 + half * 64;
      const uint32_t t_d2 = t_lane + D2_COL + (uint32_t)(half * 32);
      for (int k = 0; k < my_tiles; ++k) {
        const int ty = (int)(((uint32_t)rem * mtx) >> 16), tx = rem - ty * p.TX;
        mbar_wait(BAR(B_D2_FULL), k & 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(t_d2, r);
        tmem_ld_wait();
        tmem_st16(t_d2, biasu);
        tmem_st16(t_d2 + 16u, biasu + 16);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(BAR(B_D2_EMPTY));
        uint32_t packed[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) packed[e] = pack_lrelu_bf16x2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
        if (row_ok && 7 * tx + w2l < p.Wox2) {
          uint8_t* o = p.out + (thr_off + (long long)img * p.out_pitch_n_b + (long long)(14 * ty) * p.out_pitch_y_b +
                                (long long)(7 * tx) * 128);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st_global_v4(o + 16 * c, packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
        }
        rem += step_rem; img += step_img;
        if (rem >= tiles_per_img) { rem -= tiles_per_img; ++img; }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();   // the peer's shared memory and TMEM are in use until the leader is done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_ff() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

bool upload(const void* src, size_t bytes, void** dst, std::vector<void*>& allocs) {
  if (cudaMalloc(dst, bytes) != cudaSuccess) return false;
  allocs.push_back(*dst);
  return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
}

inline size_t sw128_byte(int stage, int rows, int n, int kk) {   // [stage][rows][128 B], K-major, 128B swizzle
  return ((size_t)stage * rows + n) * 128 + (size_t)((((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2);
}

}  // namespace

// g0 / g1: geometry of the two convs (g0: 7x7 s1 Cin 2; g1: 5x5 s2); x: the block input tensor with halo 5.
int conv_fused_prepare(FusedPlan& plan, const std::vector<float>& wk0, const std::vector<float>& bias0,
                       const std::vector<float>& wk1, const std::vector<float>& bias1, const ConvGeom& g0,
                       const ConvGeom& g1, const Tensor& x, std::vector<void*>& allocs, std::string& err) {
  plan.enabled = 0;
  const int C1 = g0.Cout, C2 = g1.Cout;
  if (!(g0.KH == 7 && g0.stride == 1 && g0.Cin == 2 && g1.KH == 5 && g1.stride == 2 && g1.Cin == C1)) return 0;
  if (!((C1 == 8 && C2 == 16) || (C1 == 16 && C2 == 32))) return 0;
  if (x.ph != 5 || x.pwl != 5 || g1.Ho % 14 || g1.Wo % (64 / C2)) return 0;
  const int G1 = 64 / C1, xb2 = 64 / C2;
  plan.C1 = C1;
  plan.TX = (g1.Wo / xb2 + 6) / 7;
  plan.TY = g1.Ho / 14;
  plan.Wox2 = g1.Wo / xb2;
  plan.H1 = g0.Ho; plan.W1 = g0.Wo;
  // ---- tensor map over the block input: {32 el window, conv-1 group, padded row, image} ----
  PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode_ff();
  if (!encode) { err = "cuTensorMapEncodeTiled entry point not found"; return -2; }
  const cuuint64_t gdim[4] = {32, 48, (cuuint64_t)x.Hp, (cuuint64_t)x.N};
  const cuuint64_t gstr[3] = {(cuuint64_t)G1 * 2 * 2, (cuuint64_t)x.pitch_y() * 2, (cuuint64_t)x.pitch_n * 2};
  const cuuint32_t box[4] = {32, 8, IN_ROWS, 1}, estr[4] = {1, 1, 1, 1};
  if (encode(reinterpret_cast<CUtensorMap*>(plan.tmap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x.p, gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 0;
  // ---- B1: window rows j = 0..7 of a conv-1 row PAIR, two per 64-element stage; N = 128:
  //      element (n = r*64 + i*C1 + co, k = (j&1)*32 + xi*2 + c) of stage j>>1 = W1[co][c][ky = j - r][kx = xi - i]
  std::vector<uint16_t> b1((size_t)B1_STAGES * 128 * 64, 0);
  for (int j = 0; j < 8; ++j)
    for (int r = 0; r < 2; ++r) {
      const int ky = j - r;
      if (ky < 0 || ky >= 7) continue;
      for (int xi = 0; xi < G1 + 6; ++xi)
        for (int c = 0; c < 2; ++c)
          for (int i = 0; i < G1; ++i) {
            const int kx = xi - i;
            if (kx < 0 || kx >= 7) continue;
            for (int co = 0; co < C1; ++co) {
              const float w = wk0[(size_t)((ky * 7 + kx) * 2 + c) * C1 + co];
              b1[sw128_byte(j >> 1, 128, r * 64 + i * C1 + co, (j & 1) * 32 + xi * 2 + c) / 2] = f32_to_bf16_host(w);
            }
          }
    }
  // ---- B2: stage (ky, chunk); element (n = xo*C2 + co, k = q - 64*chunk), q = xi*C1 + c1, W2[co][c1][ky][xi - 2*xo]
  std::vector<uint16_t> b2((size_t)B2_STAGES * 64 * 64, 0);
  const int run2 = (2 * (xb2 - 1) + 5) * C1;
  for (int ky = 0; ky < 5; ++ky)
    for (int qel = 0; qel < run2; ++qel) {
      const int xi = qel / C1, c1 = qel % C1;
      for (int xo = 0; xo < xb2; ++xo) {
        const int kx = xi - 2 * xo;
        if (kx < 0 || kx >= 5) continue;
        for (int co = 0; co < C2; ++co) {
          const float w = wk1[(size_t)((ky * 5 + kx) * C1 + c1) * C2 + co];
          b2[sw128_byte(ky * 2 + qel / 64, 64, xo * C2 + co, qel % 64) / 2] = f32_to_bf16_host(w);
        }
      }
    }
  std::vector<float> bx1(64), bx2(64);
  for (int n = 0; n < 64; ++n) { bx1[n] = bias0[n % C1]; bx2[n] = bias1[n % C2]; }
  if (!upload(b1.data(), b1.size() * 2, &plan.b1_image, allocs) || !upload(b2.data(), b2.size() * 2, &plan.b2_image, allocs) ||
      !upload(bx1.data(), 256, (void**)&plan.bias1_x, allocs) || !upload(bx2.data(), 256, (void**)&plan.bias2_x, allocs)) {
    err = "cudaMalloc/cudaMemcpy (fused conv operands)";
    return -2;
  }
  plan.enabled = 1;
  if (getenv("UAHN_DEBUG"))
    fprintf(stderr, "[uahn] fused front: 7x7 2->%d + 5x5s2 ->%d, tiles %dx%d per image, smem %d B\n", C1, C2, plan.TX,
            plan.TY, SMEM_BYTES);
  return 0;
}

cudaError_t launch_conv_fused(const FusedPlan& plan, void* out, const ConvGeom& g1, int n_img, int num_sms,
                              cudaStream_t st) {
  FusedParams p{};
  p.b1_image = (const uint8_t*)plan.b1_image;
  p.b2_image = (const uint8_t*)plan.b2_image;
  p.bias1_x = plan.bias1_x;
  p.bias2_x = plan.bias2_x;
  p.out = (uint8_t*)out;
  p.n_img = n_img; p.TX = plan.TX; p.TY = plan.TY;
  p.H1 = plan.H1; p.W1 = plan.W1; p.Wox2 = plan.Wox2;
  p.out_pitch_n_b = g1.out_pitch_n * 2;
  p.out_pitch_y_b = (int)(g1.out_pitch_y * 2);
  p.out_origin_b = g1.out_origin * 2;
  p.magic_tiles = ((1ull << 40) + p.TX * p.TY - 1) / (p.TX * p.TY);
  p.magic_tx = ((1ull << 40) + p.TX - 1) / p.TX;
  const int tiles = n_img * p.TX * p.TY;      // tiles per image are 48 / 24: always even, so CTA pairs always divide the work
  const int grid = std::min(tiles, num_sms) & ~1;
  if (grid < 2) return cudaErrorInvalidConfiguration;
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(plan.tmap);
  // per kernel and per device (common.cuh); NOT a static inside the generic lambda: both instantiations have the same
  // function-pointer type and would share it
  static SmemOptIn optins[2];
  auto launch = [&](auto kern, SmemOptIn& optin) -> cudaError_t {
    if (cudaError_t e = optin.ensure(kern, SMEM_BYTES); e != cudaSuccess) {
      if (getenv("UAHN_DEBUG")) fprintf(stderr, "[uahn] fused front: shared-memory opt-in of %d B failed: %s\n", SMEM_BYTES, cudaGetErrorString(e));
      return e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(FF_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, *tm, p);
    if (le != cudaSuccess && getenv("UAHN_DEBUG"))
      fprintf(stderr, "[uahn] fused front launch (grid %d, %d threads, %d B): %s\n", grid, FF_THREADS, SMEM_BYTES, cudaGetErrorString(le));
    return le;
  };
  return plan.C1 == 8 ? launch(conv_fused_front_kernel<8, 2>, optins[0]) : launch(conv_fused_front_kernel<16, 3>, optins[1]);
}

}  // namespace uahn
