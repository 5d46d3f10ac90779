// Fused front of cascade blocks 3 and 4: conv 7x7 s1 (2 -> C1) + LeakyReLU  ->  conv 5x5 s2 (C1 -> C2) + LeakyReLU
// in ONE persistent tcgen05 kernel.  The first conv's output (1.15 MB / 0.57 MB per pair, the largest activations of
// the network) never exists in HBM: its tiles go TMEM -> registers -> shared memory, laid out exactly as the second
// conv's A operand, and the second conv's MMAs read them from there.  This removes 31 % (block 4) + 15 % (block 3) of
// the conv DRAM traffic, which is what bounds the conv stacks.
//
// Tile = 7 x 14 patch of conv-2 pixel groups (one group = 64/C2 output pixels = G1 = 64/C1 conv-1 pixels = one
// 128-byte K row).  Per tile:
//   1. TMA: one 4-D box {64 el, 8 groups, 38 rows} of the 2-channel block input (overlapping x windows, halo 5),
//      double-buffered and fetched a tile ahead.
//   2. conv 1 as two 128-row MMA tiles (rows 0-15 / 16-31 of the 32 x 8-group region conv 2 needs); kernel row ky is
//      the same plane shifted by ky rows (shifted-window trick of conv_bf16_tma.cu); weights packed two taps per
//      64-element B stage.  The accumulators were pre-loaded with the bias (tcgen05.st), every MMA accumulates.
//   3. conv-1 epilogue warps (2..9): D1 -> bf16 -> LeakyReLU on the packed pair, zero outside the image (conv 2's zero
//      padding) -> shared-memory planes [row parity][t = row/2][group], 128B-swizzled by ADDRESS (the UMMA swizzle is
//      purely address-based — tools/umma_offset_test.cu — so operands may start at any 128-byte row); then they re-arm
//      D1 with the bias.
//   4. conv 2: tap ky = rho + 2a, chunk c reads plane rho at row offset (a*8 + c): "chunk 1 of group w" is "chunk 0
//      of group w+1", so nothing is duplicated.  B2 and B1 stay resident in shared memory.
//   5. conv-2 epilogue warps (10..17): D2 -> bf16 -> LeakyReLU -> each thread stores its 64 contiguous bytes of the
//      (haloed NHWC) conv-2 output.
// The MMA warp issues conv 1 of tile k+1 AHEAD of conv 2 of tile k (D1 double-buffered in TMEM), so the tensor pipe
// works through the D1 -> planes hand-off instead of idling on it.  PAIR: clusters of two CTAs share every UMMA
// (tcgen05.mma.cta_group::2, M = 256), each holding half of B; the freed shared memory double-buffers the planes.
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

constexpr int FF_EPI1_WARPS = 8;                     // conv-1 epilogue: 2 per TMEM lane quadrant, 32 columns each
constexpr int FF_EPI2_WARPS = 8;                     // conv-2 epilogue: 2 per quadrant, 32 columns each
constexpr int FF_THREADS = 64 + 32 * (FF_EPI1_WARPS + FF_EPI2_WARPS);
constexpr int IN_ROWS = 38;                          // 32 conv-1 rows + 6 (7x7 halo)
constexpr int IN_BYTES = IN_ROWS * 8 * 128;          // 38 912
constexpr int B1_STAGES = 4, B2_STAGES = 10;         // 7 taps packed in pairs; 5 taps x 2 chunks
constexpr int BSTAGE = 64 * 128;                     // N = 64 rows x 128 B
constexpr int PLANE_ROWS = 18;                       // t = 0..15 written, +2 rows read only by dummy M rows
constexpr int PLANE_BYTES = PLANE_ROWS * 8 * 128;    // 18 432
constexpr int smem_bytes(bool pair) {   // pair: half of B per CTA, the conv-2 operand planes double-buffered
  return 1024 + 2 * IN_BYTES + (B1_STAGES + B2_STAGES) * (pair ? BSTAGE / 2 : BSTAGE) + (pair ? 2 : 1) * 2 * PLANE_BYTES +
         2 * 64 * 4 + 32 * 8;
}
constexpr int SMEM_BYTES = smem_bytes(false) > smem_bytes(true) ? smem_bytes(false) : smem_bytes(true);

#ifndef UAHN_FF_PROFILE
#define UAHN_FF_PROFILE 0
#endif
__device__ __forceinline__ long long ff_clock() { return UAHN_FF_PROFILE ? clock64() : 0ll; }

// event trace of a few tiles of CTAs 0/1 (UAHN_FF_PROFILE only): TR(role, k, event)
#define FF_TR(role, k, ev)                                                                                 \
  do {                                                                                                     \
    if (UAHN_FF_PROFILE && p.dbg && blockIdx.x < 2 && (k) >= 40 && (k) < 48)                                \
      p.dbg[24 * 1024 + ((role) * 8 + ((k) - 40)) * 8 + (ev)] = (unsigned long long)clock64();             \
  } while (0)

struct FusedParams {
  unsigned long long* dbg;   // optional [grid][24] cycle counters (-DUAHN_FF_PROFILE=1 + UAHN_FF_DEBUG)
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
// PAIR: launched as clusters of two CTAs that share every UMMA (tcgen05.mma.cta_group::2, M = 256): each CTA keeps its
// own tile stream, input planes, D1 -> planes epilogue and output, but holds only HALF of the resident B operands (32 of
// the 64 N rows) — the tensor core's operand fetch from shared memory, which bounds this kernel, drops from 6 KB to 5 KB
// per CTA and MMA.  Only the leader (cluster rank 0) issues MMAs; its hand-off barriers count arrivals from both CTAs.
template <int C1, int KS2B, bool PAIR>
__global__ void __launch_bounds__(FF_THREADS, 1) conv_fused_front_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                          const __grid_constant__ FusedParams p) {
  constexpr int G1 = 64 / C1;                        // conv-1 pixels per group
  constexpr int BST = PAIR ? BSTAGE / 2 : BSTAGE;    // bytes of one resident B stage in THIS CTA
  constexpr int NCTA = PAIR ? 2 : 1;
  constexpr int NPL = PAIR ? 2 : 1;                  // buffers of the conv-2 operand planes (the pair has the room)
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const bool cta_leader = crank == 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sIn = smem;                                             // conv-1 input plane, double-buffered
  uint8_t* sB1 = sIn + 2 * IN_BYTES;
  uint8_t* sB2 = sB1 + B1_STAGES * BST;
  uint8_t* sPl = sB2 + B2_STAGES * BST;                            // [2 parities][PLANE_BYTES]
  float* sBias = reinterpret_cast<float*>(sPl + NPL * 2 * PLANE_BYTES);  // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_img = p.TX * p.TY;
  const int total_tiles = p.n_img * tiles_per_img;

  // bars: 0-1 in_full[buf], 2 bres, 3-6 d1_full[buf][jt], 7-10 d1_empty[buf][jt], 11 d2_full, 12 d2_empty,
  //       13-14 in_empty[buf], 15-16 planes_full[pb], 17-18 planes_empty[pb]
  constexpr int B_IN_FULL = 0, B_RES = 2, B_D1_FULL = 3, B_D1_EMPTY = 7, B_D2_FULL = 11, B_D2_EMPTY = 12, B_IN_EMPTY = 13,
                B_PL_FULL = 15, B_PL_EMPTY = 17;
  constexpr uint32_t TMEM_COLS = 512;                    // D1: 2 buffers x 2 row tiles x 64 columns; D2: 64 columns at 256
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_IN_FULL + i), 1); mbar_init(BAR(B_IN_EMPTY + i), 1); }
      mbar_init(BAR(B_RES), 1);
      for (int i = 0; i < 4; ++i) { mbar_init(BAR(B_D1_FULL + i), 1); mbar_init(BAR(B_D1_EMPTY + i), NCTA * FF_EPI1_WARPS); }
      for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_PL_FULL + i), NCTA * FF_EPI1_WARPS); mbar_init(BAR(B_PL_EMPTY + i), 1); }
      mbar_init(BAR(B_D2_FULL), 1); mbar_init(BAR(B_D2_EMPTY), NCTA * FF_EPI2_WARPS);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int i = tid; i < 64; i += FF_THREADS) { sBias[i] = p.bias1_x[i]; sBias[64 + i] = p.bias2_x[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Epilogue warps: TMEM lane quadrant q = warp % 4 (hardware rule).  Warps 2..9 run the conv-1 epilogue, warps
  // 10..17 the conv-2 epilogue (32 accumulator columns each), so the D1 -> planes hand-off that conv 2 waits for
  // never queues behind global stores.  The accumulators are pre-loaded with the bias (tcgen05.st): every MMA
  // accumulates, the epilogues do no bias add, and each drain re-arms its columns.
  const int q = warp & 3, half = ((warp - 2) >> 2) & 1;
  const bool epi2_warp = warp >= 2 + FF_EPI1_WARPS;
  const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
  uint32_t biasu[32];                                     // this warp's 32 columns of bias1 (conv-1 side) / bias2
  if (warp >= 2) {
#pragma unroll
    for (int c = 0; c < 32; ++c) biasu[c] = __float_as_uint(sBias[(epi2_warp ? 64 : 0) + half * 32 + c]);
    if (!epi2_warp) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { tmem_st16(t_lane + (uint32_t)(i * 64), biasu); tmem_st16(t_lane + (uint32_t)(i * 64 + 16), biasu + 16); }
    } else {
      tmem_st16(t_lane + 256u, biasu); tmem_st16(t_lane + 272u, biasu + 16);
    }
    tmem_st_wait();
  }
  if (warp == 0 && lane == 0) {
    // resident B operands: all 64 N rows of every stage, or this CTA's 32 rows of them (PAIR)
    tma_prefetch_desc(&tmap);
    mbar_arrive_expect_tx(BAR(B_RES), (uint32_t)((B1_STAGES + B2_STAGES) * BST));
    for (int s = 0; s < B1_STAGES; ++s)
      bulk_g2s(smem_u32(sB1 + s * BST), p.b1_image + (size_t)s * BSTAGE + (size_t)crank * BST, BST, BAR(B_RES));
    for (int s = 0; s < B2_STAGES; ++s)
      bulk_g2s(smem_u32(sB2 + s * BST), p.b2_image + (size_t)s * BSTAGE + (size_t)crank * BST, BST, BAR(B_RES));
    if (PAIR) mbar_wait(BAR(B_RES), 0);        // the leader's MMAs read the peer's half: resident before the cluster sync
  }
  __syncwarp();                                // barrier.cluster is .aligned: every warp arrives converged
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  pdl_wait();                 // everything above touched only shared memory, TMEM and weights
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      long long pw = 0;
      const long long pbeg = ff_clock();
      for (int k = 0; k < my_tiles; ++k) {
        const int tile = blockIdx.x + k * gridDim.x;
        const int img = fast_div(tile, p.magic_tiles), rem = tile - img * tiles_per_img;
        const int ty = fast_div(rem, p.magic_tx), tx = rem - ty * p.TX;
        const long long t0 = ff_clock();
        const int ib = k & 1;
        mbar_wait(BAR(B_IN_EMPTY + ib), ((k >> 1) & 1) ^ 1);  // conv-1 MMAs of tile k-2 have read this buffer
        pw += ff_clock() - t0;
        if (PAIR) {   // both CTAs' planes complete on the leader's barrier, which the leader arms for both
          if (cta_leader) mbar_arrive_expect_tx(BAR(B_IN_FULL + ib), 2 * IN_BYTES);
          tma_load_4d_pair(smem_u32(sIn + ib * IN_BYTES), &tmap, 0, 7 * tx, 28 * ty, img, BAR(B_IN_FULL + ib));
        } else {
          mbar_arrive_expect_tx(BAR(B_IN_FULL + ib), IN_BYTES);
          tma_load_4d(smem_u32(sIn + ib * IN_BYTES), &tmap, 0, 7 * tx, 28 * ty, img, BAR(B_IN_FULL + ib));
        }
      }
      if (p.dbg) { p.dbg[blockIdx.x * 24 + 0] = pw; p.dbg[blockIdx.x * 24 + 1] = ff_clock() - pbeg; }
    }
  } else if (warp == 1 && cta_leader) {
    // ===================== MMA issuer (the leader CTA of a pair issues for both) =====================
    // Issue order: conv1(0), then per tile k: conv1(k+1), conv2(k).  conv 2 of tile k has to wait for the conv-1
    // epilogue of tile k (TMEM -> registers -> planes); with conv 1 of the NEXT tile queued in front of it the tensor
    // pipe works through that wait instead of idling (D1 is double-buffered in TMEM for this).
    constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, 64);
    constexpr uint64_t DESC_HI = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
      if (PAIR) tc_mma_bf16_pair(d, a, b, id, 1u); else tc_mma_bf16(d, a, b, id, 1u);   // accumulators hold the bias
    };
    auto commit = [](uint32_t bar) { if (PAIR) tc_commit_pair(bar); else tc_commit(bar); };
    auto wait = [](uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); };
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t in16 = (smem_u32(sIn) & 0x3FFFFu) >> 4, b1_16 = (smem_u32(sB1) & 0x3FFFFu) >> 4;
    const uint32_t b2_16 = (smem_u32(sB2) & 0x3FFFFu) >> 4, pl16 = (smem_u32(sPl) & 0x3FFFFu) >> 4;
    if (!PAIR) mbar_wait(BAR(B_RES), 0);
    long long mw_in = 0, mw_d1e = 0, mw_pl = 0, mw_d2e = 0, tq;
    const long long mbeg = ff_clock();
    const bool leader = elect_one();
    auto conv1 = [&](int k) {
      const int buf = k & 1;
      if (lane == 0) FF_TR(0, k, 0);                        // conv1(k) issue starts
      tq = ff_clock();
      wait(BAR(B_IN_FULL + buf), (k >> 1) & 1);
      mw_in += ff_clock() - tq;
      tc_fence_after();
#pragma unroll
      for (int jt = 0; jt < 2; ++jt) {
        tq = ff_clock();
        wait(BAR(B_D1_EMPTY + buf * 2 + jt), ((k >> 1) & 1) ^ 1);        // epilogue has drained this D1 of tile k-2
        mw_d1e += ff_clock() - tq;
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int ky = 0; ky < 7; ++ky)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint32_t alo = in16 + (uint32_t)(buf * (IN_BYTES / 16) + (16 * jt + ky) * (8 * 128 / 16) + kk * 2);
              const uint32_t blo = b1_16 + (uint32_t)((ky >> 1) * (BST / 16) + (ky & 1) * 4 + kk * 2);
              mma(tmem_u + (uint32_t)(buf * 128 + jt * 64), DESC_HI | (uint64_t)alo, DESC_HI | (uint64_t)blo, idesc);
            }
          commit(BAR(B_D1_FULL + buf * 2 + jt));
          if (jt == 1) commit(BAR(B_IN_EMPTY + buf));         // this input buffer may be refilled
        }
        __syncwarp();
      }
      if (lane == 0) FF_TR(0, k, 1);                        // conv1(k) issued
    };
    if (my_tiles > 0) conv1(0);
    for (int k = 0; k < my_tiles; ++k) {
      if (k + 1 < my_tiles) conv1(k + 1);
      // ---- conv 2: 5 taps x (4 + KS2B) k-steps on the planes written by the conv-1 epilogue ----
      const int pb = k % NPL;
      tq = ff_clock();
      if (lane == 0) FF_TR(0, k, 2);                        // waiting planes_full(k)
      wait(BAR(B_PL_FULL + pb), (k / NPL) & 1);
      if (lane == 0) FF_TR(0, k, 3);                        // planes_full(k) seen
      mw_pl += ff_clock() - tq;
      tq = ff_clock();
      wait(BAR(B_D2_EMPTY), (k & 1) ^ 1);                     // D2 of the previous tile has been read
      mw_d2e += ff_clock() - tq;
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
              const uint32_t blo = b2_16 + (uint32_t)((ky * 2 + c) * (BST / 16) + kk * 2);
              mma(tmem_u + 256u, DESC_HI | (uint64_t)alo, DESC_HI | (uint64_t)blo, idesc);
            }
        }
        commit(BAR(B_D2_FULL));
        commit(BAR(B_PL_EMPTY + pb));                         // these planes may be rewritten
      }
      __syncwarp();
      if (lane == 0) FF_TR(0, k, 4);                        // conv2(k) issued
    }
    if (p.dbg && lane == 0) {
      unsigned long long* d = p.dbg + blockIdx.x * 24;
      d[2] = mw_in; d[3] = mw_d1e; d[4] = mw_pl; d[5] = mw_d2e; d[6] = ff_clock() - mbeg;
    }
    tc_fence_before();
  } else if (warp >= 2) {
    // ===================== epilogue warps (2..17) =====================
    // arrivals on the barriers the MMA issuer waits on go to the leader CTA of the pair
    auto arrive_mma = [](uint32_t bar) { if (PAIR) mbar_arrive_leader(bar); else mbar_arrive(bar); };
    const int m = q * 32 + lane;                              // accumulator row
    // tile coordinates advance incrementally (no divisions in the loop)
    const int step_img = (int)gridDim.x / tiles_per_img, step_rem = (int)gridDim.x - step_img * tiles_per_img;
    int img = (int)blockIdx.x / tiles_per_img, rem = (int)blockIdx.x - img * tiles_per_img;
    const uint32_t mtx = (uint32_t)((65536 + p.TX - 1) / p.TX);                                     // rem < 65536 / TX
    long long laps[6] = {0, 0, 0, 0, 0, 0};
    long long t_prev = ff_clock();
    const long long ebeg = t_prev;
    auto lap = [&](int i) {
      if (UAHN_FF_PROFILE) { const long long now = clock64(); laps[i] += now - t_prev; t_prev = now; }
    };
    if (!epi2_warp) {
      // ---- conv-1 epilogue: D1[buf][jt] -> LeakyReLU -> bf16 planes (conv 2's A operand) ----
      const int rr1 = m >> 3, g = m & 7;                      // row within the 16-row tile, group
      // plane rows of this thread's two conv-1 rows (j1 = rr1, 16 + rr1): plane rho = j1 & 1, row (j1 >> 1) * 8 + g;
      // this warp's 32 columns are 16-byte chunks 4*half .. 4*half+3, address-swizzled with (row & 7) = g
      const uint32_t pl_addr = smem_u32(sPl);
      uint32_t prow[2], pch[4];
#pragma unroll
      for (int jt = 0; jt < 2; ++jt) {
        const int j1 = 16 * jt + rr1;
        prow[jt] = pl_addr + (uint32_t)((j1 & 1) * PLANE_BYTES + ((j1 >> 1) * 8 + g) * 128);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) pch[c] = (uint32_t)(((4 * half + c) ^ g) << 4);
      for (int k = 0; k < my_tiles; ++k) {
        const int ty = (int)(((uint32_t)rem * mtx) >> 16), tx = rem - ty * p.TX;
        const int buf = k & 1;
        const int y0 = 28 * ty - 2, x0 = G1 * 7 * tx - 2;     // first conv-1 row / pixel of the tile region
        // conv 2's zero padding: conv-1 outputs outside the image must be stored as zeros (border tiles only)
        const bool border = y0 < 0 || y0 + 31 >= p.H1 || x0 < 0 || x0 + G1 * 8 - 1 >= p.W1;
        const int pb = k % NPL;
        if (warp == 2 && lane == 0) FF_TR(1 + blockIdx.x, k, 0);
        mbar_wait(BAR(B_PL_EMPTY + pb), ((k / NPL) & 1) ^ 1);   // conv 2 of tile k - NPL has read these planes
        if (warp == 2 && lane == 0) FF_TR(1 + blockIdx.x, k, 1);
        lap(0);
#pragma unroll
        for (int jt = 0; jt < 2; ++jt) {
          mbar_wait(BAR(B_D1_FULL + buf * 2 + jt), (k >> 1) & 1);
          if (warp == 2 && lane == 0) FF_TR(1 + blockIdx.x, k, 2 + jt);
          lap(1);
          tc_fence_after();
          uint32_t r[32];
          tmem_ld32(t_lane + (uint32_t)(buf * 128 + jt * 64), r);
          tmem_ld_wait();
          lap(4);
          uint32_t packed[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) packed[e] = pack_lrelu_bf16x2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
          if (border) {
            const bool yok = (unsigned)(y0 + 16 * jt + rr1) < (unsigned)p.H1;
#pragma unroll
            for (int c = 0; c < 4; ++c) {                     // 16-byte chunk c = columns half*32 + 8c .. +7: one pixel
              const int x = x0 + G1 * g + (half * 32 + 8 * c) / C1;
              const bool ok = yok && (unsigned)x < (unsigned)p.W1;
#pragma unroll
              for (int e = 0; e < 4; ++e) packed[4 * c + e] = ok ? packed[4 * c + e] : 0u;
            }
          }
          lap(5);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st_shared_v4(prow[jt] + (uint32_t)(pb * 2 * PLANE_BYTES) + pch[c], packed[4 * c], packed[4 * c + 1], packed[4 * c + 2],
                         packed[4 * c + 3]);
          lap(2);
        }
        fence_proxy_async();                                  // generic-proxy writes -> visible to the UMMA reads
        __syncwarp();
        if (lane == 0) arrive_mma(BAR(B_PL_FULL + pb));
        if (warp == 2 && lane == 0) FF_TR(1 + blockIdx.x, k, 4);
        // off the critical path: re-arm both D1 tiles with the bias and hand them back to the MMA warp
#pragma unroll
        for (int jt = 0; jt < 2; ++jt) {
          tmem_st16(t_lane + (uint32_t)(buf * 128 + jt * 64), biasu);
          tmem_st16(t_lane + (uint32_t)(buf * 128 + jt * 64 + 16), biasu + 16);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { arrive_mma(BAR(B_D1_EMPTY + buf * 2)); arrive_mma(BAR(B_D1_EMPTY + buf * 2 + 1)); }
        if (warp == 2 && lane == 0) FF_TR(1 + blockIdx.x, k, 5);
        rem += step_rem; img += step_img;
        if (rem >= tiles_per_img) { rem -= tiles_per_img; ++img; }
        lap(3);
      }
    } else {
      // ---- conv-2 epilogue: D2 -> LeakyReLU -> bf16 -> global; each thread stores its 64 contiguous bytes ----
      const int rr2 = m >> 3, w2l = m & 7;
      const bool row_ok = rr2 < 14 && w2l < 7;
      const long long thr_off = p.out_origin_b + (long long)rr2 * p.out_pitch_y_b + (long long)w2l * 128 + half * 64;
      for (int k = 0; k < my_tiles; ++k) {
        const int ty = (int)(((uint32_t)rem * mtx) >> 16), tx = rem - ty * p.TX;
        mbar_wait(BAR(B_D2_FULL), k & 1);
        if (warp == 10 && lane == 0 && blockIdx.x == 0) FF_TR(3, k, 0);
        lap(0);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(t_lane + 256u, r);
        tmem_ld_wait();
        tmem_st16(t_lane + 256u, biasu);
        tmem_st16(t_lane + 272u, biasu + 16);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_mma(BAR(B_D2_EMPTY));
        if (warp == 10 && lane == 0 && blockIdx.x == 0) FF_TR(3, k, 1);
        lap(1);
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
        lap(2);
      }
    }
    if (p.dbg && lane == 0 && (warp == 2 || warp == 10)) {    // warp 2: conv-1 epilogue, warp 10: conv-2 epilogue
      unsigned long long* d = p.dbg + blockIdx.x * 24 + (warp == 2 ? 8 : 14);
      for (int i = 0; i < 4; ++i) d[i] = laps[i];
      d[4] = ff_clock() - ebeg;
      if (warp == 2) { d[12] = laps[4]; d[13] = laps[5]; }
    }
  }
  __syncwarp();
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // (PAIR: the peer's shared memory and TMEM are in use until the leader is done)
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
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

inline size_t sw128_byte(int stage, int n, int kk) {   // [stage][64 rows][128 B], K-major, 128B swizzle
  return ((size_t)stage * 64 + n) * 128 + (size_t)((((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2);
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
  // ---- tensor map over the block input: {64 el window, conv-1 group, padded row, image} ----
  PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode_ff();
  if (!encode) { err = "cuTensorMapEncodeTiled entry point not found"; return -2; }
  const cuuint64_t gdim[4] = {64, 48, (cuuint64_t)x.Hp, (cuuint64_t)x.N};
  const cuuint64_t gstr[3] = {(cuuint64_t)G1 * 2 * 2, (cuuint64_t)x.pitch_y() * 2, (cuuint64_t)x.pitch_n * 2};
  const cuuint32_t box[4] = {64, 8, IN_ROWS, 1}, estr[4] = {1, 1, 1, 1};
  if (encode(reinterpret_cast<CUtensorMap*>(plan.tmap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x.p, gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 0;
  // ---- B1: two taps per 64-element stage; element (n = i*C1 + co, k = (ky&1)*32 + xi*2 + c) = W1[co][c][ky][xi - i]
  std::vector<uint16_t> b1((size_t)B1_STAGES * 64 * 64, 0);
  for (int ky = 0; ky < 7; ++ky)
    for (int xi = 0; xi < G1 + 6; ++xi)
      for (int c = 0; c < 2; ++c)
        for (int i = 0; i < G1; ++i) {
          const int kx = xi - i;
          if (kx < 0 || kx >= 7) continue;
          for (int co = 0; co < C1; ++co) {
            const float w = wk0[(size_t)((ky * 7 + kx) * 2 + c) * C1 + co];
            b1[sw128_byte(ky >> 1, i * C1 + co, (ky & 1) * 32 + xi * 2 + c) / 2] = f32_to_bf16_host(w);
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
          b2[sw128_byte(ky * 2 + qel / 64, xo * C2 + co, qel % 64) / 2] = f32_to_bf16_host(w);
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
  const int tiles = n_img * p.TX * p.TY;
  const int grid = std::min(tiles, num_sms);
  static unsigned long long* d_dbg = nullptr;
  const bool debug = UAHN_FF_PROFILE && getenv("UAHN_FF_DEBUG") != nullptr;
  if (debug && !d_dbg) cudaMalloc(&d_dbg, 24 * 8 * 1024 + 8 * 4096);
  if (debug) cudaMemsetAsync(d_dbg, 0, 24 * 8 * 1024 + 8 * 4096, st);
  p.dbg = debug ? d_dbg : nullptr;
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(plan.tmap);
  // CTA pairs (cta_group::2) need an even grid and an even tile count (tiles per image are 48 / 24: always even)
  static const bool want_pair = getenv("UAHN_FF_NO_PAIR") == nullptr;
  const bool pair = want_pair && grid >= 2 && tiles % 2 == 0;
  auto launch = [&](auto kern, bool use_pair) -> cudaError_t {
    const int smem = smem_bytes(use_pair);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(use_pair ? (grid & ~1) : grid);
    cfg.blockDim = dim3(FF_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = use_pair ? 2 : 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, *tm, p);
  };
  cudaError_t lerr;
  if (plan.C1 == 8) lerr = pair ? launch(conv_fused_front_kernel<8, 2, true>, true) : launch(conv_fused_front_kernel<8, 2, false>, false);
  else lerr = pair ? launch(conv_fused_front_kernel<16, 3, true>, true) : launch(conv_fused_front_kernel<16, 3, false>, false);
  if (lerr != cudaSuccess) return lerr;
  if (debug) {
    std::vector<unsigned long long> h(24 * grid);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), d_dbg, h.size() * 8, cudaMemcpyDeviceToHost);
    const double tl = (double)tiles / grid;
    for (int par = 0; par < 2; ++par) {       // even (leader) and odd (peer) CTAs separately
      double a[24] = {0};
      int cnt = 0;
      for (int i = par; i < grid; i += 2, ++cnt) for (int j = 0; j < 24; ++j) a[j] += (double)h[i * 24 + j] / tl;
      for (int j = 0; j < 24; ++j) a[j] /= cnt;
      fprintf(stderr, "[uahn-ff] C1=%d %s CTAs, per tile (cycles): producer wait_in_empty %.0f of %.0f | mma wait in_full %.0f d1_empty %.0f planes_full %.0f d2_empty %.0f of %.0f\n",
              plan.C1, par ? "odd " : "even", a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
      fprintf(stderr, "[uahn-ff]   epi1(w2): wait planes free %.0f | wait d1_full %.0f | tmem ld %.0f | pack %.0f | sts %.0f | fence+arrive+rearm %.0f | total %.0f   epi2(w10): wait d2_full %.0f | drain %.0f | pack+store %.0f | total %.0f\n",
              a[8], a[9], a[20], a[21], a[10], a[11], a[12], a[14], a[15], a[16], a[18]);
    }
    {
      std::vector<unsigned long long> tr(4 * 8 * 8);
      cudaMemcpy(tr.data(), d_dbg + 24 * 1024, tr.size() * 8, cudaMemcpyDeviceToHost);
      const unsigned long long t0 = tr[0];
      if (t0)
        for (int k = 0; k < 8; ++k) {
          fprintf(stderr, "[uahn-ff-trace] tile %d:", 40 + k);
          for (int role = 0; role < 4; ++role) {
            fprintf(stderr, "  r%d", role);
            for (int ev = 0; ev < 6; ++ev) fprintf(stderr, " %lld", (long long)(tr[(role * 8 + k) * 8 + ev] ? tr[(role * 8 + k) * 8 + ev] - t0 : 0));
          }
          fprintf(stderr, "\n");
        }
    }
  }
  return cudaGetLastError();
}

}  // namespace uahn
