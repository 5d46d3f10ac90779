// Fused front of cascade blocks 3 and 4: conv 7x7 s1 (2 -> C1) + LeakyReLU  ->  conv 5x5 s2 (C1 -> C2) + LeakyReLU
// in ONE persistent tcgen05 kernel.  The first conv's output (1.15 MB / 0.57 MB per pair, the largest activations of
// the network) never exists in HBM: its tiles go TMEM -> registers -> shared memory, laid out exactly as the second
// conv's A operand, and the second conv's MMAs read them from there.  This removes 31 % (block 4) + 15 % (block 3) of
// the conv DRAM traffic, which is what bounds the conv stacks.
//
// Tile = 7 x 14 patch of conv-2 pixel groups (one group = 64/C2 output pixels = G1 = 64/C1 conv-1 pixels = one
// 128-byte K row).  Per tile:
//   1. TMA: one 4-D box {64 el, 8 groups, 38 rows} of the 2-channel block input (overlapping x windows, halo 5).
//   2. conv 1 as two 128-row MMA tiles (rows 0-15 / 16-31 of the 32 x 8-group region conv 2 needs); kernel row ky is
//      the same plane shifted by ky rows (shifted-window trick of conv_bf16_tma.cu); weights packed two taps per
//      64-element B stage.
//   3. 16 epilogue warps: D1 -> +bias, LeakyReLU, zero outside the image (that is conv 2's zero padding) -> bf16 ->
//      shared-memory planes [row parity][t = row/2][group], 128B-swizzled by ADDRESS (the UMMA swizzle is purely
//      address-based — tools/umma_offset_test.cu — so operands may start at any 128-byte row).
//   4. conv 2: tap ky = rho + 2a, chunk c reads plane rho at row offset (a*8 + c): "chunk 1 of group w" is "chunk 0
//      of group w+1", so nothing is duplicated.  B2 (80 KB) and B1 (32 KB) stay resident in shared memory.
//   5. epilogue: D2 -> +bias, LeakyReLU -> bf16 -> staged, coalesced stores into the (haloed NHWC) conv-2 output.
// conv 1 of tile i+1 overlaps the conv-2 epilogue of tile i (separate TMEM accumulators, mbarrier hand-offs).
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

constexpr int FF_EPI_WARPS = 16;
constexpr int FF_THREADS = 64 + 32 * FF_EPI_WARPS;
constexpr int IN_ROWS = 38;                          // 32 conv-1 rows + 6 (7x7 halo)
constexpr int IN_BYTES = IN_ROWS * 8 * 128;          // 38 912
constexpr int B1_STAGES = 4, B2_STAGES = 10;         // 7 taps packed in pairs; 5 taps x 2 chunks
constexpr int BSTAGE = 64 * 128;                     // N = 64 rows x 128 B
constexpr int PLANE_ROWS = 18;                       // t = 0..15 written, +2 rows read only by dummy M rows
constexpr int PLANE_BYTES = PLANE_ROWS * 8 * 128;    // 18 432
constexpr int SROW = 64 * 2 + 16;                    // conv-2 epilogue staging row
constexpr int SMEM_BYTES = 1024 + IN_BYTES + B1_STAGES * BSTAGE + B2_STAGES * BSTAGE + 2 * PLANE_BYTES + 128 * SROW +
                           2 * 64 * 4 + 128 * 8 + 32 * 8;

#ifndef UAHN_FF_PROFILE
#define UAHN_FF_PROFILE 0
#endif
__device__ __forceinline__ long long ff_clock() { return UAHN_FF_PROFILE ? clock64() : 0ll; }

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

// C1: conv-1 output channels (8 or 16); KS2B: k-steps of conv-2 chunk 1 (2 for block 4, 3 for block 3)
template <int C1, int KS2B>
__global__ void __launch_bounds__(FF_THREADS, 1) conv_fused_front_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                          const __grid_constant__ FusedParams p) {
  constexpr int G1 = 64 / C1;                        // conv-1 pixels per group
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sIn = smem;                                             // conv-1 input plane
  uint8_t* sB1 = sIn + IN_BYTES;
  uint8_t* sB2 = sB1 + B1_STAGES * BSTAGE;
  uint8_t* sPl = sB2 + B2_STAGES * BSTAGE;                         // [2 parities][PLANE_BYTES]
  uint8_t* sOut = sPl + 2 * PLANE_BYTES;                           // [128][SROW]
  float* sBias = reinterpret_cast<float*>(sOut + 128 * SROW);      // [2][64]
  long long* sRowOff = reinterpret_cast<long long*>(sBias + 128);  // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRowOff + 128);
  // bars: 0 in_full, 1 in_empty, 2 bres, 3-4 d1_full, 5-6 d1_empty, 7 planes_full, 8 planes_empty, 9 d2_full, 10 d2_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_img = p.TX * p.TY;
  const int total_tiles = p.n_img * tiles_per_img;

  if (warp == 1) {
    if (lane == 0) {
      mbar_init(BAR(0), 1); mbar_init(BAR(1), 1); mbar_init(BAR(2), 1);
      mbar_init(BAR(3), 1); mbar_init(BAR(4), 1);
      mbar_init(BAR(5), FF_EPI_WARPS); mbar_init(BAR(6), FF_EPI_WARPS);
      mbar_init(BAR(7), FF_EPI_WARPS); mbar_init(BAR(8), 1);
      mbar_init(BAR(9), 1); mbar_init(BAR(10), FF_EPI_WARPS);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 64; i += FF_THREADS) { sBias[i] = p.bias1_x[i]; sBias[64 + i] = p.bias2_x[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmap);
      mbar_arrive_expect_tx(BAR(2), (uint32_t)((B1_STAGES + B2_STAGES) * BSTAGE));
      for (int s = 0; s < B1_STAGES; ++s) bulk_g2s(smem_u32(sB1 + s * BSTAGE), p.b1_image + (size_t)s * BSTAGE, BSTAGE, BAR(2));
      for (int s = 0; s < B2_STAGES; ++s) bulk_g2s(smem_u32(sB2 + s * BSTAGE), p.b2_image + (size_t)s * BSTAGE, BSTAGE, BAR(2));
      int tcount = 0;
      long long pw = 0;
      const long long pbeg = ff_clock();
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int img = fast_div(tile, p.magic_tiles), rem = tile - img * tiles_per_img;
        const int ty = fast_div(rem, p.magic_tx), tx = rem - ty * p.TX;
        const long long t0 = ff_clock();
        mbar_wait(BAR(1), (tcount & 1) ^ 1);                  // conv-1 MMAs of the previous tile have read the plane
        pw += ff_clock() - t0;
        mbar_arrive_expect_tx(BAR(0), IN_BYTES);
        tma_load_4d(smem_u32(sIn), &tmap, 0, 7 * tx, 28 * ty, img, BAR(0));
      }
      if (p.dbg) { p.dbg[blockIdx.x * 24 + 0] = pw; p.dbg[blockIdx.x * 24 + 1] = ff_clock() - pbeg; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    constexpr uint64_t DESC_HI = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t in16 = (smem_u32(sIn) & 0x3FFFFu) >> 4, b1_16 = (smem_u32(sB1) & 0x3FFFFu) >> 4;
    const uint32_t b2_16 = (smem_u32(sB2) & 0x3FFFFu) >> 4, pl16 = (smem_u32(sPl) & 0x3FFFFu) >> 4;
    mbar_wait(BAR(2), 0);
    int tcount = 0;
    long long mw_in = 0, mw_d1e = 0, mw_pl = 0, mw_d2e = 0, tq;
    const long long mbeg = ff_clock();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t par = tcount & 1;
      const bool leader = elect_one();
      // ---- conv 1: two 128-row tiles, 7 taps x 2 k-steps each ----
      tq = ff_clock();
      mbar_wait(BAR(0), par);
      mw_in += ff_clock() - tq;
      tc_fence_after();
#pragma unroll
      for (int jt = 0; jt < 2; ++jt) {
        tq = ff_clock();
        mbar_wait(BAR(5 + jt), par ^ 1);                      // epilogue has drained D1[jt] of the previous tile
        mw_d1e += ff_clock() - tq;
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int ky = 0; ky < 7; ++ky)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint32_t alo = in16 + (uint32_t)((16 * jt + ky) * (8 * 128 / 16) + kk * 2);
              const uint32_t blo = b1_16 + (uint32_t)((ky >> 1) * (BSTAGE / 16) + (ky & 1) * 4 + kk * 2);
              tc_mma_bf16(tmem_u + (uint32_t)(jt * 64), DESC_HI | (uint64_t)alo, DESC_HI | (uint64_t)blo, idesc,
                          (ky | kk) ? 1u : 0u);
            }
          tc_commit(BAR(3 + jt));
          if (jt == 1) tc_commit(BAR(1));                     // input plane may be refilled
        }
        __syncwarp();
      }
      // ---- conv 2: 5 taps x (4 + KS2B) k-steps on the planes written by the conv-1 epilogue ----
      tq = ff_clock();
      mbar_wait(BAR(7), par);
      mw_pl += ff_clock() - tq;
      tq = ff_clock();
      mbar_wait(BAR(10), par ^ 1);                            // D2 of the previous tile has been read
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
              const uint32_t alo = pl16 + (uint32_t)(rho * (PLANE_BYTES / 16) + (a * 8 + c) * (128 / 16) + kk * 2);
              const uint32_t blo = b2_16 + (uint32_t)((ky * 2 + c) * (BSTAGE / 16) + kk * 2);
              tc_mma_bf16(tmem_u + 128u, DESC_HI | (uint64_t)alo, DESC_HI | (uint64_t)blo, idesc,
                          (ky | c | kk) ? 1u : 0u);
            }
        }
        tc_commit(BAR(9));
        tc_commit(BAR(8));                                    // planes may be rewritten
      }
      __syncwarp();
    }
    if (p.dbg && lane == 0) {
      unsigned long long* d = p.dbg + blockIdx.x * 24;
      d[2] = mw_in; d[3] = mw_d1e; d[4] = mw_pl; d[5] = mw_d2e; d[6] = ff_clock() - mbeg;
    }
    tc_fence_before();
  } else {
    // ===================== epilogue warps (2..17) =====================
    const int q = warp & 3, cg = (warp - 2) >> 2;             // TMEM lane quadrant, 16-column group
    const int m = q * 32 + lane;                              // accumulator row
    const int rr1 = m >> 3, g = m & 7;                        // conv-1: row within the 16-row tile, group
    float bias1[16], bias2[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { bias1[c] = sBias[cg * 16 + c]; bias2[c] = sBias[64 + cg * 16 + c]; }
    uint8_t* qOut = sOut + q * 32 * SROW;
    long long* qRow = sRowOff + q * 32;
    const uint32_t pl_addr = smem_u32(sPl);
    int tcount = 0;
    long long ew_ple = 0, ew_d1[2] = {0, 0}, ec_e1 = 0, ew_d2 = 0, ec_e2 = 0, ec_st = 0, tq;
    const long long ebeg = ff_clock();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t par = tcount & 1;
      const int img = fast_div(tile, p.magic_tiles), rem = tile - img * tiles_per_img;
      const int ty = fast_div(rem, p.magic_tx), tx = rem - ty * p.TX;
      // ---- conv-1 epilogue: D1[jt] -> planes ----
      tq = ff_clock();
      mbar_wait(BAR(8), par ^ 1);                             // conv-2 MMAs of the previous tile have read the planes
      ew_ple += ff_clock() - tq;
#pragma unroll
      for (int jt = 0; jt < 2; ++jt) {
        tq = ff_clock();
        mbar_wait(BAR(3 + jt), par);
        ew_d1[jt] += ff_clock() - tq;
        tq = ff_clock();
        tc_fence_after();
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(jt * 64 + cg * 16), r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(5 + jt));
        const int j1 = 16 * jt + rr1;                         // conv-1 row inside the 32-row region
        const int y1 = 28 * ty - 2 + j1;
        const bool yok = (unsigned)y1 < (unsigned)p.H1;
        const int x1_0 = G1 * (7 * tx + g) - 2;               // first conv-1 pixel of this group
        uint32_t packed[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          // columns n = cg*16 + 2e, +1  <->  pixel i = n / C1 of the group, channel n % C1
          const int i = (cg * 16 + 2 * e) / C1;
          const bool ok = yok && (unsigned)(x1_0 + i) < (unsigned)p.W1;
          float v0 = __uint_as_float(r[2 * e]) + bias1[2 * e], v1 = __uint_as_float(r[2 * e + 1]) + bias1[2 * e + 1];
          v0 = fmaxf(v0, v0 * LRELU_SLOPE);
          v1 = fmaxf(v1, v1 * LRELU_SLOPE);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(ok ? v0 : 0.f, ok ? v1 : 0.f);
          packed[e] = *reinterpret_cast<uint32_t*>(&h2);
        }
        // plane rho = j1 & 1, row R = (j1 >> 1) * 8 + g, 16-byte chunks 2cg, 2cg+1, address-swizzled with R & 7 = g
        const uint32_t row = pl_addr + (uint32_t)((j1 & 1) * PLANE_BYTES + ((j1 >> 1) * 8 + g) * 128);
        st_shared_v4(row + (uint32_t)(((2 * cg) ^ g) << 4), packed[0], packed[1], packed[2], packed[3]);
        st_shared_v4(row + (uint32_t)(((2 * cg + 1) ^ g) << 4), packed[4], packed[5], packed[6], packed[7]);
        ec_e1 += ff_clock() - tq;
      }
      fence_proxy_async();                                    // generic-proxy writes -> visible to the UMMA reads
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(7));
      // ---- conv-2 epilogue: D2 -> global ----
      if (cg == 0) {
        const int rr = m >> 3, w = m & 7, w2 = 7 * tx + w;
        long long off = -1;
        if (rr < 14 && w < 7 && w2 < p.Wox2)
          off = p.out_origin_b + (long long)img * p.out_pitch_n_b + (long long)(14 * ty + rr) * p.out_pitch_y_b +
                (long long)w2 * 128;
        qRow[lane] = off;
      }
      tq = ff_clock();
      mbar_wait(BAR(9), par);
      ew_d2 += ff_clock() - tq;
      tq = ff_clock();
      tc_fence_after();
      {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(128 + cg * 16), r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(10));
        uint32_t packed[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v0 = __uint_as_float(r[2 * e]) + bias2[2 * e], v1 = __uint_as_float(r[2 * e + 1]) + bias2[2 * e + 1];
          v0 = fmaxf(v0, v0 * LRELU_SLOPE);
          v1 = fmaxf(v1, v1 * LRELU_SLOPE);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
          packed[e] = *reinterpret_cast<uint32_t*>(&h2);
        }
        uint4* o = reinterpret_cast<uint4*>(qOut + lane * SROW + cg * 32);
        o[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        o[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
      }
      ec_e2 += ff_clock() - tq;
      tq = ff_clock();
      asm volatile("bar.sync %0, %1;" ::"r"(q + 1), "r"(128) : "memory");
      {
        const int rsub = lane >> 3, ch = lane & 7;            // 8 x 16-byte chunks per 128-byte row, 4 rows per store
#pragma unroll
        for (int r0 = 0; r0 < 8; r0 += 4) {
          const int row = cg * 8 + r0 + rsub;
          const long long off = qRow[row];
          const uint4 v = *reinterpret_cast<const uint4*>(qOut + row * SROW + ch * 16);
          if (off >= 0) *reinterpret_cast<uint4*>(p.out + off + ch * 16) = v;
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(q + 1), "r"(128) : "memory");
      ec_st += ff_clock() - tq;
    }
    if (p.dbg && warp == 2 && lane == 0) {
      unsigned long long* d = p.dbg + blockIdx.x * 24;
      d[8] = ew_ple; d[9] = ew_d1[0]; d[10] = ew_d1[1]; d[11] = ec_e1; d[12] = ew_d2; d[13] = ec_e2; d[14] = ec_st;
      d[15] = ff_clock() - ebeg;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
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
  if (debug && !d_dbg) cudaMalloc(&d_dbg, 24 * 8 * 1024);
  if (debug) cudaMemsetAsync(d_dbg, 0, 24 * 8 * 1024, st);
  p.dbg = debug ? d_dbg : nullptr;
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(plan.tmap);
  static bool attr8 = false, attr16 = false;
  if (plan.C1 == 8) {
    if (!attr8) {
      cudaError_t e = cudaFuncSetAttribute(conv_fused_front_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e != cudaSuccess) return e;
      attr8 = true;
    }
    conv_fused_front_kernel<8, 2><<<grid, FF_THREADS, SMEM_BYTES, st>>>(*tm, p);
  } else {
    if (!attr16) {
      cudaError_t e = cudaFuncSetAttribute(conv_fused_front_kernel<16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e != cudaSuccess) return e;
      attr16 = true;
    }
    conv_fused_front_kernel<16, 3><<<grid, FF_THREADS, SMEM_BYTES, st>>>(*tm, p);
  }
  if (debug) {
    std::vector<unsigned long long> h(24 * grid);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), d_dbg, h.size() * 8, cudaMemcpyDeviceToHost);
    double a[24] = {0};
    const double tl = (double)tiles / grid;
    for (int i = 0; i < grid; ++i) for (int j = 0; j < 24; ++j) a[j] += (double)h[i * 24 + j] / grid / tl;
    fprintf(stderr, "[uahn-ff] C1=%d tiles/CTA=%.1f per tile (cycles): producer wait_in_empty %.0f of %.0f | mma wait in_full %.0f d1_empty %.0f planes_full %.0f d2_empty %.0f of %.0f | epi(w2) wait planes_empty %.0f d1_full0 %.0f d1_full1 %.0f epi1 %.0f wait d2_full %.0f epi2 %.0f store %.0f of %.0f\n",
            plan.C1, tl, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15]);
  }
  return cudaGetLastError();
}

}  // namespace uahn
