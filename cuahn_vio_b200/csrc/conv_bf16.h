// bf16 tcgen05 implicit-GEMM convolution (conv_bf16.cu): host-side operand preparation + launch.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

namespace uahn {

// Shifted-window TMA plan (conv_bf16_tma.cu): the tile is a BW x BR patch of output pixel groups of one image;
// every input row segment is loaded ONCE per tile by a 4-D tiled TMA (one "plane" per row parity and 64-element
// chunk of the K run) and the KH kernel rows are MMAs whose A descriptor is shifted by whole patch rows.
struct TmaPlan {
  int enabled = 0;
  alignas(64) unsigned char tmap[128];   // CUtensorMap over the haloed NHWC input (overlapping x windows)
  int xb = 1, BW = 8, BR = 16, PX = 0, PY = 0;
  int chunks = 1, run_elems = 0, n_planes = 0, n_slots = 0, slot_bytes = 0;
  int plane_rows[8], plane_taps[8], plane_rho[8], plane_chunk[8];
  void* b_image = nullptr;   // [KH*chunks][N][128 B]
  float* bias_x = nullptr;
  int n_total = 0;
};

// Fused block front (conv_fused_front.cu): 7x7 s1 (2 -> C1) + 5x5 s2 (C1 -> C2) in one kernel.
struct FusedPlan {
  int enabled = 0;
  alignas(64) unsigned char tmap[128];
  int C1 = 0, TX = 0, TY = 0, Wox2 = 0, H1 = 0, W1 = 0;
  void *b1_image = nullptr, *b2_image = nullptr;
  float *bias1_x = nullptr, *bias2_x = nullptr;
};
int conv_fused_prepare(FusedPlan& plan, const std::vector<float>& wk0, const std::vector<float>& bias0,
                       const std::vector<float>& wk1, const std::vector<float>& bias1, const ConvGeom& g0,
                       const ConvGeom& g1, const Tensor& x, std::vector<void*>& allocs, std::string& err);
cudaError_t launch_conv_fused(const FusedPlan& plan, void* out, const ConvGeom& g1, int n_img, int num_sms,
                              cudaStream_t st);

// block_2_1 (conv_s2_first.cu): 7x7 stride-2 conv of the 2-channel block input, TMA-staged, N = 128.
struct S2Plan {
  int enabled = 0;
  alignas(64) unsigned char tmap[128];
  int TX = 0, TY = 0, Ho = 0, Wog = 0;
  void* b_image = nullptr;
  float* bias_x = nullptr;
};
int conv_s2first_prepare(S2Plan& plan, const std::vector<float>& wk, const std::vector<float>& bias, const ConvGeom& g,
                         const Tensor& x, std::vector<void*>& allocs, std::string& err);
cudaError_t launch_conv_s2first(const S2Plan& plan, void* out, const ConvGeom& g, int num_sms, cudaStream_t st);

struct ConvBf16Weights {
  TmaPlan tma;
  S2Plan s2;
  void* b_image = nullptr;   // pre-swizzled B-operand stages in global memory
  int ready = 0;
  int xb = 1;                // output pixels per GEMM row (Toeplitz expansion along x)
  int n_total = 0;           // xb * Cout
  int bn = 0;                // N tile
  int k_total = 0;           // padded K (elements)
  int runs = 0, run_granules = 0;
  float* bias_x = nullptr;   // bias replicated xb times
  // im2col TMA A producer (conv_bf16.cu, Cin % 64 == 0 layers): tensor map over the haloed NHWC input
  int im2col = 0;
  alignas(64) unsigned char im2col_map[128];
  // CTA-pair mode: tiled tensor maps over the pre-swizzled B image with boxes of 128 / 64 / 32 rows (half an N tile)
  int pair_ok = 0;
  alignas(64) unsigned char b_half_map[3][128];
};

// wk: [K][Cout] fp32 with k = (ky*KW + kx)*Cin + c.  Appends device allocations to `allocs`.
int conv_bf16_prepare(ConvBf16Weights& wb, const std::vector<float>& wk, const std::vector<float>& bias,
                      const ConvGeom& g, const Tensor& in, const Tensor& out, std::vector<void*>& allocs,
                      std::string& err);
// conv_bf16_tma.cu
int conv_tma_prepare(TmaPlan& plan, const std::vector<float>& wk, const std::vector<float>& bias, const ConvGeom& g,
                     const Tensor& in, std::vector<void*>& allocs, std::string& err);
cudaError_t launch_conv_tma(const TmaPlan& plan, const float* bias, void* out, const ConvGeom& g, int num_sms,
                            cudaStream_t st);
uint16_t f32_to_bf16_host(float f);

cudaError_t launch_mc_gemm_bf16(const ConvBf16Weights& wb, const void* feat, const uint8_t* mc_bits, void* out, int n,
                                cudaStream_t st);
cudaError_t launch_conv_bf16(const ConvBf16Weights& wb, const void* in, const float* bias, void* out,
                             const ConvGeom& g, cudaStream_t st);
// conv_small_m.cu — latency path (M <= 160 GEMM rows, Cin % 64 == 0): split-K over a cluster of 8 CTAs, mma.sync
int conv_small_m_ok(const ConvBf16Weights& wb, const ConvGeom& g);
cudaError_t launch_conv_small_m(const ConvBf16Weights& wb, const void* in, const float* bias, void* out, const ConvGeom& g,
                                cudaStream_t st);

}  // namespace uahn
