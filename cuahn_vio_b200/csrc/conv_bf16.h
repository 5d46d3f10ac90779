// bf16 tcgen05 implicit-GEMM convolution (conv_bf16.cu): host-side operand preparation + launch.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

namespace uahn {

struct ConvBf16Weights {
  void* b_image = nullptr;   // pre-swizzled B-operand stages in global memory
  int ready = 0;
  int xb = 1;                // output pixels per GEMM row (Toeplitz expansion along x)
  int n_total = 0;           // xb * Cout
  int bn = 0;                // N tile
  int k_total = 0;           // padded K (elements)
  int runs = 0, run_granules = 0;
  float* bias_x = nullptr;   // bias replicated xb times
};

// wk: [K][Cout] fp32 with k = (ky*KW + kx)*Cin + c.  Appends device allocations to `allocs`.
int conv_bf16_prepare(ConvBf16Weights& wb, const std::vector<float>& wk, const std::vector<float>& bias,
                      const ConvGeom& g, const Tensor& in, const Tensor& out, std::vector<void*>& allocs,
                      std::string& err);
cudaError_t launch_conv_bf16(const ConvBf16Weights& wb, const void* in, const float* bias, void* out,
                             const ConvGeom& g, cudaStream_t st);

}  // namespace uahn
