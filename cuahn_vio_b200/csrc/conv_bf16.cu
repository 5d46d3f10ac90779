#include "conv_bf16.h"
namespace uahn {
int conv_bf16_prepare(ConvBf16Weights&, const std::vector<float>&, const std::vector<float>&, const ConvGeom&,
                      const Tensor&, const Tensor&, std::vector<void*>&, std::string& err) {
  err = "bf16 path not built yet";
  return -5;
}
cudaError_t launch_conv_bf16(const ConvBf16Weights&, const void*, const float*, void*, const ConvGeom&, cudaStream_t) {
  return cudaErrorNotSupported;
}
}  // namespace uahn
