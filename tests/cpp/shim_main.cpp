// Host program exercising the C++ drop-in class (include/HomographyNet.h) the way VioManager does
// (VioManager.cpp:107,188,236,257-259): ctor, load_current_img per frame, network_inference, getters.
// usage: shim_main <weights.bin> <frames.u8 (k x 224 x 320)> <k> <priors.f64 (k x 8)> <precision> [<iterative weights.bin>]
// With an iterative model (max_IEKF_iteration = 2) every frame also runs iteration 1 on the second model slot
// (HomographyNet.cpp:209-230) and prints its outputs as ITER lines.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "HomographyNet.h"

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  std::string model = argv[1], iter_model = "";
  const int k = atoi(argv[3]);
  std::vector<unsigned char> frames((size_t)k * UAHN_IMG_PIXELS);
  std::vector<double> priors((size_t)k * 8);
  FILE* f = fopen(argv[2], "rb");
  if (!f || fread(frames.data(), 1, frames.size(), f) != frames.size()) return 3;
  fclose(f);
  f = fopen(argv[4], "rb");
  if (!f || fread(priors.data(), 8, priors.size(), f) != priors.size()) return 3;
  fclose(f);
  if (argc > 6) iter_model = argv[6];
  pytorch::HomographyNet net(model, iter_model, /*use_prior=*/true, /*num_of_iteration=*/argc > 6 ? 2 : 1, /*show_imgs=*/false,
                             atoi(argv[5]));
  net.set_seed(9);
  for (int i = 0; i < k; ++i) {
    net.load_current_img(frames.data() + (size_t)i * UAHN_IMG_PIXELS, UAHN_IMG_H, UAHN_IMG_W, UAHN_IMG_W, 10.0 + i);
    net.network_inference(priors.data() + (size_t)i * 8, 0);
    printf("RESULT %d %d %.3f", i, net.img_counter, net.get_latest_inference_time());
    for (int j = 0; j < 8; ++j) printf(" %.9g", net.pred_mean()[j]);
    for (int j = 0; j < 8; ++j) printf(" %.9g", net.pred_cov()[j * 9]);
    printf("\n");
    if (argc > 6 && net.img_counter >= 2) {
      net.network_inference(priors.data() + (size_t)i * 8, 1);
      printf("ITER %d", i);
      for (int j = 0; j < 8; ++j) printf(" %.9g", net.pred_mean()[j]);
      printf("\n");
    }
  }
  return 0;
}
