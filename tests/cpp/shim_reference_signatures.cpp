// Type-checks the reference-signature flavour of the drop-in class (include/HomographyNet.h with
// -DUAHN_WITH_EIGEN_OPENCV) against the calls cuahn's VioManager makes (VioManager.cpp:107,188,230-236,257-259,288):
// if this translation unit compiles, `#include "HomographyNet.h"` in VioManager.h keeps compiling unchanged.
// Built with -fsyntax-only against the stub headers in tests/cpp/stubs (the container has neither Eigen nor OpenCV).
#include <memory>
#include <string>

#include "HomographyNet.h"

double vio_manager_like(cv::Mat& img, double timestamp) {
  std::string network_model_path = "traced_model_3_blocks_using_prior_showError.bin", network_model_iterative_path = "";
  bool use_prior = true, show_img = false;
  int max_IEKF_iteration = 1;
  // VioManager.cpp:107
  std::unique_ptr<pytorch::HomographyNet> HNet(
      new pytorch::HomographyNet(network_model_path, network_model_iterative_path, use_prior, max_IEKF_iteration, show_img));
  HNet->load_current_img(img, timestamp);                                   // VioManager.cpp:188
  Eigen::Matrix<double, 8, 1> prior_4pt_offset_vec;                          // VioManager.cpp:230-234
  for (int i = 0; i < 8; ++i) prior_4pt_offset_vec(i) = 0.0;
  HNet->network_inference(prior_4pt_offset_vec, 0);                          // VioManager.cpp:236
  double used = 0;
  if (HNet->get_latest_inference_time() == timestamp && HNet->img_counter > 10) {   // VioManager.cpp:257
    Eigen::Matrix<double, 8, 1> pred_mean = HNet->get_pred_mean();           // VioManager.cpp:258
    Eigen::Matrix<double, 8, 8> pred_Cov = HNet->get_pred_Cov();             // VioManager.cpp:259
    used = pred_mean(0) + pred_Cov(0, 0);
  }
  return used + HNet->img_counter;                                           // VioManager.cpp:288
}
