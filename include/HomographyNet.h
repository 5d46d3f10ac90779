// Drop-in C++ replacement for the reference's `pytorch::HomographyNet`
// (cuahn_ros/homography_network/src/HomographyNet.h:23-67, HomographyNet.cpp:7-252) on top of the C ABI in uahn.h.
//
// Same class name, namespace, public members and argument meaning, so `cuahn`'s VioManager
// (VioManager.cpp:107,188,236,257-259) compiles against it unchanged.  The libtorch members are gone: the
// TorchScript `forward` (HomographyNet.cpp:183) is replaced by uahn_infer(), images live in a 2-slot device ring.
//
// Two flavours:
//   * with -DUAHN_WITH_EIGEN_OPENCV (the ROS build, where Eigen and OpenCV exist) the signatures are exactly the
//     reference's (cv::Mat, Eigen::Matrix<double,8,1> ...);
//   * otherwise a dependency-free flavour with raw pointers (what the tests in this repository compile, since this
//     container has neither Eigen nor OpenCV — SURVEY §0 fact 7).
//
// `network_model_path` now names the flat weight file exported once from the reference checkpoint by
// cuahn_vio_b200.weights.export_checkpoint(); the `_showError` file-name convention of HomographyNet.cpp:96-100 is
// kept.  Errors: like the reference, nothing throws — failures print to stderr and leave the outputs untouched.
#ifndef UAHN_HOMOGRAPHYNET_H_
#define UAHN_HOMOGRAPHYNET_H_

#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "uahn.h"

#ifdef UAHN_WITH_EIGEN_OPENCV
#include <Eigen/Eigen>
#include <opencv2/core/core.hpp>
#include <opencv2/highgui/highgui.hpp>
#endif

namespace pytorch {

class HomographyNet {
 public:
  HomographyNet(std::string& network_model_path, std::string& network_model_iterative_path, bool use_prior,
                int num_of_iteration, bool show_imgs, int precision = UAHN_PRECISION_BF16, int device = 0,
                int iterative_variant = UAHN_VARIANT_AUTO) {
    use_prior_4pt_offset = use_prior;
    cv_imshow = show_imgs;
    std::cout << "CUDA (sm_100a) UAHN engine: running on GPU " << device << "." << std::endl;
    show_phtometric_error = network_model_path.find("_showError") != std::string::npos;   // HomographyNet.cpp:96-100
    main_ = create(network_model_path, use_prior ? UAHN_VARIANT_PRIOR3 : UAHN_VARIANT_FULL, show_phtometric_error,
                   precision, device);
    if (num_of_iteration > 1) {                                                           // HomographyNet.cpp:20-24
      iteration = true;
      const bool it_err = network_model_iterative_path.find("_showError") != std::string::npos;
      // The reference's second slot runs whatever graph the file holds (HomographyNet.cpp:104-124; blocks_to_run in
      // {1, 2, 3}, model_to_trace.py:72,131-132).  UAHN_VARIANT_AUTO takes the variant the exporter recorded in the flat
      // file (weights.export_torchscript); files without a record (exported from a bare checkpoint) fall back to the
      // 2-block schedule.
      iter_ = create(network_model_iterative_path, iterative_variant, it_err, precision, device, /*quiet=*/iterative_variant == UAHN_VARIANT_AUTO);
      if (!iter_ && iterative_variant == UAHN_VARIANT_AUTO)
        iter_ = create(network_model_iterative_path, UAHN_VARIANT_PRIOR2, it_err, precision, device);
      std::cout << "IEKF! Load the Network for Iteration!" << std::endl;
    }
    // warm-up forward on constant images 0.2 / 0.5 and an all-ones prior (HomographyNet.cpp:29-63)
    if (main_) {
      static unsigned char a[UAHN_IMG_PIXELS], b[UAHN_IMG_PIXELS];
      memset(a, 51, sizeof(a));
      memset(b, 128, sizeof(b));
      const auto t0 = std::chrono::steady_clock::now();
      for (uahn_handle* h : {main_, iter_}) {
        if (!h) continue;
        uahn_load_image(h, a, UAHN_IMG_H, UAHN_IMG_W, UAHN_IMG_W, -1.0);
        uahn_load_image(h, b, UAHN_IMG_H, UAHN_IMG_W, UAHN_IMG_W, -1.0);
        double ones[8] = {1, 1, 1, 1, 1, 1, 1, 1}, m[8], c[64];
        if (uahn_infer(h, ones, nullptr, m, c, nullptr) != UAHN_OK) std::cerr << uahn_last_error(h) << std::endl;
      }
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      printf("[TIME]: %.4f milliseconds for the first network inference\n", ms);
      warmup_images_ = 2;
    }
  }
  ~HomographyNet() {
    uahn_destroy(main_);
    uahn_destroy(iter_);
    std::cout << "HomographyNet Object is being deleted! End of this run ..." << std::endl;
  }
  HomographyNet(const HomographyNet&) = delete;
  HomographyNet& operator=(const HomographyNet&) = delete;

  int img_counter = 0;   // HomographyNet.h:33 (public in the reference; VioManager.cpp:257,288-299 reads it)

  double get_latest_inference_time() { return _latest_inference_time_stamp; }

#ifdef UAHN_WITH_EIGEN_OPENCV
  void load_current_img(const cv::Mat& img, const double& time_stamp) {
    if (cv_imshow) { cv::imshow("Image", img); cv::waitKey(1); }
    load_current_img(img.data, img.rows, img.cols, img.step, time_stamp);
  }
  Eigen::Matrix<double, 8, 1> get_pred_mean() { return Eigen::Map<Eigen::Matrix<double, 8, 1>>(_pred_mean); }
  // The reference maps the row-major 8x8 as column-major (HomographyNet.cpp:197); Cov is symmetric, keep that.
  Eigen::Matrix<double, 8, 8> get_pred_Cov() { return Eigen::Map<Eigen::Matrix<double, 8, 8>>(_pred_Cov); }
  void network_inference(Eigen::Matrix<double, 8, 1>& prior_4pt_offset_vec, int num_of_inference) {
    network_inference(prior_4pt_offset_vec.data(), num_of_inference);
    if (cv_imshow && show_phtometric_error && have_err_) {
      cv::Mat resultImg(UAHN_IMG_H, UAHN_IMG_W, CV_8UC1, err_map_);
      cv::imshow("Photometric Error", resultImg);
      cv::waitKey(1);
    }
  }
#endif

  // ---- dependency-free flavour (same semantics, raw buffers) ----
  void load_current_img(const unsigned char* gray, int rows, int cols, size_t stride, const double& time_stamp) {
    img_counter++;
    if (img_counter == 1) std::cout << "First Image Comes into the Network Object!" << std::endl;
    for (uahn_handle* h : {main_, iter_})
      if (h && uahn_load_image(h, gray, rows, cols, stride, time_stamp) != UAHN_OK)
        std::cerr << "load_current_img: " << uahn_last_error(h) << std::endl;
    if (img_counter >= 2) _latest_inference_time_stamp = time_stamp;   // HomographyNet.cpp:148
  }
  void network_inference(const double* prior_4pt_offset_vec, int num_of_inference) {
    if (img_counter < 2) {   // HomographyNet.cpp:155-158
      std::cout << "HNet cannot inference! Only has one image!" << std::endl;
      return;
    }
    uahn_handle* h = (num_of_inference == 0 || !iter_) ? main_ : iter_;
    if (!h) { std::cerr << "error loading the model !!!\n"; return; }
    uahn_rng rng{seed_, rng_calls_++, nullptr};   // every forward draws fresh masks (model_to_trace.py:266-273)
    const bool want_err = cv_imshow && show_phtometric_error && (num_of_inference != 0 || !iteration);
    const auto t0 = std::chrono::steady_clock::now();
    double mean[8], cov[64];
    const int rc = uahn_infer(h, use_prior_4pt_offset ? prior_4pt_offset_vec : nullptr, &rng, mean, cov,
                              want_err ? err_map_ : nullptr);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc != UAHN_OK) { std::cerr << "network_inference: " << uahn_last_error(h) << std::endl; return; }
    memcpy(_pred_mean, mean, sizeof(mean));
    memcpy(_pred_Cov, cov, sizeof(cov));
    have_err_ = want_err;
    if (num_of_inference == 0) {   // running average after 100 inferences (HomographyNet.cpp:245-251)
      inference_counting++;
      if (inference_counting > 100) {
        sum_nn_time += ms;
        printf("[TIME]: %.3f (avg. = %.3f) milliseconds for pure network inference\n", ms,
               sum_nn_time / (double)(inference_counting - 100));
      }
    }
  }
  const double* pred_mean() const { return _pred_mean; }   // 8 doubles
  const double* pred_cov() const { return _pred_Cov; }     // 64 doubles, row-major
  const unsigned char* error_map() const { return have_err_ ? err_map_ : nullptr; }
  void set_seed(uint64_t s) { seed_ = s; }

 private:
  static uahn_handle* create(const std::string& path, int variant, bool show_err, int precision, int device,
                             bool quiet = false) {
    std::cout << "Loading the Network Model (UAHN flat weights) ..." << std::endl;
    uahn_config cfg{};
    cfg.weights_path = path.c_str();
    cfg.variant = variant;
    cfg.show_error = show_err;
    cfg.precision = precision;
    cfg.device = device;
    cfg.max_batch = 1;
    uahn_handle* h = nullptr;
    if (uahn_create(&cfg, &h) != UAHN_OK) {   // print-and-continue like HomographyNet.cpp:87-93
      if (!quiet) std::cerr << "error loading the model !!! " << uahn_last_error(nullptr) << "\n";
      return nullptr;
    }
    std::cerr << path << std::endl;
    return h;
  }

  uahn_handle* main_ = nullptr;
  uahn_handle* iter_ = nullptr;
  uint64_t inference_counting = 0;
  uint64_t rng_calls_ = 0;
  uint64_t seed_ = 0;
  int warmup_images_ = 0;
  bool cv_imshow = false, use_prior_4pt_offset = false, show_phtometric_error = false, iteration = false;
  bool have_err_ = false;
  double sum_nn_time = 0;
  double _latest_inference_time_stamp = -1.0;
  double _pred_mean[8] = {0};
  double _pred_Cov[64] = {0};
  unsigned char err_map_[UAHN_IMG_PIXELS];
};

}  // namespace pytorch

#endif  // UAHN_HOMOGRAPHYNET_H_
