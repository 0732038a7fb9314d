// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <opencv2/core/eigen.hpp>
// (cv::cv2eigen for CV_64FC1 sources only; used at model.cpp:157).
#pragma once
#include <Eigen/Core>
#include "opencv2/core.hpp"
namespace cv {
template <typename T, int R, int C, int O, int MR, int MC>
inline void cv2eigen(const Mat& src, Eigen::Matrix<T, R, C, O, MR, MC>& dst) {
  if (src.type() != CV_64FC1) std::abort();
  if (R != Eigen::Dynamic && (src.rows != R)) std::abort();
  if (C != Eigen::Dynamic && (src.cols != C)) std::abort();
  dst.resize(src.rows, src.cols);
  for (int r = 0; r < src.rows; r++)
    for (int c = 0; c < src.cols; c++) dst(r, c) = (T)src.at<double>(r, c);
}
}  // namespace cv
