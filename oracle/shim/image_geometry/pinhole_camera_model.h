// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for
// image_geometry/pinhole_camera_model.h. Only the zero-distortion, identity-
// rectification case is supported (calib/DVS-playroom.yaml is such a camera):
// rectifyPoint is the identity and projectPixelTo3dRay follows image_geometry:
//   ray = ((u - cx - Tx)/fx, (v - cy - Ty)/fy, 1) with fx=P[0], fy=P[5],
//   cx=P[2], cy=P[6], Tx=P[3], Ty=P[7].
// Used by /root/reference/src/utils/event_pano_warper.cpp:11,27-41.
#pragma once
#include <cstdlib>
#include <cstdio>
#include "opencv2/core.hpp"
#include "sensor_msgs/CameraInfo.h"
namespace image_geometry {
class PinholeCameraModel {
public:
  bool fromCameraInfo(const sensor_msgs::CameraInfo& msg) {
    info_ = msg;
    for (double d : msg.D) {
      if (d != 0.0) {
        std::fprintf(stderr, "oracle shim: distorted cameras are not supported\n");
        std::abort();
      }
    }
    return true;
  }
  cv::Size fullResolution() const { return cv::Size((int)info_.width, (int)info_.height); }
  cv::Point2d rectifyPoint(const cv::Point2d& uv_raw) const { return uv_raw; }
  cv::Point3d projectPixelTo3dRay(const cv::Point2d& uv_rect) const {
    const double fx = info_.P[0], fy = info_.P[5], cx = info_.P[2], cy = info_.P[6];
    const double Tx = info_.P[3], Ty = info_.P[7];
    cv::Point3d ray;
    ray.x = (uv_rect.x - cx - Tx) / fx;
    ray.y = (uv_rect.y - cy - Ty) / fy;
    ray.z = 1.0;
    return ray;
  }
private:
  sensor_msgs::CameraInfo info_;
};
}  // namespace image_geometry
