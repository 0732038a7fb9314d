// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for sensor_msgs/CameraInfo.h
#pragma once
#include <array>
#include <string>
#include <vector>
#include <cstdint>
namespace sensor_msgs {
struct CameraInfo {
  uint32_t height = 0, width = 0;
  std::string distortion_model;
  std::vector<double> D;
  std::array<double, 9> K{};
  std::array<double, 9> R{};
  std::array<double, 12> P{};
};
}  // namespace sensor_msgs
