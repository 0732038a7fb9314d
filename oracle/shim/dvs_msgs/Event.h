// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for dvs_msgs/Event.h
// (fields as in the dvs_msgs ROS message: uint16 x, uint16 y, time ts, bool polarity).
#pragma once
#include <cstdint>
#include "ros/time.h"
namespace dvs_msgs {
struct Event {
  uint16_t x = 0;
  uint16_t y = 0;
  ros::Time ts;
  uint8_t polarity = 0;
};
}  // namespace dvs_msgs
