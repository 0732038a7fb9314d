// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for dvs_msgs/EventArray.h
#pragma once
#include <vector>
#include "dvs_msgs/Event.h"
namespace dvs_msgs {
struct EventArray {
  uint32_t height = 0, width = 0;
  std::vector<Event> events;
};
}  // namespace dvs_msgs
