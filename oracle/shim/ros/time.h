// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <ros/time.h> so that the
// reference's hot-path translation units compile here without ROS.
// Only the surface used by /root/reference/src/emba/model.cpp,
// src/utils/trajectory.cpp and include/utils/trajectory.h is provided.
//
// Semantics follow rostime: Time/Duration are {sec, nsec}; Duration(double)
// uses sec=floor(d), nsec=round((d-sec)*1e9) then normalises;
// Duration*double goes through toSec(). rostime itself is not vendored in the
// reference, so this is an assumption (it only moves the batch mid-time of
// model.cpp:116-119 by <= 1 ns); see DESIGN.md "parity pinning".
#pragma once
#include <cstdint>
#include <cmath>
#include <ostream>

namespace ros {

class Duration {
public:
  int32_t sec = 0;
  int32_t nsec = 0;
  Duration() {}
  Duration(int32_t s, int32_t ns) { fromNSec(int64_t(s) * 1000000000LL + ns); }
  explicit Duration(double d) { fromSec(d); }
  Duration& fromSec(double d) {
    int64_t s = (int64_t)std::floor(d);
    int64_t ns = (int64_t)std::round((d - (double)s) * 1e9);
    return fromNSec(s * 1000000000LL + ns);
  }
  Duration& fromNSec(int64_t t) {
    int64_t s = t / 1000000000LL;
    int64_t ns = t % 1000000000LL;
    if (ns < 0) { ns += 1000000000LL; s -= 1; }
    sec = (int32_t)s; nsec = (int32_t)ns;
    return *this;
  }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  int64_t toNSec() const { return (int64_t)sec * 1000000000LL + (int64_t)nsec; }
  Duration operator*(double scale) const { return Duration(toSec() * scale); }
  Duration operator+(const Duration& o) const { Duration d; d.fromNSec(toNSec() + o.toNSec()); return d; }
  Duration operator-(const Duration& o) const { Duration d; d.fromNSec(toNSec() - o.toNSec()); return d; }
};

class Time {
public:
  uint32_t sec = 0;
  uint32_t nsec = 0;
  Time() {}
  Time(uint32_t s, uint32_t ns) : sec(s), nsec(ns) {}
  explicit Time(double t) { fromSec(t); }
  Time& fromSec(double t) {
    int64_t s = (int64_t)std::floor(t);
    int64_t ns = (int64_t)std::round((t - (double)s) * 1e9);
    return fromNSec((uint64_t)(s * 1000000000LL + ns));
  }
  Time& fromNSec(uint64_t t) {
    sec = (uint32_t)(t / 1000000000ULL);
    nsec = (uint32_t)(t % 1000000000ULL);
    return *this;
  }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  uint64_t toNSec() const { return (uint64_t)sec * 1000000000ULL + (uint64_t)nsec; }
  Duration operator-(const Time& o) const {
    Duration d; d.fromNSec((int64_t)toNSec() - (int64_t)o.toNSec()); return d;
  }
  Time operator+(const Duration& d) const { Time t; t.fromNSec((uint64_t)((int64_t)toNSec() + d.toNSec())); return t; }
  Time operator-(const Duration& d) const { Time t; t.fromNSec((uint64_t)((int64_t)toNSec() - d.toNSec())); return t; }
  bool operator<(const Time& o) const { return toNSec() < o.toNSec(); }
  bool operator>(const Time& o) const { return toNSec() > o.toNSec(); }
  bool operator<=(const Time& o) const { return toNSec() <= o.toNSec(); }
  bool operator>=(const Time& o) const { return toNSec() >= o.toNSec(); }
  bool operator==(const Time& o) const { return toNSec() == o.toNSec(); }
  bool operator!=(const Time& o) const { return toNSec() != o.toNSec(); }
};

inline std::ostream& operator<<(std::ostream& os, const Time& t) { return os << t.toSec(); }
inline std::ostream& operator<<(std::ostream& os, const Duration& d) { return os << d.toSec(); }

}  // namespace ros
