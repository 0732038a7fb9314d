// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <glog/logging.h>.
// CHECK_* abort with a message (as glog does); LOG/VLOG swallow their stream.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>
namespace oracle_shim {
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  std::ostringstream ss;
  const char* what;
  FatalStream(const char* w) : what(w) {}
  template <typename T> FatalStream& operator<<(const T& v) { ss << v; return *this; }
  ~FatalStream() { std::fprintf(stderr, "CHECK failed: %s %s\n", what, ss.str().c_str()); std::abort(); }
};
}  // namespace oracle_shim
#define ORACLE_SHIM_CHECK_OP(a, op, b) \
  if ((a) op (b)) {} else ::oracle_shim::FatalStream(#a " " #op " " #b)
#define CHECK(c) if (c) {} else ::oracle_shim::FatalStream(#c)
#define CHECK_EQ(a, b) ORACLE_SHIM_CHECK_OP(a, ==, b)
#define CHECK_NE(a, b) ORACLE_SHIM_CHECK_OP(a, !=, b)
#define CHECK_GE(a, b) ORACLE_SHIM_CHECK_OP(a, >=, b)
#define CHECK_GT(a, b) ORACLE_SHIM_CHECK_OP(a, >, b)
#define CHECK_LE(a, b) ORACLE_SHIM_CHECK_OP(a, <=, b)
#define CHECK_LT(a, b) ORACLE_SHIM_CHECK_OP(a, <, b)
#define LOG(sev) ::oracle_shim::NullStream()
#define VLOG(n) ::oracle_shim::NullStream()
