// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the OpenCV C++ API
// surface that the reference's hot-path translation units use
// (/root/reference/src/emba/model.cpp, src/utils/event_pano_warper.cpp,
// src/utils/trajectory.cpp and the headers they include). Written from the
// OpenCV documentation of each call; no OpenCV source is used. OpenCV C++
// headers are not installed in this image (python cv2 is, and is used by
// tests/test_oracle_ref.py to validate Sobel below bit-for-bit up to summation
// order).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_64FC2 CV_MAKETYPE(CV_64F, 2)

namespace cv {

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<double> Point3d;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

template <typename T, int n> struct Vec {
  T val[n];
  Vec() { for (int i = 0; i < n; i++) val[i] = T(0); }
  Vec(T a, T b, T c) { static_assert(n == 3, "Vec3 ctor"); val[0] = a; val[1] = b; val[2] = c; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<int, 3> Vec3i;
typedef Vec<double, 2> Vec2d;

enum MarkerTypes { MARKER_CROSS = 0, MARKER_TILTED_CROSS = 1, MARKER_STAR = 2, MARKER_DIAMOND = 3,
                   MARKER_SQUARE = 4, MARKER_TRIANGLE_UP = 5, MARKER_TRIANGLE_DOWN = 6 };
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_MINMAX = 32 };
enum ColormapTypes { COLORMAP_JET = 2 };
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2,
                   BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };

inline size_t shimElemSize(int type) {
  const int depth = type & 7;
  const int cn = (type >> CV_CN_SHIFT) + 1;
  size_t d = 1;
  switch (depth) {
    case CV_8U: d = 1; break;
    case CV_32S: d = 4; break;
    case CV_32F: d = 4; break;
    case CV_64F: d = 8; break;
    default: std::abort();
  }
  return d * (size_t)cn;
}

// Reference-counted dense 2-D matrix (header/data split like cv::Mat: copying a
// Mat shares the buffer; clone()/copyTo() make deep copies).
class Mat {
public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;
    rows = r; cols = c; type_ = type;
    const size_t bytes = (size_t)r * (size_t)c * shimElemSize(type);
    buf_ = std::shared_ptr<std::vector<unsigned char>>(new std::vector<unsigned char>(bytes));
    data = buf_->data();
  }
  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    if (m.data) std::memset(m.data, 0, m.bytes());
    return m;
  }
  static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }

  int type() const { return type_; }
  Size size() const { return Size(cols, rows); }
  size_t total() const { return (size_t)rows * (size_t)cols; }
  bool empty() const { return data == nullptr || total() == 0; }
  size_t bytes() const { return total() * shimElemSize(type_); }

  template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(data)[(size_t)r * cols + c]; }
  template <typename T> const T& at(int r, int c) const {
    return reinterpret_cast<const T*>(data)[(size_t)r * cols + c];
  }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data) + (size_t)r * cols; }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data) + (size_t)r * cols; }
  template <typename T> T& at(Point2i p) { return at<T>(p.y, p.x); }
  template <typename T> const T& at(Point2i p) const { return at<T>(p.y, p.x); }

  Mat clone() const {
    Mat m;
    if (data) { m.create(rows, cols, type_); std::memcpy(m.data, data, bytes()); }
    return m;
  }
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    if (data) std::memcpy(dst.data, data, bytes());
  }
  Mat& setTo(double v) {
    if (v != 0.0) std::abort();  // only setTo(0) is used by the reference hot path
    if (data) std::memset(data, 0, bytes());
    return *this;
  }

private:
  int type_ = 0;
  std::shared_ptr<std::vector<unsigned char>> buf_;
};

inline Mat operator*(double s, const Mat& a) {
  if (a.type() != CV_64FC1) std::abort();
  Mat r(a.rows, a.cols, a.type());
  const double* pa = reinterpret_cast<const double*>(a.data);
  double* pr = reinterpret_cast<double*>(r.data);
  for (size_t i = 0; i < a.total(); i++) pr[i] = s * pa[i];
  return r;
}
inline Mat operator*(const Mat& a, double s) { return s * a; }
inline Mat operator+(const Mat& a, const Mat& b) {
  if (a.type() != CV_64FC1 || b.type() != CV_64FC1 || a.rows != b.rows || a.cols != b.cols) std::abort();
  Mat r(a.rows, a.cols, a.type());
  const double* pa = reinterpret_cast<const double*>(a.data);
  const double* pb = reinterpret_cast<const double*>(b.data);
  double* pr = reinterpret_cast<double*>(r.data);
  for (size_t i = 0; i < a.total(); i++) pr[i] = pa[i] + pb[i];
  return r;
}

// Small fixed-size matrix, row-major.
template <typename T, int m, int n> struct Matx {
  T val[m * n];
  Matx() { for (int i = 0; i < m * n; i++) val[i] = T(0); }
  Matx(T v0, T v1, T v2, T v3, T v4, T v5, T v6, T v7, T v8) {
    static_assert(m * n == 9, "9-arg Matx ctor");
    val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; val[4] = v4;
    val[5] = v5; val[6] = v6; val[7] = v7; val[8] = v8;
  }
  T& operator()(int i, int j) { return val[i * n + j]; }
  const T& operator()(int i, int j) const { return val[i * n + j]; }
};
typedef Matx<double, 2, 3> Matx23d;
typedef Matx<double, 3, 3> Matx33d;

template <typename T, int m, int l, int n>
inline Matx<T, m, n> operator*(const Matx<T, m, l>& a, const Matx<T, l, n>& b) {
  Matx<T, m, n> c;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      T s = 0;
      for (int k = 0; k < l; k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}

// Matx (m x l) times a CV_64FC1 Mat (l x n) -> Mat (m x n), as OpenCV's
// MatExpr operator*(const Matx&, const Mat&) evaluates to.
template <int m, int l>
inline Mat operator*(const Matx<double, m, l>& a, const Mat& b) {
  if (b.type() != CV_64FC1 || b.rows != l) std::abort();
  Mat c(m, b.cols, CV_64FC1);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < b.cols; j++) {
      double s = 0;
      for (int k = 0; k < l; k++) s += a(i, k) * b.at<double>(k, j);
      c.at<double>(i, j) = s;
    }
  return c;
}

inline int shimReflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

// cv::Sobel(src, dst, ddepth, dx, dy) with the documented defaults: ksize=3,
// scale=1, delta=0, BORDER_DEFAULT (= BORDER_REFLECT_101). For ksize=3 the
// kernel is separable: derivative [-1 0 1] along the differentiated axis and
// smoothing [1 2 1] along the other one. Only first derivatives on CV_64FC1
// are supported (all the reference uses, model.cpp:89-92).
inline void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize = 3,
                  double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT) {
  if (src.type() != CV_64FC1 || (ddepth != CV_64F && ddepth != CV_64FC1) || ksize != 3 ||
      borderType != BORDER_REFLECT_101 || !((dx == 1 && dy == 0) || (dx == 0 && dy == 1)))
    std::abort();
  const int R = src.rows, C = src.cols;
  Mat out(R, C, CV_64FC1);
  const double kd[3] = {-1.0, 0.0, 1.0};
  const double ks[3] = {1.0, 2.0, 1.0};
  const double* kx = (dx == 1) ? kd : ks;
  const double* ky = (dy == 1) ? kd : ks;
  // row pass then column pass (OpenCV's separable filter order)
  std::vector<double> tmp((size_t)R * C);
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++) {
      double s = 0;
      for (int k = -1; k <= 1; k++) s += kx[k + 1] * src.at<double>(r, shimReflect101(c + k, C));
      tmp[(size_t)r * C + c] = s;
    }
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++) {
      double s = 0;
      for (int k = -1; k <= 1; k++) s += ky[k + 1] * tmp[(size_t)shimReflect101(r + k, R) * C + c];
      out.at<double>(r, c) = s * scale + delta;
    }
  dst = out;
}

// Drawing / display helpers referenced by headers on the include path but never
// reached on the hot path: no-ops.
template <typename P>
inline void drawMarker(Mat&, const P&, const Vec3i&, int = 0, int = 20, int = 1, int = 8) {}
inline void normalize(const Mat&, Mat&, double = 1, double = 0, int = NORM_L2, int = -1) {}
inline void applyColorMap(const Mat&, Mat&, int) {}

}  // namespace cv
