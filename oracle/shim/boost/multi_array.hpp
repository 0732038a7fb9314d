// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for boost::multi_array<T, N> as used by the reference's
// Poisson solver (/root/reference/src/image_rec/laplace.cpp, poisson_reconstruction.cpp): construction from
// boost::extents[a][b], value-initialised storage, operator[][] access, shape(), resize(extents) that keeps the
// overlapping elements, deep copy. Written from the Boost.MultiArray documentation; no Boost source is used.
// Boost is not installed in this image.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>  // the real header provides assert() transitively; laplace.cpp relies on that
#include <cstddef>
#include <vector>

namespace boost {

namespace detail_shim {
template <std::size_t N> struct extent_gen {
  std::array<std::size_t, N> ext;
  extent_gen<N + 1> operator[](std::size_t e) const {
    extent_gen<N + 1> r;
    for (std::size_t i = 0; i < N; i++) r.ext[i] = ext[i];
    r.ext[N] = e;
    return r;
  }
};
template <> struct extent_gen<0> {
  extent_gen<1> operator[](std::size_t e) const {
    extent_gen<1> r;
    r.ext[0] = e;
    return r;
  }
};
}  // namespace detail_shim

static const detail_shim::extent_gen<0> extents = detail_shim::extent_gen<0>();

template <typename T, std::size_t N> class multi_array;

template <typename T> class multi_array<T, 2> {
public:
  typedef std::size_t size_type;
  typedef std::ptrdiff_t index;
  multi_array() { shape_[0] = shape_[1] = 0; }
  explicit multi_array(const detail_shim::extent_gen<2>& e) {
    shape_[0] = e.ext[0];
    shape_[1] = e.ext[1];
    data_.assign(shape_[0] * shape_[1], T());
  }
  multi_array(const multi_array& o) : data_(o.data_) { shape_[0] = o.shape_[0]; shape_[1] = o.shape_[1]; }
  multi_array& operator=(const multi_array& o) {
    data_ = o.data_;
    shape_[0] = o.shape_[0];
    shape_[1] = o.shape_[1];
    return *this;
  }
  const size_type* shape() const { return shape_; }
  size_type num_elements() const { return data_.size(); }
  T* data() { return data_.data(); }
  const T* data() const { return data_.data(); }
  T* operator[](index i) { return data_.data() + (size_type)i * shape_[1]; }
  const T* operator[](index i) const { return data_.data() + (size_type)i * shape_[1]; }
  // Boost semantics: elements whose indices exist in both the old and the new shape keep their value
  multi_array& resize(const detail_shim::extent_gen<2>& e) {
    std::vector<T> nd(e.ext[0] * e.ext[1], T());
    const size_type r = std::min(shape_[0], e.ext[0]), c = std::min(shape_[1], e.ext[1]);
    for (size_type i = 0; i < r; i++)
      for (size_type j = 0; j < c; j++) nd[i * e.ext[1] + j] = data_[i * shape_[1] + j];
    data_.swap(nd);
    shape_[0] = e.ext[0];
    shape_[1] = e.ext[1];
    return *this;
  }

private:
  std::vector<T> data_;
  size_type shape_[2];
};

}  // namespace boost
