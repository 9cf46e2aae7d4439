// Minimal stand-in for the reference's caffe/util/math_functions.hpp so that
// /root/reference/src/caffe/util/im2col.cpp compiles VERBATIM into oracle/_ref
// (it needs only caffe_set, TypedConsts, DCHECK_LT and std::vector).
// Test infrastructure only; not part of the product.
#ifndef B2O_REF_SHIM_MATH_FUNCTIONS_HPP_
#define B2O_REF_SHIM_MATH_FUNCTIONS_HPP_
#include <cassert>
#include <cstring>
#include <vector>
namespace caffe {
using std::vector;
template <typename Dtype> struct TypedConsts { static const Dtype zero, one; };
template <typename Dtype> const Dtype TypedConsts<Dtype>::zero = Dtype(0);
template <typename Dtype> const Dtype TypedConsts<Dtype>::one = Dtype(1);
template <typename Dtype>
inline void caffe_set(const size_t n, const Dtype alpha, Dtype* y) {
  for (size_t i = 0; i < n; ++i) y[i] = alpha;
}
}  // namespace caffe
#define DCHECK_LT(a, b) assert((a) < (b))
#endif
