// Functional MOCK of the few Stan Math / Eigen / Boost names stan/gp_lml_stan.hpp uses (test infrastructure).
// Stan Math and Eigen are absent from this image; with this mock the header is not only type-checked
// (tests/test_abi.py) but compiled, linked against libgpb200.so and EXECUTED on the GPU box
// (stan/mock/run.cpp, tests/test_stan_header_gpu.py): `var` is a tape node that records its operands and
// partials, precomputed_gradients builds such a node, grad() runs the reverse sweep -- the mechanism the
// reference's own plug-in relies on (models/cubic_interpolated_gp.hpp:28, precomp_v_vari).
#pragma once
#include <ostream>
#include <type_traits>
#include <vector>
namespace Eigen {
constexpr int Dynamic = -1;
template <typename T, int R, int C>
struct Matrix {
  std::vector<T> v_;
  Matrix() {}
  explicit Matrix(int n) : v_(n) {}
  const T *data() const { return v_.data(); }
  T *data() { return v_.data(); }
  int size() const { return (int)v_.size(); }
  T &operator[](int i) { return v_[i]; }
  const T &operator[](int i) const { return v_[i]; }
};
typedef Matrix<double, Dynamic, 1> VectorXd;
}  // namespace Eigen
namespace stan {
namespace math {
struct vari {
  double val_, adj_;
  std::vector<vari *> operands_;
  std::vector<double> partials_;
  explicit vari(double v) : val_(v), adj_(0.0) {}
};
inline std::vector<vari *> &tape() {
  static std::vector<vari *> t;
  return t;
}
struct var {
  vari *vi_;
  var() : vi_(nullptr) {}
  var(double v) : vi_(new vari(v)) { tape().push_back(vi_); }
  explicit var(vari *vi) : vi_(vi) {}
  double val() const { return vi_->val_; }
  double adj() const { return vi_->adj_; }
};
inline double value_of(double x) { return x; }
inline double value_of(const var &v) { return v.val(); }
inline var precomputed_gradients(double value, const std::vector<var> &operands, const std::vector<double> &gradients) {
  vari *vi = new vari(value);
  for (size_t i = 0; i < operands.size(); i++) {
    vi->operands_.push_back(operands[i].vi_);
    vi->partials_.push_back(gradients[i]);
  }
  tape().push_back(vi);
  return var(vi);
}
// reverse sweep from v over everything recorded so far
inline void grad(const var &v) {
  for (vari *n : tape()) n->adj_ = 0.0;
  v.vi_->adj_ = 1.0;
  for (size_t i = tape().size(); i-- > 0;) {
    vari *n = tape()[i];
    for (size_t k = 0; k < n->operands_.size(); k++) n->operands_[k]->adj_ += n->adj_ * n->partials_[k];
  }
}
}  // namespace math
template <typename T> struct is_constant { static constexpr bool value = !std::is_same<T, math::var>::value; };
}  // namespace stan
template <typename... T> struct any_var : std::false_type {};
template <typename H, typename... T> struct any_var<H, T...> : std::integral_constant<bool, std::is_same<H, stan::math::var>::value || any_var<T...>::value> {};
namespace boost { namespace math { namespace tools {
template <typename... T> struct promote_args { typedef typename std::conditional<any_var<T...>::value, stan::math::var, double>::type type; };
}}}  // namespace boost::math::tools
