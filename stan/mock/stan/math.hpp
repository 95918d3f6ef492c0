// MOCK of the few Stan Math / Boost names stan/gp_lml_stan.hpp uses, for `g++ -fsyntax-only` in an image
// without Stan (tests/test_abi.py).  Never shipped to a Stan build.
#pragma once
#include <ostream>
#include <vector>
#include <type_traits>
namespace Eigen {
constexpr int Dynamic = -1;
template <typename T, int R, int C> struct Matrix { const T *data() const; int size() const; };
typedef Matrix<double, Dynamic, 1> VectorXd;
}  // namespace Eigen
namespace stan {
namespace math {
struct var { double val() const; };
inline double value_of(double x) { return x; }
inline double value_of(const var &v) { return v.val(); }
var precomputed_gradients(double value, const std::vector<var> &operands, const std::vector<double> &gradients);
}  // namespace math
template <typename T> struct is_constant { static constexpr bool value = !std::is_same<T, math::var>::value; };
}  // namespace stan
template <typename... T> struct any_var : std::false_type {};
template <typename H, typename... T> struct any_var<H, T...> : std::integral_constant<bool, std::is_same<H, stan::math::var>::value || any_var<T...>::value> {};
namespace boost { namespace math { namespace tools {
template <typename... T> struct promote_args { typedef typename std::conditional<any_var<T...>::value, stan::math::var, double>::type type; };
}}}  // namespace boost::math::tools
