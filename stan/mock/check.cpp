// type-checks stan/gp_lml_stan.hpp against the mock: both overloads must instantiate
#include "stan/math.hpp"
#include "../gp_lml_stan.hpp"
using stan::math::var;
var f1(const std::vector<double> &x, const Eigen::VectorXd &y, var a, var r, var s) { return gp_lml(x, y, a, r, s, nullptr); }
var f2(const std::vector<double> &x, const Eigen::VectorXd &y, double a, var r, double s) { return gp_lml(x, y, a, r, s, nullptr); }
double f3(const std::vector<double> &x, const Eigen::VectorXd &y) { return gp_lml(x, y, 1.0, 1.0, 0.3, nullptr); }
var f4(const Eigen::VectorXd &t, const Eigen::VectorXd &dx, var a, var l2, var s2) { return gp_lml_dd(t, dx, a, l2, s2, nullptr); }
var f5(const Eigen::VectorXd &t, const Eigen::VectorXd &dx, var a, double l2, var s2) { return gp_lml_dd(t, dx, a, l2, s2, nullptr); }
double f6(const Eigen::VectorXd &t, const Eigen::VectorXd &dx) { return gp_lml_dd(t, dx, 1.0, 2.0, 0.04, nullptr); }
typedef Eigen::Matrix<var, Eigen::Dynamic, 1> VectorXv;
var f7(const Eigen::VectorXd &t, const Eigen::VectorXd &y, var a, var r, const VectorXv &nz) { return gp_lml_joint(t, y, a, r, nz, nullptr); }
var f8(const Eigen::VectorXd &t, const Eigen::VectorXd &y, double a, var r, const Eigen::VectorXd &nz) { return gp_lml_joint(t, y, a, r, nz, nullptr); }
double f9(const Eigen::VectorXd &t, const Eigen::VectorXd &y, const Eigen::VectorXd &nz) { return gp_lml_joint(t, y, 1.0, 1.0, nz, nullptr); }
