// Executes stan/gp_lml_stan.hpp against the functional mock of Stan Math (stan/mock/stan/math.hpp) and the real
// libgpb200.so: every overload a Stan model would instantiate is called, the reverse sweep is run, and values and
// adjoints are printed as one JSON object for tests/test_stan_header_gpu.py to compare with the oracle.
// Inputs are deterministic closed forms so that the test can rebuild them.
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include "stan/math.hpp"
#include "../gp_lml_stan.hpp"
using stan::math::var;
typedef Eigen::Matrix<var, Eigen::Dynamic, 1> VectorXv;

static void print3(const char *k, double v, double a, double b, double c, bool last = false) {
  std::printf("\"%s\": {\"value\": %.17g, \"adj\": [%.17g, %.17g, %.17g]}%s\n", k, v, a, b, c, last ? "" : ",");
}

int main() {
  const int n = 200;
  std::vector<double> x(n);
  Eigen::VectorXd y(n), t(n), dx(n), ystack(3 * n), nz(3);
  for (int i = 0; i < n; i++) {
    x[i] = 0.05 * i + 0.01 * std::sin(1.7 * i);
    y[i] = std::sin(x[i]) + 0.3 * std::cos(5.0 * x[i]);
    t[i] = 10.0 * i / (n - 1);
    dx[i] = std::cos(t[i]) + 0.05 * std::sin(11.0 * t[i]);
    ystack[i] = std::sin(t[i]) + 0.05 * std::cos(7.0 * t[i]);
    ystack[n + i] = std::cos(t[i]) + 0.05 * std::sin(9.0 * t[i]);
    ystack[2 * n + i] = -std::sin(t[i]) + 0.05 * std::cos(13.0 * t[i]);
  }
  nz[0] = 0.2; nz[1] = 0.25; nz[2] = 0.3;
  std::printf("{\n");
  try {
    {  // all three hyper-parameters are parameters (models/fit_hyperparameters.stan:12-16)
      var a(1.1), r(0.9), s(0.3);
      var lp = gp_lml(x, y, a, r, s, nullptr);
      stan::math::grad(lp);
      print3("lml_vvv", lp.val(), a.adj(), r.adj(), s.adj());
    }
    {  // mixed: only rho is a parameter
      var r(0.9);
      var lp = gp_lml(x, y, 1.1, r, 0.3, nullptr);
      stan::math::grad(lp);
      print3("lml_dvd", lp.val(), 0.0, r.adj(), 0.0);
    }
    print3("lml_ddd", gp_lml(x, y, 1.1, 0.9, 0.3, nullptr), 0, 0, 0);
    {  // gpderivs.py:62-83 parametrisation (sf2, l2, s2)
      var sf2(1.3), l2(1.7), s2(0.04);
      var lp = gp_lml_dd(t, dx, sf2, l2, s2, nullptr);
      stan::math::grad(lp);
      print3("dd_vvv", lp.val(), sf2.adj(), l2.adj(), s2.adj());
    }
    print3("dd_ddd", gp_lml_dd(t, dx, 1.3, 1.7, 0.04, nullptr), 0, 0, 0);
    {  // joint (y, y', y'') with a var noise vector
      var a(1.2), r(1.1);
      VectorXv nzv(3);
      for (int b = 0; b < 3; b++) nzv[b] = var(nz[b]);
      var lp = gp_lml_joint(t, ystack, a, r, nzv, nullptr);
      stan::math::grad(lp);
      std::printf("\"joint_vvv\": {\"value\": %.17g, \"adj\": [%.17g, %.17g, %.17g, %.17g, %.17g]},\n", lp.val(), a.adj(), r.adj(),
                  nzv[0].adj(), nzv[1].adj(), nzv[2].adj());
    }
    print3("joint_ddd", gp_lml_joint(t, ystack, 1.2, 1.1, nz, nullptr), 0, 0, 0);
    // a non-positive-definite covariance must surface as std::domain_error (NUTS then rejects the proposal)
    bool threw = false;
    try {
      std::vector<double> xd(x);
      xd[7] = xd[6];  // duplicated input, no noise, no jitter: exactly singular
      (void)gp_lml(xd, y, 1.0, 1.0, 0.0, nullptr);
    } catch (const std::domain_error &e) {
      threw = true;
    }
    std::printf("\"domain_error_on_singular\": %s\n", threw ? "true" : "false");
  } catch (const std::exception &e) {
    std::printf("\"error\": \"%s\"\n}\n", e.what());
    return 1;
  }
  std::printf("}\n");
  return 0;
}
