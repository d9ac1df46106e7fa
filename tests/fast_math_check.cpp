// Host-side accuracy check of gstpeaq_b200/csrc/peaq_math.cuh (compiled by tests/test_fast_math.py
// with -DPEAQ_MATH_HOST): the same sequence of IEEE operations as on the device, against long
// double libm.  Prints the maxima; the Python test asserts on them.
#include "peaq_math.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
int main() {
  std::mt19937_64 rng(12345);
  long double max_rel_exp = 0, max_rel_log = 0, max_abs_log = 0;
  double worst_e = 0, worst_l = 0;
  std::uniform_real_distribution<double> ue(-708.0, 708.0), us(-40, 40), ul(-300, 300), um(0.4, 2.6), un(-1e-3, 1e-3);
  for (int i = 0; i < (1 << 22); i++) {
    double x = (i & 1) ? ue(rng) : us(rng);
    long double want = expl((long double)x);
    long double rel = fabsl(((long double)peaq::peaq_exp(x) - want) / want);
    if (rel > max_rel_exp) { max_rel_exp = rel; worst_e = x; }
    double y;
    switch (i & 3) { case 0: y = pow(10.0, ul(rng)); break; case 1: y = um(rng); break; case 2: y = 1.0 + un(rng); break; default: y = pow(2.0, us(rng)) ; }
    long double wl = logl((long double)y);
    long double got = peaq::peaq_log(y);
    long double ab = fabsl(got - wl);
    long double rl = wl != 0 ? ab / fabsl(wl) : ab;
    if (rl > max_rel_log) { max_rel_log = rl; worst_l = y; }
    if (ab / fmaxl(1.0L, fabsl(wl)) > max_abs_log) max_abs_log = ab / fmaxl(1.0L, fabsl(wl));
  }
  // branch-free variants
  long double max_div = 0, max_sqrt = 0, max_logp = 0, max_expc = 0;
  std::uniform_real_distribution<double> ud(-200, 200);
  for (int i = 0; i < (1 << 22); i++) {
    const double a = (i & 1 ? -1.0 : 1.0) * pow(2.0, ud(rng)) * um(rng), b = (i & 2 ? -1.0 : 1.0) * pow(2.0, ud(rng)) * um(rng);
    long double want = (long double)a / (long double)b;
    long double rel = fabsl(((long double)peaq::peaq_div(a, b) - want) / want);
    if (rel > max_div) max_div = rel;
    const double x = pow(2.0, 2 * ud(rng)) * um(rng);
    want = sqrtl((long double)x);
    rel = fabsl(((long double)peaq::peaq_sqrt(x) - want) / want);
    if (rel > max_sqrt) max_sqrt = rel;
    double y;
    switch (i & 3) { case 0: y = pow(10.0, ul(rng)); break; case 1: y = um(rng); break; case 2: y = 1.0 + un(rng); break; default: y = pow(2.0, us(rng)); }
    long double wl = logl((long double)y);
    long double ab = fabsl((long double)peaq::peaq_log_pos(y) - wl);
    long double rl = wl != 0 ? ab / fabsl(wl) : ab;
    if (rl > max_logp) max_logp = rl;
    const double z = (i & 1) ? ue(rng) : us(rng);
    want = expl((long double)z);
    rel = fabsl(((long double)peaq::peaq_exp_clamped(z) - want) / want);
    if (rel > max_expc) max_expc = rel;
  }
  printf("div max rel %.3Le\nsqrt max rel %.3Le\nlog_pos max rel %.3Le\nexp_clamped max rel %.3Le\n", max_div, max_sqrt, max_logp, max_expc);
  printf("sqrt(0)=%g exp_clamped(-1e6)=%g exp_clamped(nan)=%g div(1,3)=%.17g sqrt(2)=%.17g\n", peaq::peaq_sqrt(0.0),
         peaq::peaq_exp_clamped(-1e6), peaq::peaq_exp_clamped(NAN), peaq::peaq_div(1.0, 3.0), peaq::peaq_sqrt(2.0));
  printf("exp max rel %.3Le at %.17g\nlog max rel %.3Le at %.17g  (abs/max(1,|ln|) %.3Le)\n", max_rel_exp, worst_e, max_rel_log, worst_l, max_abs_log);
  // special values go through the library
  double sp[] = {0.0, -1.0, INFINITY, NAN, 1e-310, 1.0};
  for (double v : sp) printf("log(%g)=%g exp(%g)=%g\n", v, peaq::peaq_log(v), v, peaq::peaq_exp(v));
  printf("exp(-745)=%g exp(709.5)=%g exp(-1e6)=%g log10(1000)=%.17g\n", peaq::peaq_exp(-745), peaq::peaq_exp(709.5), peaq::peaq_exp(-1e6), peaq::peaq_log10(1000.0));
  return 0;
}
