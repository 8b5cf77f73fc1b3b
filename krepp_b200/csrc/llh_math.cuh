// llh_math.cuh -- the floating-point core of K4/K5: optimize::HDistHistLLH and the Brent minimiser, written so that the
// same source compiles for the device (solve.cu, with --fmad=false) and for the host (tests/test_llh_math_cpu.py builds
// it with g++ -ffp-contract=off and checks it against the CPU restatement of the reference).
//
//   objective  : optimize::HDistHistLLH::operator()          ref src/hdhistllh.hpp:71-89
//   tables     : HDistHistLLH ctor                            ref src/hdhistllh.hpp:51-69
//   minimiser  : boost::math::tools::brent_find_minima(f, 1e-10, 0.5, 16)
//                ref external/boost/libs/math/include/boost/math/tools/minima.hpp:23-138, called at src/query.cpp:430
//
// Rounding contract.  The reference runs on x86-64 without FMA contraction, and Brent's parabolic-vs-golden decisions
// (tolerance only 2^-15) must follow the same iteration sequence, so every product and sum
// of the objective rounds separately and each accumulation chain keeps the reference's order:
//   sum   : x = 0..th      sum  -= (k ln(1-d) + x (ln d - ln(1-d))) * hist[x]
//   lv_m  : x = 0..k       lv_m += w[x] * powdc ; powdc *= d/(1-d)          (w = hnk for x <= th, C(k,x) above)
// The two chains never feed each other inside the loop, so they are run as two loops (same values, same order within
// each chain).  The only libm calls are log() and pow((1-d), k): pow with an integer exponent is formed by a
// double-double square-and-multiply chain and rounded once, i.e. the correctly rounded value, which is what glibc's
// pow returns except in near-tie cases (its error bound is 0.52 ulp).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define KLLH_FN __host__ __device__ __forceinline__
#else
#define KLLH_FN inline
#endif
#ifdef __CUDA_ARCH__
#define KLLH_UNROLL _Pragma("unroll")
#else
#define KLLH_UNROLL
#endif

namespace krepp {

constexpr int kLlhMaxK = 32;

// Binomial tables of optimize::HDistHistLLH (ref src/hdhistllh.hpp:51-69), exact in double (values < 2^53).
struct LlhTables {
  double w[kLlhMaxK + 1]; // x <= th: C(k, x) - C(k-h, x) (0 for x = 0); th < x <= k: C(k, x)
};

inline void llh_tables(LlhTables& t, uint32_t k, uint32_t h, uint32_t th)
{
  uint64_t ck[kLlhMaxK + 1] = {0}, vc = 1;
  ck[0] = 1;
  for (uint32_t i = 0; i < k; ++i) ck[i + 1] = (ck[i] * (k - i)) / (i + 1);
  for (uint32_t i = 0; i <= (uint32_t)kLlhMaxK; ++i) t.w[i] = i <= k ? (double)ck[i] : 0.0;
  t.w[0] = 0;
  const uint32_t nh = k - h;
  for (uint32_t i = 1; i <= th && i <= k; ++i) { vc = (vc * (nh - i + 1)) / i; t.w[i] = (double)(ck[i] - vc); }
}

// x^n (n >= 1) from a double-double product chain, rounded once.
KLLH_FN double powi_rounded(double x, uint32_t n)
{
  if (n == 0) return 1.0;
  int top = 31;
  while (!((n >> top) & 1u)) --top;
  double hi = x, lo = 0.0;
  for (int b = top - 1; b >= 0; --b) {
    double p = hi * hi;
    double e = fma(hi, hi, -p);
    e = fma(hi + hi, lo, e);
    double s = p + e;
    lo = e - (s - p);
    hi = s;
    if ((n >> b) & 1u) {
      p = hi * x;
      e = fma(hi, x, -p);
      e = fma(lo, x, e);
      s = p + e;
      lo = e - (s - p);
      hi = s;
    }
  }
  return hi;
}

// The part of the objective that depends on d alone.
struct DTerms { double logdn_k, logdp, lv_m; };

KLLH_FN DTerms d_terms(const LlhTables& t, double d, uint32_t k)
{
  DTerms s;
  const double omd = 1.0 - d;
  double powdc = powi_rounded(omd, k);
  double logdn = log(omd);
  s.logdp = log(d) - logdn;
  s.logdn_k = logdn * (double)k;
  const double dratio = d / omd;
  double lv_m = 0.0;
KLLH_UNROLL
  for (int x = 0; x <= kLlhMaxK; ++x) {
    if ((uint32_t)x > k) break;
    lv_m += t.w[x] * powdc;
    powdc *= dratio;
  }
  s.lv_m = lv_m;
  return s;
}

// N = number of histogram bins held (th + 1 <= N).
template <int N>
struct Objective {
  double mc[N];
  double uc, rho;
  uint32_t k, th;
  KLLH_FN double finish(const DTerms& s) const
  {
    double sum = 0.0;
KLLH_UNROLL
    for (int x = 0; x < N; ++x)
      if ((uint32_t)x <= th) sum -= (s.logdn_k + (double)x * s.logdp) * mc[x];
    return sum - log(rho * s.lv_m + 1.0 - rho) * uc;
  }
  KLLH_FN double eval(const LlhTables& t, double d) const { return finish(d_terms(t, d, k)); }
};

// The abscissae every minimisation visits first: Brent starts at max = 0.5, and its first two steps are golden-section
// steps whatever the function is (delta2 = 0 at the first; p = q = 0 at the second because two of the three points still
// coincide), so evaluation 0 is at u[0], evaluation 1 at u[1], evaluation 2 at u[2] or u[3].  The d-only terms at these
// four points are computed once per CTA (solve.cu) and looked up BY VALUE, so nothing depends on this being exhaustive.
KLLH_FN void brent_first_points(double (&u)[4])
{
  const double lo = 1e-10, hi = 0.5;
  const double golden = (double)0.3819660f;
  u[0] = hi;
  u[1] = hi + golden * (lo - hi);
  u[2] = u[1] + golden * (lo - u[1]);      // after f(u1) <= f(0.5): bracket [lo, 0.5], x = u1 >= mid
  u[3] = hi + golden * (u[1] - hi);        // after f(u1) >  f(0.5): bracket [u1, 0.5], x = 0.5 >= mid
}

// boost::math::tools::brent_find_minima(f, 1e-10, 0.5, 16) with unlimited iterations; f(u, i) = objective at u, i = index
// of the evaluation (0, 1, 2, ...).
template <class F>
KLLH_FN void brent_minimum(const F& f, double& xo, double& fo)
{
  double min = 1e-10, max = 0.5;
  const double tolerance = 3.0517578125e-05; // ldexp(1.0, 1 - 16)
  const double golden = (double)0.3819660f;
  double x, w, v, u, delta, delta2, fu, fv, fw, fx, mid, fract1, fract2;
  x = w = v = max;
  fw = fv = fx = f(x, 0);
  delta2 = delta = 0;
  for (int it = 1;; ++it) {
    mid = (min + max) / 2;
    fract1 = tolerance * fabs(x) + tolerance / 4;
    fract2 = 2 * fract1;
    if (fabs(x - mid) <= (fract2 - (max - min) / 2)) break;
    if (fabs(delta2) > fract1) {
      double r = (x - w) * (fx - fv);
      double q = (x - v) * (fx - fw);
      double p = (x - v) * q - (x - w) * r;
      q = 2 * (q - r);
      if (q > 0) p = -p;
      q = fabs(q);
      const double td = delta2;
      delta2 = delta;
      if ((fabs(p) >= fabs(q * td / 2)) || (p <= q * (min - x)) || (p >= q * (max - x))) {
        delta2 = (x >= mid) ? min - x : max - x;
        delta = golden * delta2;
      } else {
        delta = p / q;
        u = x + delta;
        if (((u - min) < fract2) || ((max - u) < fract2)) delta = (mid - x) < 0 ? -fabs(fract1) : fabs(fract1);
      }
    } else {
      delta2 = (x >= mid) ? min - x : max - x;
      delta = golden * delta2;
    }
    u = (fabs(delta) >= fract1) ? (x + delta) : (delta > 0 ? (x + fabs(fract1)) : (x - fabs(fract1)));
    fu = f(u, it);
    if (fu <= fx) {
      if (u >= x) min = x; else max = x;
      v = w; w = x; x = u; fv = fw; fw = fx; fx = fu;
    } else {
      if (u < x) min = u; else max = u;
      if ((fu <= fw) || (w == x)) { v = w; w = u; fv = fw; fw = fu; }
      else if ((fu <= fv) || (v == x) || (v == w)) { v = u; fv = fu; }
    }
  }
  xo = x; fo = fx;
}

} // namespace krepp
