// Host build of krepp_b200/csrc/llh_math.cuh for tests/test_llh_math_cpu.py (compiled with -ffp-contract=off, the host
// counterpart of solve.cu's --fmad=false).  Test infrastructure only.
#include "../../krepp_b200/csrc/llh_math.cuh"

using namespace krepp;

template <int N>
struct Plain {
  const Objective<N>* o; const LlhTables* t; mutable uint32_t evals = 0;
  double operator()(double u, int) const { ++evals; return o->eval(*t, u); }
};
template <int N>
struct Memo {
  const Objective<N>* o; const LlhTables* t; const double* su; const DTerms* st; mutable uint32_t hits = 0;
  double operator()(double u, int it) const
  {
    if (it < 3) {
      const int slot = it < 2 ? it : (u == su[2] ? 2 : 3);
      if (u == su[slot]) { ++hits; return o->finish(st[slot]); }
    }
    return o->eval(*t, u);
  }
};

template <int N>
static Objective<N> make(uint32_t k, uint32_t th, const double* hist, double uc, double rho)
{
  Objective<N> f;
  for (int x = 0; x < N; ++x) f.mc[x] = (uint32_t)x <= th ? hist[x] : 0.0;
  f.uc = uc; f.rho = rho; f.k = k; f.th = th;
  return f;
}

extern "C" {
double llh_powi(double x, uint32_t n) { return powi_rounded(x, n); }
void llh_tables_out(uint32_t h, uint32_t k, uint32_t th, double* w) { LlhTables t; llh_tables(t, k, h, th); for (int i = 0; i <= kLlhMaxK; ++i) w[i] = t.w[i]; }
double llh_eval(uint32_t h, uint32_t k, uint32_t th, const double* hist, double uc, double rho, double d)
{
  LlhTables t; llh_tables(t, k, h, th);
  return make<17>(k, th, hist, uc, rho).eval(t, d);
}
// memo = 1: the first abscissae come from the table, as in solve_kernel; *hits = how many evaluations the table served
void llh_brent(uint32_t h, uint32_t k, uint32_t th, const double* hist, double uc, double rho, int memo, double* d, double* v, uint32_t* hits)
{
  LlhTables t; llh_tables(t, k, h, th);
  if (!memo) {
    const Objective<17> f = make<17>(k, th, hist, uc, rho);
    Plain<17> e{&f, &t};
    brent_minimum(e, *d, *v);
    *hits = e.evals;
  } else {
    const Objective<5> f = make<5>(k, th, hist, uc, rho);
    double su[4]; DTerms st[4];
    brent_first_points(su);
    for (int i = 0; i < 4; ++i) st[i] = d_terms(t, su[i], k);
    Memo<5> e{&f, &t, su, st};
    brent_minimum(e, *d, *v);
    *hits = e.hits;
  }
}
}
