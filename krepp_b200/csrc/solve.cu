// solve.cu -- K4 of the query path: maximum-likelihood distance per (read, strand, leaf) record, then the per-read
// strand merge / closest selection and the likelihood-ratio statistic.
//
// THIS FILE IS COMPILED WITH --fmad=false: the reference runs on x86-64 without FMA contraction, and Brent's
// parabolic-vs-golden decisions (tolerance only 2^-15) must follow the same iteration
// sequence, so every product and sum below rounds separately, in the reference's literal evaluation order.
//
//   objective  : optimize::HDistHistLLH::operator()          ref src/hdhistllh.hpp:71-89
//   tables     : HDistHistLLH ctor (host, passed by value)   ref src/hdhistllh.hpp:51-69
//   minimiser  : boost::math::tools::brent_find_minima(f, 1e-10, 0.5, 16)
//                ref external/boost/libs/math/include/boost/math/tools/minima.hpp:23-138, called at src/query.cpp:430
//   gate/merge : IBatch::summarize_matches                   ref src/query.cpp:96-139
//   chisq      : Minfo::likelihood_ratio                     ref src/query.cpp:420-424
#include "device.cuh"
#include "solve.cuh"

#include <cfloat>

namespace krepp {

struct Objective {
  const LlhTables* t;
  const double* mc; // hist as doubles
  double uc, rho;
  uint32_t k, th;
  __device__ double operator()(double d) const
  {
    double sum = 0.0, lv_m = 0.0;
    double powdc = pow((1.0 - d), (double)k);
    double logdn = log(1.0 - d);
    double logdp = log(d) - logdn;
    logdn *= (double)k;
    const double dratio = d / (1.0 - d);
    for (uint32_t x = 0; x <= k; ++x) {
      if (x <= th) {
        sum -= (logdn + (double)x * logdp) * mc[x];
        lv_m += t->hnk[x] * powdc;
      } else {
        lv_m += powdc * t->ck[x];
      }
      powdc *= dratio;
    }
    return sum - log(rho * lv_m + 1.0 - rho) * uc;
  }
};

__device__ void brent_minimum(const Objective& f, double& xo, double& fo)
{
  double min = 1e-10, max = 0.5;
  const double tolerance = 3.0517578125e-05; // ldexp(1.0, 1 - 16)
  const double golden = (double)0.3819660f;
  double x, w, v, u, delta, delta2, fu, fv, fw, fx, mid, fract1, fract2;
  x = w = v = max;
  fw = fv = fx = f(x);
  delta2 = delta = 0;
  for (;;) {
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
    fu = f(u);
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

// One thread per record: match_count / hdist_min from the histogram, the hdist_filt gate, then Brent.
__global__ void __launch_bounds__(128) solve_kernel(const SolveArgs a, const LlhTables tab)
{
  const uint32_t n = a.counters[0] < a.n_records ? a.counters[0] : a.n_records;
  const uint32_t stride = a.th + 1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t read = a.rec_read[i], slot = a.rec_slot[i];
    const uint32_t strand = slot >> 31, se = slot & 0x7FFFFFFFu;
    double mc[kMaxTh + 1];
    uint32_t match = 0, hdmin = 0xFFFFFFFFu;
    for (uint32_t x = 0; x < stride; ++x) {
      const uint32_t c = a.rec_hist[(size_t)i * stride + x];
      mc[x] = (double)c;
      match += c;
      if (c && hdmin == 0xFFFFFFFFu) hdmin = x;
    }
    a.rec_match[i] = match; a.rec_hdmin[i] = hdmin;
    const uint32_t filt = 2u * a.hdfilt[2 * read + strand] + 1u; // uint32 wrap kept (ref src/query.cpp:101-102)
    double d = DBL_MAX, v = nan(""); // Minfo defaults (ref src/query.hpp:225-226)
    uint32_t flags = 0;
    if (!(hdmin > filt)) {
      Objective f{&tab, mc, (double)a.onmers[read] - (double)match, a.rho[se], a.k, a.th};
      brent_minimum(f, d, v);
      flags = 1u; // KREPP_REC_SOLVED
    }
    a.rec_d[i] = d; a.rec_v[i] = v; a.rec_flags[i] = flags; a.rec_chisq[i] = nan("");
  }
}

// One thread per read: closest + per-leaf strand choice, in the fixed visiting order (forward leaves by ascending se,
// then reverse leaves by ascending se; `<=` kept so the last tied entry wins -- SURVEY.md section 0 fact 6).
__global__ void __launch_bounds__(128) merge_kernel(const SolveArgs a)
{
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += gridDim.x * blockDim.x) {
    const uint32_t b = a.rec_begin[r], n = a.rec_count[r];
    int32_t cl = -1;
    double best = DBL_MAX;
    uint32_t nf = 0;
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t g = b + i;
      if (!(a.rec_slot[g] >> 31)) ++nf;
      if ((a.rec_flags[g] & 1u) && a.rec_d[g] <= best) { best = a.rec_d[g]; cl = (int32_t)g; }
    }
    // two-pointer walk over forward [b, b+nf) and reverse [b+nf, b+n), both ascending in se
    uint32_t i = b, j = b + nf;
    const uint32_t ie = b + nf, je = b + n;
    while (i < ie || j < je) {
      const uint32_t si = i < ie ? (a.rec_slot[i] & 0x7FFFFFFFu) : 0xFFFFFFFFu;
      const uint32_t sj = j < je ? (a.rec_slot[j] & 0x7FFFFFFFu) : 0xFFFFFFFFu;
      if (si < sj) { if (a.rec_flags[i] & 1u) a.rec_flags[i] |= 2u; ++i; }
      else if (sj < si) { if (a.rec_flags[j] & 1u) a.rec_flags[j] |= 2u; ++j; }
      else {
        const bool fs = a.rec_flags[i] & 1u, rs = a.rec_flags[j] & 1u;
        if (rs) {
          const double dr = a.rec_d[j], df = a.rec_d[i];
          const bool fwd_wins = (dr > df) || ((dr == df) && (a.rec_match[j] < a.rec_match[i]));
          if (fwd_wins) a.rec_flags[i] |= 2u; else a.rec_flags[j] |= 2u;
        } else if (fs) a.rec_flags[i] |= 2u;
        ++i; ++j;
      }
    }
    if (cl >= 0) { // node_to_minfo[nd_closest] = mi_closest (ref src/query.cpp:136-138)
      const uint32_t se = a.rec_slot[cl] & 0x7FFFFFFFu;
      for (uint32_t q = b; q < b + n; ++q)
        if ((a.rec_slot[q] & 0x7FFFFFFFu) == se) a.rec_flags[q] &= ~2u;
      a.rec_flags[cl] |= 2u | 4u;
    }
    a.closest[r] = cl;
  }
}

// One thread per selected record: chisq = 2 * (f_closest(d_record) - v_closest).
__global__ void __launch_bounds__(128) chisq_kernel(const SolveArgs a, const LlhTables tab)
{
  const uint32_t n = a.counters[0] < a.n_records ? a.counters[0] : a.n_records;
  const uint32_t stride = a.th + 1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!(a.rec_flags[i] & 2u)) continue;
    const uint32_t read = a.rec_read[i];
    const int32_t cl = a.closest[read];
    if (cl < 0) continue;
    double mc[kMaxTh + 1];
    for (uint32_t x = 0; x < stride; ++x) mc[x] = (double)a.rec_hist[(size_t)cl * stride + x];
    Objective f{&tab, mc, (double)a.onmers[read] - (double)a.rec_match[cl], a.rho[a.rec_slot[cl] & 0x7FFFFFFFu], a.k, a.th};
    a.rec_chisq[i] = 2 * (f(a.rec_d[i]) - a.rec_v[cl]);
  }
}

cudaError_t launch_solve(const SolveArgs& a, const LlhTables& tab, int sms, cudaStream_t stream)
{
  const int grid = sms * 8;
  solve_kernel<<<grid, 128, 0, stream>>>(a, tab);
  merge_kernel<<<grid, 128, 0, stream>>>(a);
  if (a.want_chisq) chisq_kernel<<<grid, 128, 0, stream>>>(a, tab);
  return cudaGetLastError();
}

} // namespace krepp
