// fixed5.h -- a distance as the integer its "%.5f" text shows: round(d * 1e5) of the EXACT binary value, ties to even, i.e. the
// digits glibc printf / std::fixed << setprecision(5) print (ref src/query.cpp:152-153, src/query.hpp:210 DISTANCE_FIELDS).
// The device rounds so that the `dist` rows that leave the GPU are 4 or 8 bytes instead of a double each.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define KREPP_HD __host__ __device__
#else
#define KREPP_HD
#endif

namespace krepp {

// 0 <= d < 40000 (so the result fits 32 bits); anything else (negative, NaN, DBL_MAX of an unsolved record) gives 0xffffffff.
KREPP_HD inline uint32_t fixed5_units(double d)
{
  if (!(d >= 0.0 && d < 40000.0)) return 0xFFFFFFFFu;
  const double x = d * 100000.0;            // rounded product
  const double r = fma(d, 100000.0, -x);    // d * 1e5 = x + r exactly, |r| <= ulp(x) / 2
  const double fl = floor(x);
  const double s = (x - fl) - 0.5;          // x - fl is exact; the subtraction of 0.5 is exact whenever the result is small
  // x - fl is a multiple of ulp(x), so a non-zero s outweighs r; r decides only when x sits exactly on the half
  bool up;
  if (s > 0.0) up = true;
  else if (s < 0.0) up = false;
  else if (r > 0.0) up = true;
  else if (r < 0.0) up = false;
  else up = (((uint64_t)fl) & 1ull) != 0;   // an exact tie: to even, as printf does in round-to-nearest
  return (uint32_t)fl + (up ? 1u : 0u);
}

} // namespace krepp
