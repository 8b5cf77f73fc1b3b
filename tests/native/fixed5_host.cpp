// Host build of krepp_b200/csrc/fixed5.h for tests/test_fixed5_cpu.py (the same function the finalize kernels call).
#include "../../krepp_b200/csrc/fixed5.h"
extern "C" void fixed5_many(const double* d, uint32_t* out, uint64_t n) { for (uint64_t i = 0; i < n; ++i) out[i] = krepp::fixed5_units(d[i]); }
