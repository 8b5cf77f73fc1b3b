// The opaque index handle of include/krepp_b200.h, shared by the GPU layer (api.cu) and the host I/O layer (host_io.cpp).
#pragma once
#include "device.cuh"
#include "index_image.hpp"

#include <vector>

struct krepp_index {
  krepp::HostIndex host;
  krepp::DevIndex dev{};
  int device = 0, sms = 0, resident_warps = 0;
  bool staged = false; // match.cu phase B strategy: bucket streaming through shared memory (large buckets) or lane-per-bucket
  bool sorted_ok = false;      // flattened colour lists are resident: the bucket-sorted pipeline (sorted.cu) can run
  bool sorted_default = false; // ... and is what batch slots use unless KREPP_PIPELINE says otherwise
  uint64_t device_bytes = 0;
  std::vector<void*> allocs;
};

namespace krepp {
// sets the thread-local krepp_last_error() text and returns `code`
int set_error(int code, const char* fmt, ...);
}
