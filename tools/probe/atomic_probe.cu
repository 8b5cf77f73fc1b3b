// probe: does the tagged atomicMin scheme telescope correctly when many lanes of one warp hit the same leaf at once?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void leaf_hit(uint32_t* marker, uint32_t* acc, uint32_t hd, uint32_t tagbase)
{
  const uint32_t old = atomicMin(marker, tagbase | hd);
  if ((old ^ tagbase) >> 5) atomicAdd(&acc[hd], 1u);
  else if ((old & 31u) > hd) { atomicSub(&acc[old & 31u], 1u); atomicAdd(&acc[hd], 1u); }
}
__global__ void probe(uint32_t* marker, uint32_t* acc, const uint32_t* hds, int nlook, uint32_t* olds)
{
  const uint32_t lane = threadIdx.x;
  uint32_t tag = 0x07FFFFFEu;
  for (int l = 0; l < nlook; ++l) {
    --tag;
    const uint32_t hd = hds[l * 32 + lane];
    if (hd <= 4) {
      const uint32_t old = atomicMin(marker, (tag << 5) | hd);
      olds[l * 32 + lane] = old;
      if ((old ^ (tag << 5)) >> 5) atomicAdd(&acc[hd], 1u);
      else if ((old & 31u) > hd) { atomicSub(&acc[old & 31u], 1u); atomicAdd(&acc[hd], 1u); }
    }
    __syncwarp();
  }
}
int main()
{
  uint32_t *marker, *acc, *hds, *olds;
  cudaMalloc(&marker, 4); cudaMalloc(&acc, 64); cudaMalloc(&hds, 4 * 32 * 4); cudaMalloc(&olds, 4 * 32 * 4);
  cudaMemset(marker, 0xFF, 4); cudaMemset(acc, 0, 64); cudaMemset(olds, 0, 4 * 32 * 4);
  uint32_t h[4 * 32];
  for (int i = 0; i < 128; ++i) h[i] = 99;
  // lookup 0: hits hd 4,3,2,1,4,4,4 on lanes 1,3,4,5,8,12,14 (read 1405 of the 20k test); lookup 1: 4,3,2,0; lookup 2: 0 only; lookup 3: 3,3,3
  int l0[7] = {1, 3, 4, 5, 8, 12, 14}; uint32_t v0[7] = {4, 3, 2, 1, 4, 4, 4};
  for (int i = 0; i < 7; ++i) h[l0[i]] = v0[i];
  h[32 + 1] = 4; h[32 + 2] = 3; h[32 + 5] = 2; h[32 + 6] = 0;
  h[64 + 9] = 0;
  h[96 + 0] = 3; h[96 + 1] = 3; h[96 + 31] = 3;
  cudaMemcpy(hds, h, sizeof h, cudaMemcpyHostToDevice);
  probe<<<1, 32>>>(marker, acc, hds, 4, olds);
  uint32_t a[16], o[128], m;
  cudaMemcpy(a, acc, 64, cudaMemcpyDeviceToHost); cudaMemcpy(o, olds, sizeof o, cudaMemcpyDeviceToHost); cudaMemcpy(&m, marker, 4, cudaMemcpyDeviceToHost);
  printf("err=%s\nhist: %u %u %u %u %u   (expect 2 1 0 1 0)\nmarker %08x\n", cudaGetErrorString(cudaDeviceSynchronize()), a[0], a[1], a[2], a[3], a[4], m);
  for (int l = 0; l < 4; ++l) { printf("lookup %d olds:", l); for (int i = 0; i < 32; ++i) if (h[l * 32 + i] <= 4) printf(" lane%d(hd%u)=%08x", i, h[l * 32 + i], o[l * 32 + i]); printf("\n"); }
  return 0;
}
