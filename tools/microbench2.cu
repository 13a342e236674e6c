// microbench2.cu - does a 16-byte-misaligned streaming read over-fetch?  (ncu: dram__bytes_read, lts tex reads)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
template <int MODE> __device__ __forceinline__ uint4 ld(const uint8_t* p) {
  uint4 r;
  if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  if (MODE == 1) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
template <int MODE, int OFF>
__global__ void __launch_bounds__(256) k(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t nvec) {
  uint32_t base = blockIdx.x * 256 * 8 + threadIdx.x;
  uint4 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) if (base + i * 256 < nvec) a[i] = ld<MODE>(src + OFF + 16ull * (base + i * 256));
#pragma unroll
  for (int i = 0; i < 8; ++i) if (base + i * 256 < nvec) *reinterpret_cast<uint4*>(dst + 16ull * (base + i * 256)) = a[i];
}
int main() {
  const uint32_t P = 4u << 20, nvec = P / 16;
  uint8_t *s, *d;
  CK(cudaMalloc(&s, P + 4096)); CK(cudaMalloc(&d, P));
  CK(cudaMemset(s, 1, P + 4096));
  const uint32_t g = (nvec + 2047) / 2048;
  k<0, 0><<<g, 256>>>(s, d, nvec);  k<0, 16><<<g, 256>>>(s, d, nvec); k<1, 16><<<g, 256>>>(s, d, nvec);
  k<2, 16><<<g, 256>>>(s, d, nvec); k<3, 16><<<g, 256>>>(s, d, nvec); k<0, 32><<<g, 256>>>(s, d, nvec);
  k<0, 48><<<g, 256>>>(s, d, nvec); k<0, 64><<<g, 256>>>(s, d, nvec);
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
