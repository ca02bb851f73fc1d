// pipe probe: issue rate of VIMNMX3.U16x2 (alu pipe), HMNMX2 (which pipe?), IMAD, and mixes (debug aid, not product)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t hmin2u(uint32_t a, uint32_t b) { uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmax2u(uint32_t a, uint32_t b) { uint32_t r; asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + i + 1); b[i] = seed + i * 77 + threadIdx.x; a[i] &= 0x00ff00ffu; b[i] &= 0x00ff00ffu; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { a[i] = __vimin3_u16x2(a[i], b[i], b[(i + 1) & 7]); b[i] = __vimax3_u16x2(b[i], a[i], a[(i + 3) & 7]); }
      if (MODE == 1) { a[i] = hmin2u(a[i], b[i]); b[i] = hmax2u(b[i], a[(i + 3) & 7]); }
      if (MODE == 2) { a[i] = __vimin3_u16x2(a[i], b[i], b[(i + 1) & 7]); b[i] = hmax2u(b[i], a[(i + 3) & 7]); }
      if (MODE == 3) { a[i] = a[i] * 3 + b[i]; b[i] = b[i] * 5 + a[(i + 3) & 7]; }
      if (MODE == 4) { a[i] = __vimin3_u16x2(a[i], b[i], b[(i + 1) & 7]); b[i] = b[i] * 5 + a[(i + 3) & 7]; }
      if (MODE == 5) { a[i] = __vminu2(a[i], b[i]); b[i] = __vmaxu2(b[i], a[(i + 3) & 7]); }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  uint32_t* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const char* names[] = {"VIMNMX3.U16x2 only", "HMNMX2 only", "VIMNMX3 + HMNMX2 1:1", "IMAD only", "VIMNMX3 + IMAD 1:1", "VIMNMX.U16x2 (2-input) only"};
  const int iters = 4000;
  for (int m = 0; m < 6; m++) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      switch (m) {
        case 0: k<0><<<148 * 8, 256>>>(out, iters, 12345); break;
        case 1: k<1><<<148 * 8, 256>>>(out, iters, 12345); break;
        case 2: k<2><<<148 * 8, 256>>>(out, iters, 12345); break;
        case 3: k<3><<<148 * 8, 256>>>(out, iters, 12345); break;
        case 4: k<4><<<148 * 8, 256>>>(out, iters, 12345); break;
        case 5: k<5><<<148 * 8, 256>>>(out, iters, 12345); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winstr = (double)148 * 8 * 8 * iters * 16;  // warps * instr
    printf("%-28s %.3f ms  %.1f G warp-instr/s  = %.2f instr/clk/SMSP @1.9GHz\n", names[m], ms, winstr / ms / 1e6, winstr / ms / 1e6 / (148 * 4 * 1.9));
  }
  return 0;
}
