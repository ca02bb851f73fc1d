// Issue rate of the legacy mma.sync int8 path on sm_100a (IMMA.16832.U8.U8): is an all-pairs Hamming distance
// (popc(a & b) as a 0/1-byte dot product) on it worth building?  Prints G warp-instructions/s and the equivalent
// descriptor-pair rate (8 IMMAs of k = 32 make one 16 x 8 tile of 256-bit distances).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void __launch_bounds__(256) k(const uint32_t* a, int* out, int iters) {
  uint32_t a0 = a[threadIdx.x], a1 = a[threadIdx.x + 32], a2 = a[threadIdx.x + 64], a3 = a[threadIdx.x + 96];
  uint32_t b0 = a[threadIdx.x + 128], b1 = a[threadIdx.x + 160];
  int c[NACC][4];
  for (int q = 0; q < NACC; q++) c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < NACC; q++)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[q][0]), "+r"(c[q][1]), "+r"(c[q][2]), "+r"(c[q][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  int s = 0;
  for (int q = 0; q < NACC; q++) s += c[q][0] + c[q][1] + c[q][2] + c[q][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(const char* name) {
  uint32_t* a; int* o;
  cudaMalloc(&a, 4096); cudaMemset(a, 1, 4096);
  cudaMalloc(&o, 148 * 8 * 256 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0);
    k<NACC><<<148 * 8, 256>>>(a, o, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  const double winstr = 148.0 * 8 * 8 * (double)iters * NACC;
  printf("%s: %.3f ms, %.1f G IMMA warp-instr/s, %.2f T pair-distances/s (8 IMMA per 128 distances), err=%s\n", name, best,
         winstr / best / 1e6, winstr / best / 1e6 * 16 / 1e3, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<1>("1 accumulator chain"); run<4>("4 chains"); run<8>("8 chains"); return 0; }
