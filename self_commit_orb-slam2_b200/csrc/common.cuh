// common.cuh — shared helpers for libb200slam.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/b200slam.h"

namespace b2s {

extern thread_local std::string g_last_error;
void set_error(const char* fmt, ...);

#define B2S_CUDA(call)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) {                                                                           \
      b2s::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));            \
      return (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? B2S_ERR_NO_DEVICE       \
                                                                               : B2S_ERR_CUDA;          \
    }                                                                                                   \
  } while (0)

int select_device(int device);  // returns B2S_OK / B2S_ERR_NO_DEVICE

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- device helpers ----
__device__ __forceinline__ int warp_reduce_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// 256-bit descriptor load/store: one 32-byte global transaction per thread (LDG.E.ENL2.256 / STG.E.ENL2.256)
struct __align__(32) u256 {
  uint32_t w[8];
};
__device__ __forceinline__ u256 ld_u256(const void* p) {
  u256 r;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                 "=r"(r.w[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_u256(void* p, const u256& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]),
               "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7])
               : "memory");
}
__device__ __forceinline__ int hamming256(const u256& a, const u256& b) {
  int d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d += __popc(a.w[i] ^ b.w[i]);
  return d;
}

}  // namespace b2s
