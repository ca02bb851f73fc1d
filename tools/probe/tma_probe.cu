// standalone TMA probe: which descriptor / instruction variants work on this box (debug aid, not product)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, const CUtensorMap* gtm, int useGlobal, int bytes, int x, int y, int z, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t b = s32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    const CUtensorMap* t = useGlobal ? gtm : &tm;
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(sm)), "l"(t), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm)), "l"(t), "r"(x), "r"(y), "r"(b) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  for (int i = 0; i < 2000000 && !ok; i++)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
  if (threadIdx.x == 0) { out[0] = ok; out[1] = sm[0] | (sm[1] << 8) | (sm[2] << 16) | (sm[3] << 24); out[2] = sm[bytes - 1]; }
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
  printf("entry point: %s qr=%d p=%p\n", cudaGetErrorString(e), (int)qr, p);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  const int W = 640, H = 480, B = 2, pitch = 640; const size_t img = 400384;  // multiple of 256
  uint8_t* d; cudaMalloc(&d, img * B);
  std::vector<uint8_t> h(img * B);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 8));
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  uint32_t* out; cudaMalloc(&out, 64);
  CUtensorMap* gtm; cudaMalloc(&gtm, sizeof(CUtensorMap));
  struct V { int rank, bw, bh; int useGlobal; int x, y, z; CUtensorMapL2promotion l2; const char* name; };
  V vs[] = {
    {3, 160, 38, 0, 128, 29, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "3d 160x38 l2-128 x=128 y=29"},
    {3, 160, 38, 0, 112, 29, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "3d 160x38 l2-128 x=112 y=29"},
    {3, 160, 38, 0, -16, -3, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "3d 160x38 l2-128 x=-16 y=-3"},
    {3, 160, 38, 1, -16, -3, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "3d 160x38 GLOBAL desc x=-16 y=-3"},
    {3, 160, 38, 0, 624, 477, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "3d 160x38 past the edge x=624 y=477"},
    {3, 144, 38, 0, 120, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE, "3d 144x38 none x=120 (8 mod 16)"},
  };
  for (V& v : vs) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)img};
    cuuint32_t box[3] = {(cuuint32_t)v.bw, (cuuint32_t)v.bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, v.rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, v.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    cudaMemset(out, 0, 64);
    const int bytes = v.bw * v.bh;
    if (v.rank == 3) k<3><<<1, 32, bytes + 128>>>(tm, gtm, v.useGlobal, bytes, v.x, v.y, v.z, out);
    else k<2><<<1, 32, bytes + 128>>>(tm, gtm, v.useGlobal, bytes, v.x, v.y, v.z, out);
    cudaError_t ce = cudaDeviceSynchronize();
    uint32_t ho[4] = {0, 0, 0, 0};
    if (ce == cudaSuccess) cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
    size_t o0 = (size_t)v.z * img + (size_t)(v.y < 0 ? 0 : v.y) * pitch + (v.x < 0 ? 0 : v.x);
    printf("%-45s encode=%d run=%s done=%u first=%08x expect(first in-image bytes)=%02x%02x%02x%02x last=%02x\n", v.name, (int)r, cudaGetErrorString(ce),
           ho[0], ho[1], h[o0 + 3], h[o0 + 2], h[o0 + 1], h[o0], ho[2]);
    if (ce != cudaSuccess) { printf("sticky error, stopping\n"); break; }
  }
  return 0;
}
