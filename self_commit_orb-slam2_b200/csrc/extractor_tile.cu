// extractor_tile.cu — the fused front end of the B200 ORB extractor: ONE tile kernel per pyramid level that
//   * stages a 128 x 32 pixel tile (+ halo) of level l in shared memory with TMA (cp.async.bulk.tensor, mbarrier),
//   * computes the FAST-9/16 arc strength of EVERY pixel of the tile once (packed u16x2 DPX min/max; one strength map
//     serves both thresholds),
//   * computes the Q8 7x7 Gaussian blur of the tile (REFLECT_101 at the image border),
//   * produces the part of pyramid level l+1 whose bilinear source samples start inside the tile,
// and writes the strength map / blurred tile back with TMA stores.  A light per-cell kernel then applies the reference's
// cell rules (3x3 NMS that never crosses a cell seam, iniThFAST -> minThFAST fallback per cell) and emits the candidate
// lists the quadtree kernel consumes.
//
// Replaces ComputePyramid (/root/reference/src/ORBextractor.cc:1674-1734), the cell loop + cv::FAST of
// ComputeKeyPointsOctTree (:1060-1157) and the GaussianBlur of operator() (:1626-1634).  Bit-exact: the arithmetic is
// the same integer recipe as the per-stage kernels in extractor.cu (which remain as the B2S_EXTRACT_PATH=0 fallback and
// as the parity cross-check).
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

#include <algorithm>
#include <vector>

#include "extractor.cuh"

namespace b2s {

constexpr int TW = 128, TH = 32;          // tile interior
// halo left / top (right / bottom are HX / HY as well).  Only 3 pixels are needed, but the innermost TMA coordinate must
// be a multiple of 16 BYTES (x = tx*128 - 8 raises "illegal instruction" on B200; measured with tools/probe/tma_probe.cu)
constexpr int HX = 16, HY = 3;
constexpr int BW = TW + 2 * HX;           // 160: TMA box width in bytes (multiple of 16)
constexpr int BH = TH + 2 * HY;           // 38
constexpr int PWD = BW / 2;               // 80 pair words per plane row
constexpr int TILE_THREADS = 256;

struct TileSmem {
  alignas(128) uint8_t raw[BH * BW];        // TMA destination: rows ty*32-3 .., columns tx*128-16 ..
  alignas(128) uint8_t scoreT[TH * TW];     // TMA store source: FAST arc strength M (0..255) per interior pixel
  alignas(128) uint8_t blurT[TH * TW];      // TMA store source: blurred interior
  alignas(16) uint32_t planeE[BH * PWD];    // (raw[2j], raw[2j+1]) as u16x2
  alignas(16) uint32_t planeO[BH * PWD];    // (raw[2j+1], raw[2j+2])
  alignas(16) uint16_t blurH[BH * TW];      // horizontal pass of the blur, rows of the box x interior columns
  alignas(8) unsigned long long bar;        // mbarrier of the tile load
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(src), "r"(x),
               "r"(y), "r"(z)
               : "memory");
}

struct TileArgs {
  int level, bBase, doFast, doBlur, doResize;
  uint8_t* pyr;  // base of the batch's pyramids (image 0 of the handle), for the level l+1 stores
  const int16_t* rxOfs;
  const uint32_t* rxAlpha;
  const int16_t* ryOfs;
  const uint32_t* ryBeta;
  const int16_t* tileDx;  // per (level, tile column) first destination column owned; one sentinel per level
  const int16_t* tileDy;
};

// FAST-9/16 arc strength of the pixel pair whose left pixel sits at EVEN box column 2*j (row pointers at the pair's row).
// With I_k the 16 circle pixels, A = max_s min_{9 window} I, B = min_s max_{9 window} I and c the centre:
// bright strength = A - c, dark strength = c - B, M = max(bright, dark, 0) — identical to max over the 16 arcs of
// min |v - p| with one sign (extractor.cu fast_pair_full); corner at th <=> M > th, response = M - 1.
__device__ __forceinline__ uint32_t fast_pair_strength(const uint32_t* __restrict__ E, const uint32_t* __restrict__ O, int j) {
  uint32_t d[16];
  d[0] = E[3 * PWD + j];        // ( 0, 3)
  d[1] = O[3 * PWD + j];        // ( 1, 3)
  d[2] = E[2 * PWD + j + 1];    // ( 2, 2)
  d[3] = O[1 * PWD + j + 1];    // ( 3, 1)
  d[4] = O[j + 1];              // ( 3, 0)
  d[5] = O[-1 * PWD + j + 1];   // ( 3,-1)
  d[6] = E[-2 * PWD + j + 1];   // ( 2,-2)
  d[7] = O[-3 * PWD + j];       // ( 1,-3)
  d[8] = E[-3 * PWD + j];       // ( 0,-3)
  d[9] = O[-3 * PWD + j - 1];   // (-1,-3)
  d[10] = E[-2 * PWD + j - 1];  // (-2,-2)
  d[11] = O[-1 * PWD + j - 2];  // (-3,-1)
  d[12] = O[j - 2];             // (-3, 0)
  d[13] = O[1 * PWD + j - 2];   // (-3, 1)
  d[14] = E[2 * PWD + j - 1];   // (-2, 2)
  d[15] = O[3 * PWD + j - 1];   // (-1, 3)
  const uint32_t c = E[j];
  uint32_t t3n[16], t3x[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    t3n[k] = __vimin3_u16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    t3x[k] = __vimax3_u16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
  }
  uint32_t A = 0u, B = 0xFFFFFFFFu;
#pragma unroll
  for (int s = 0; s < 16; s += 2) {
    const uint32_t a0 = __vimin3_u16x2(t3n[s], t3n[(s + 3) & 15], t3n[(s + 6) & 15]);
    const uint32_t a1 = __vimin3_u16x2(t3n[s + 1], t3n[(s + 4) & 15], t3n[(s + 7) & 15]);
    A = __vimax3_u16x2(A, a0, a1);
    const uint32_t b0 = __vimax3_u16x2(t3x[s], t3x[(s + 3) & 15], t3x[(s + 6) & 15]);
    const uint32_t b1 = __vimax3_u16x2(t3x[s + 1], t3x[(s + 4) & 15], t3x[(s + 7) & 15]);
    B = __vimin3_u16x2(B, b0, b1);
  }
  // biased per-lane differences stay in [1, 511]: a plain 32-bit add / subtract never carries across the lanes
  const uint32_t bright = (A + 0x01000100u) - c;
  const uint32_t dark = (c + 0x01000100u) - B;
  return __vimax3_u16x2(bright, dark, 0x01000100u) - 0x01000100u;  // M(lo) | M(hi) << 16, each in [0, 255]
}

__global__ void __launch_bounds__(TILE_THREADS) k_tile(const __grid_constant__ CUtensorMap tmPyr,
                                                       const __grid_constant__ CUtensorMap tmScore,
                                                       const __grid_constant__ CUtensorMap tmBlur,
                                                       const __grid_constant__ ExtractGeom g, const TileArgs a) {
  extern __shared__ __align__(128) uint8_t smemRaw[];
  TileSmem& S = *reinterpret_cast<TileSmem*>((reinterpret_cast<uintptr_t>(smemRaw) + 127) & ~(uintptr_t)127);
  const LevelGeom& L = g.lv[a.level];
  const int tx = blockIdx.x, ty = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const int X0 = tx * TW - HX, Y0 = ty * TH - HY;  // image coordinates of box (0, 0)
  const uint32_t bar = s32(&S.bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, BH * BW);
    tma_load_3d(s32(S.raw), &tmPyr, bar, X0, Y0, a.bBase + b);  // out-of-image bytes arrive as 0
  }
  __syncthreads();  // the barrier is initialised for everyone
  while (!mbar_try_wait(bar, 0)) {
  }

  // ---- REFLECT_101 border of cv::GaussianBlur for tiles that touch the image edge: columns first, then whole rows
  const int Wd = L.w, Hd = L.h;
  const bool edgeX = (X0 + HX - 3 < 0) || (X0 + HX + TW + 3 > Wd);
  const bool edgeY = (Y0 < 0) || (Y0 + BH > Hd);
  if (a.doBlur && (edgeX || edgeY)) {
    if (edgeX) {
      for (int k = tid; k < BH * 6; k += TILE_THREADS) {
        const int ly = k / 6, c6 = k - ly * 6;
        const int gy = Y0 + ly;
        if (gy < 0 || gy >= Hd) continue;
        const int gx = (c6 < 3) ? (c6 - 3) : (Wd + c6 - 3);
        const int lx = gx - X0;
        if (lx < 0 || lx >= BW) continue;
        const int rx = (gx < 0) ? -gx : 2 * (Wd - 1) - gx;
        const int lsx = rx - X0;
        if (lsx >= 0 && lsx < BW) S.raw[ly * BW + lx] = S.raw[ly * BW + lsx];
      }
      __syncthreads();
    }
    if (edgeY) {
      for (int k = tid; k < 6 * (BW / 4); k += TILE_THREADS) {
        const int r6 = k / (BW / 4), wq = k - r6 * (BW / 4);
        const int gy = (r6 < 3) ? (r6 - 3) : (Hd + r6 - 3);
        const int ly = gy - Y0;
        if (ly < 0 || ly >= BH) continue;
        const int ry = (gy < 0) ? -gy : 2 * (Hd - 1) - gy;
        const int lsy = ry - Y0;
        if (lsy >= 0 && lsy < BH)
          reinterpret_cast<uint32_t*>(S.raw)[ly * (BW / 4) + wq] = reinterpret_cast<const uint32_t*>(S.raw)[lsy * (BW / 4) + wq];
      }
      __syncthreads();
    }
  }

  // ---- FAST: pixel-pair planes, then the dense arc strength of the interior
  if (a.doFast) {
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(S.raw);
    // (only the words the 16-point circles of the interior can reach: box columns [HX - 4, HX + TW + 4))
    constexpr int Q0 = HX / 4 - 1, QN = TW / 4 + 2;
    for (int k = tid; k < BH * QN; k += TILE_THREADS) {
      const int r = k / QN, q = Q0 + (k - r * QN);
      const uint32_t w0 = rw[r * (BW / 4) + q];
      const uint32_t w1 = rw[r * (BW / 4) + q + 1];
      uint2 e, o;
      e.x = __byte_perm(w0, 0u, 0x4140);                 // (b0, b1)
      e.y = __byte_perm(w0, 0u, 0x4342);                 // (b2, b3)
      o.x = __byte_perm(w0, 0u, 0x4241);                 // (b1, b2)
      o.y = (w0 >> 24) | ((w1 & 0xffu) << 16);           // (b3, b4)
      *reinterpret_cast<uint2*>(&S.planeE[r * PWD + 2 * q]) = e;
      *reinterpret_cast<uint2*>(&S.planeO[r * PWD + 2 * q]) = o;
    }
    __syncthreads();
    // rows outside the FAST band [16+3, maxB-3) of the level are never read by the cell kernel: skip them (a warp works
    // on half a row, so the test is warp-uniform)
    const int yLo = kMinBorder + 3, yHi = L.maxBY - 3;
#pragma unroll 1
    for (int it = 0; it < (TW / 2) * TH / TILE_THREADS; it++) {
      const int q = it * TILE_THREADS + tid;
      const int row = q >> 6, pc = q & 63;  // interior row, pair column
      const int gy = ty * TH + row;
      uint32_t M2 = 0u;
      if (gy >= yLo && gy < yHi) {
        const int ly = row + HY;
        M2 = fast_pair_strength(&S.planeE[ly * PWD], &S.planeO[ly * PWD], HX / 2 + pc);
      }
      reinterpret_cast<uint16_t*>(S.scoreT)[row * (TW / 2) + pc] = (uint16_t)((M2 & 0xffu) | ((M2 >> 8) & 0xff00u));
    }
  }

  // ---- Gaussian blur 7x7, sigma 2, Q8 taps [18,34,48,56,48,34,18] (OpenCV >= 4 fixed point)
  if (a.doBlur) {
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(S.raw);
    const uint32_t KA = 0x38302212u, KB = 0x00122230u;  // taps (18,34,48,56) and (48,34,18,0)
    for (int k = tid; k < BH * (TW / 4); k += TILE_THREADS) {
      const int r = k >> 5, q = k & 31;  // box row, group of 4 interior columns
      const int wq = r * (BW / 4) + HX / 4 + q;
      const uint32_t w0 = rw[wq - 1], w1 = rw[wq], w2 = rw[wq + 1];
      const uint32_t h0 = __dp4a(__byte_perm(w1, w2, 0x4321), KB, __dp4a(__byte_perm(w0, w1, 0x4321), KA, 0u));
      const uint32_t h1 = __dp4a(__byte_perm(w1, w2, 0x5432), KB, __dp4a(__byte_perm(w0, w1, 0x5432), KA, 0u));
      const uint32_t h2 = __dp4a(__byte_perm(w1, w2, 0x6543), KB, __dp4a(__byte_perm(w0, w1, 0x6543), KA, 0u));
      const uint32_t h3 = __dp4a(w2, KB, __dp4a(w1, KA, 0u));
      uint2 hv;
      hv.x = h0 | (h1 << 16);
      hv.y = h2 | (h3 << 16);
      *reinterpret_cast<uint2*>(&S.blurH[r * TW + 4 * q]) = hv;
    }
    __syncthreads();
    {
      // 32 column groups (4 pixels) x 8 strips of 4 output rows: a register window of 7 horizontal sums slides down
      const int q = tid & 31, strip = tid >> 5;
      const uint2* hp = reinterpret_cast<const uint2*>(&S.blurH[(strip * 4) * TW + 4 * q]);
      uint32_t win[7][4];
#pragma unroll
      for (int i = 0; i < 10; i++) {
        const uint2 hv = hp[i * (TW / 4)];
        const int u = i % 7;
        win[u][0] = hv.x & 0xffffu;
        win[u][1] = hv.x >> 16;
        win[u][2] = hv.y & 0xffffu;
        win[u][3] = hv.y >> 16;
        if (i >= 6) {
          uint32_t outw = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t q0 = win[(u + 1) % 7][j], q1 = win[(u + 2) % 7][j], q2 = win[(u + 3) % 7][j],
                           q3 = win[(u + 4) % 7][j], q4 = win[(u + 5) % 7][j], q5 = win[(u + 6) % 7][j], q6 = win[u][j];
            const uint32_t acc = 18u * (q0 + q6) + 34u * (q1 + q5) + 48u * (q2 + q4) + 56u * q3;
            outw |= ((acc + 32768u) >> 16) << (8 * j);
          }
          reinterpret_cast<uint32_t*>(S.blurT)[(strip * 4 + i - 6) * (TW / 4) + q] = outw;
        }
      }
    }
  }

  // ---- level l+1: cv::resize INTER_LINEAR fixed point, every destination pixel whose top-left source sample lies in
  // this tile's interior (the +1 neighbours are inside the halo)
  if (a.doResize && a.level + 1 < g.nlevels) {
    const LevelGeom& D = g.lv[a.level + 1];
    const int16_t* tdx = a.tileDx + L.tdxOff;
    const int16_t* tdy = a.tileDy + L.tdyOff;
    const int dx0 = tdx[tx], dx1 = tdx[tx + 1], dy0 = tdy[ty], dy1 = tdy[ty + 1];
    const int nw = dx1 - dx0, nh = dy1 - dy0;
    uint8_t* dst = a.pyr + (size_t)(a.bBase + b) * g.pyrBytes + D.off;
    for (int k = tid; k < nw * nh; k += TILE_THREADS) {
      const int ry = k / nw, rx = k - ry * nw;
      const int dx = dx0 + rx, dy = dy0 + ry;
      const int sx = a.rxOfs[D.rxOff + dx], sy = a.ryOfs[D.ryOff + dy];
      const uint32_t aa = a.rxAlpha[D.rxOff + dx], bb = a.ryBeta[D.ryOff + dy];
      const int a0 = (int16_t)(aa & 0xffffu), a1 = (int16_t)(aa >> 16);
      const int b0 = (int16_t)(bb & 0xffffu), b1 = (int16_t)(bb >> 16);
      const int sx1 = min(sx + 1, Wd - 1);
      const int sy0 = min(max(sy, 0), Hd - 1), sy1 = min(max(sy + 1, 0), Hd - 1);
      const uint8_t* r0 = S.raw + (sy0 - Y0) * BW - X0;
      const uint8_t* r1 = S.raw + (sy1 - Y0) * BW - X0;
      const int h0 = r0[sx] * a0 + r0[sx1] * a1;
      const int h1 = r1[sx] * a0 + r1[sx1] * a1;
      dst[(size_t)dy * D.pitch + dx] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
    }
  }

  // ---- results out: the generic-proxy writes to shared memory become visible to the async proxy, then one thread issues
  // the bulk tensor stores (clipped at the image edge by the hardware)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    if (a.doFast) tma_store_3d(&tmScore, s32(S.scoreT), tx * TW, ty * TH, a.bBase + b);
    if (a.doBlur) tma_store_3d(&tmBlur, s32(S.blurT), tx * TW, ty * TH, a.bBase + b);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Per-cell rules of ComputeKeyPointsOctTree (src/ORBextractor.cc:1089-1157) on the strength map: one warp per cell.
// A pixel is a keypoint at threshold th iff M > th and M is strictly greater than its 8 neighbours INSIDE the cell's
// detectable area (cv::FAST runs on the cell ROI: the 3-pixel rim scores 0, so maxima never see across a seam).  Because
// M is the threshold-independent arc strength, the same maxima serve iniThFAST and minThFAST; a cell falls back to
// minThFAST only if it has no maximum above iniThFAST.
// ------------------------------------------------------------------------------------------------
constexpr int CP = 68;    // byte pitch of a cell's strength window (<= 64 columns + ring)
constexpr int CR = 66;    // rows (<= 64 + ring)
constexpr int CLIST = 1056;

__global__ void __launch_bounds__(128) k_cells(const __grid_constant__ ExtractGeom g, const uint8_t* __restrict__ score,
                                               const uint32_t* __restrict__ cellInfo, uint32_t* __restrict__ candXY,
                                               uint32_t* __restrict__ candKey, uint8_t* __restrict__ candResp,
                                               int32_t* __restrict__ candCount, int32_t* __restrict__ status) {
  __shared__ __align__(4) uint8_t scAll[4][CR * CP];
  __shared__ uint32_t listAll[4][CLIST];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cid = blockIdx.x * 4 + warp;
  const int b = blockIdx.y;
  if (cid >= g.totalCells) return;
  const uint32_t info = cellInfo[cid];
  const int l = info >> 28, ci = (info >> 14) & 0x3fff, cj = info & 0x3fff;
  const LevelGeom& L = g.lv[l];
  const int iniY = kMinBorder + ci * L.hCell;
  int maxY = iniY + L.hCell + 6;
  if (iniY >= L.maxBY - 3) return;  // :1099
  if (maxY > L.maxBY) maxY = L.maxBY;
  const int iniX = kMinBorder + cj * L.wCell;
  int maxX = iniX + L.wCell + 6;
  if (iniX >= L.maxBX - 6) return;  // :1116
  if (maxX > L.maxBX) maxX = L.maxBX;
  const int rw = maxX - iniX, rh = maxY - iniY;
  if (rw < 7 || rh < 7) return;
  const int aw = rw - 6, ah = rh - 6;  // detectable area, starts at (iniX + 3, iniY + 3)
  uint8_t* sc = scAll[warp];
  uint32_t* list = listAll[warp];
  for (int k = lane; k < (ah + 2) * (CP / 4); k += 32) reinterpret_cast<uint32_t*>(sc)[k] = 0u;
  __syncwarp();
  const uint8_t* src = score + (size_t)b * g.pyrBytes + L.off + (size_t)(iniY + 3) * L.pitch + (iniX + 3);
  for (int r = 0; r < ah; r++)
    for (int x = lane; x < aw; x += 32) sc[(r + 1) * CP + x + 1] = src[(size_t)r * L.pitch + x];
  __syncwarp();
  const int minTh = g.minTh, iniTh = g.iniTh;
  int nAll = 0, nHi = 0;
  for (int r = 0; r < ah; r++)
    for (int x0 = 0; x0 < aw; x0 += 32) {
      const int x = x0 + lane;
      uint32_t ent = 0;
      if (x < aw) {
        const uint8_t* p = &sc[(r + 1) * CP + x + 1];
        const int M = p[0];
        if (M > minTh) {
          const int nb = max(max(max(p[-CP - 1], p[-CP]), max(p[-CP + 1], p[-1])), max(max(p[1], p[CP - 1]), max(p[CP], p[CP + 1])));
          if (M > nb) ent = (uint32_t)x | ((uint32_t)r << 8) | ((uint32_t)M << 16);
        }
      }
      const unsigned sm = __ballot_sync(0xffffffffu, ent != 0u);
      if (sm) {
        const unsigned hm = __ballot_sync(0xffffffffu, ent != 0u && (int)(ent >> 16) > iniTh);
        if (ent) {
          const int slot = nAll + __popc(sm & ((1u << lane) - 1u));
          if (slot < CLIST) list[slot] = ent;
        }
        nAll += __popc(sm);
        nHi += __popc(hm);
      }
    }
  __syncwarp();
  const bool useHi = nHi > 0;  // the cell has corners at iniThFAST: keep only those (:1132-1139)
  const int nEmit = useHi ? nHi : nAll;
  if (nEmit == 0) return;
  int base = 0;
  if (lane == 0) {
    base = atomicAdd(&candCount[b * kMaxLevels + l], nEmit);
    if (nAll > CLIST || base + nEmit > L.candCap) atomicOr(status, 1);
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  const int nList = min(nAll, CLIST);
  const size_t cbase = (size_t)b * g.totalCandCap + L.candOff;
  const int c = ci * L.nCols + cj;
  int run = 0;
  for (int k0 = 0; k0 < nList; k0 += 32) {
    const int k = k0 + lane;
    uint32_t e = 0;
    bool keep = false;
    if (k < nList) {
      e = list[k];
      keep = !useHi || (int)(e >> 16) > iniTh;
    }
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int slot = base + run + __popc(km & ((1u << lane) - 1u));
      if (slot < L.candCap) {
        const int x = e & 0xff, y = (e >> 8) & 0xff, M = e >> 16;  // detectable-area coordinates
        // ROI coordinates are +3; kp.pt += (j*wCell, i*hCell) (:1150-1151) -> border-relative level coordinates
        candXY[cbase + slot] = (uint32_t)(x + 3 + cj * L.wCell) | ((uint32_t)(y + 3 + ci * L.hCell) << 16);
        candKey[cbase + slot] = ((uint32_t)c << 12) | ((uint32_t)y << 6) | (uint32_t)x;
        candResp[cbase + slot] = (uint8_t)(M - 1);
      }
    }
    run += __popc(km);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

static int encode_level_map(CUtensorMap* tm, uint8_t* base, const LevelGeom& L, uint32_t pyrBytes, int maxBatch, int boxW,
                            int boxH) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B2S_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)maxBatch};
  const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)pyrBytes};
  const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base + L.off, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a %dx%d level, pitch %d", (int)r, L.w, L.h, L.pitch);
    return B2S_ERR_CUDA;
  }
  return B2S_OK;
}

// Called at the end of build_geometry: tensor maps of every level (pyramid load, strength store, blur store) and the
// destination ranges each source tile owns in the bilinear resize.
int tile_build(b2s_extractor* h) {
  ExtractGeom& g = h->geom;
  std::vector<int16_t> tdx, tdy;
  for (int l = 0; l < g.nlevels; l++) {
    LevelGeom& L = g.lv[l];
    L.tilesX = div_up(L.w, TW);
    L.tilesY = div_up(L.h, TH);
    L.tdxOff = (uint32_t)tdx.size();
    L.tdyOff = (uint32_t)tdy.size();
    if (l + 1 < g.nlevels) {
      // same float arithmetic as the resize tables (build_geometry): sx(dx), sy(dy) are non-decreasing
      const int dw = g.lv[l + 1].w, dh = g.lv[l + 1].h;
      const double scale_x = 1. / ((double)dw / L.w), scale_y = 1. / ((double)dh / L.h);
      int dx = 0;
      for (int t = 0; t <= L.tilesX; t++) {
        while (dx < dw) {
          float fx = (float)((dx + 0.5) * scale_x - 0.5);
          int sx = (int)floor((double)fx);
          if (sx < 0) sx = 0;
          if (sx >= L.w - 1) sx = L.w - 1;
          if (sx >= t * TW) break;
          dx++;
        }
        tdx.push_back((int16_t)dx);
      }
      int dy = 0;
      for (int t = 0; t <= L.tilesY; t++) {
        while (dy < dh) {
          float fy = (float)((dy + 0.5) * scale_y - 0.5);
          int sy = (int)floor((double)fy);
          sy = std::min(std::max(sy, 0), L.h - 1);
          if (sy >= t * TH) break;
          dy++;
        }
        tdy.push_back((int16_t)dy);
      }
    } else {
      for (int t = 0; t <= L.tilesX; t++) tdx.push_back(0);
      for (int t = 0; t <= L.tilesY; t++) tdy.push_back(0);
    }
  }
  if (tdx.size() > h->tileTabAlloc || tdy.size() > h->tileTabAlloc) {
    set_error("tile tables exceed their allocation");
    return B2S_ERR_BAD_ARG;
  }
  B2S_CUDA(cudaMemcpy(h->d.tileDx, tdx.data(), tdx.size() * 2, cudaMemcpyHostToDevice));
  B2S_CUDA(cudaMemcpy(h->d.tileDy, tdy.data(), tdy.size() * 2, cudaMemcpyHostToDevice));
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
  for (int l = 0; l < g.nlevels; l++) {
    int rc = encode_level_map(reinterpret_cast<CUtensorMap*>(h->tmPyr[l]), h->d.pyr, g.lv[l], g.pyrBytes, h->maxBatch, BW, BH);
    if (rc == B2S_OK)
      rc = encode_level_map(reinterpret_cast<CUtensorMap*>(h->tmScore[l]), h->d.score, g.lv[l], g.pyrBytes, h->maxBatch, TW, TH);
    if (rc == B2S_OK)
      rc = encode_level_map(reinterpret_cast<CUtensorMap*>(h->tmBlur[l]), h->d.blur, g.lv[l], g.pyrBytes, h->maxBatch, TW, TH);
    if (rc != B2S_OK) return rc;
  }
  B2S_CUDA(cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(TileSmem) + 128)));
  return B2S_OK;
}

// Enqueue the fused front end for `batch` images starting at image `bBase` of the handle's buffers.
// path: 1 = strength map only (pyramid and blur by the per-stage kernels), 2 = + blur, 3 = + pyramid (everything fused).
int tile_run(b2s_extractor* h, int bBase, int batch, int path, cudaStream_t st, cudaEvent_t evAfterTiles) {
  const ExtractGeom& g = h->geom;
  const DeviceBuffers& d = h->d;
  for (int l = 0; l < g.nlevels; l++) {
    TileArgs a;
    a.level = l;
    a.bBase = bBase;
    a.doFast = 1;
    a.doBlur = path >= 2;
    a.doResize = path >= 3;
    a.pyr = d.pyr;
    a.rxOfs = d.rxOfs;
    a.rxAlpha = d.rxAlpha;
    a.ryOfs = d.ryOfs;
    a.ryBeta = d.ryBeta;
    a.tileDx = d.tileDx;
    a.tileDy = d.tileDy;
    k_tile<<<dim3(g.lv[l].tilesX, g.lv[l].tilesY, batch), TILE_THREADS, sizeof(TileSmem) + 128, st>>>(
        *reinterpret_cast<const CUtensorMap*>(h->tmPyr[l]), *reinterpret_cast<const CUtensorMap*>(h->tmScore[l]),
        *reinterpret_cast<const CUtensorMap*>(h->tmBlur[l]), g, a);
    h->launches++;
  }
  if (evAfterTiles) cudaEventRecord(evAfterTiles, st);
  k_cells<<<dim3(div_up(g.totalCells, 4), batch), 128, 0, st>>>(
      g, d.score + (size_t)bBase * g.pyrBytes, d.cellInfo, d.candXY + (size_t)bBase * g.totalCandCap,
      d.candKey + (size_t)bBase * g.totalCandCap, d.candResp + (size_t)bBase * g.totalCandCap,
      d.candCount + (size_t)bBase * kMaxLevels, d.status);
  h->launches++;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

}  // namespace b2s
