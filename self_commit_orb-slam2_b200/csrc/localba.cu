// localba.cu — B200 (sm_100a) Optimizer::LocalBundleAdjustment numerics (FP64), batched over independent windows.
//
// Replaces /root/reference/src/Optimizer.cc:698-996 from graph construction to the write-back values, i.e. the g2o
// functions it executes (Thirdparty/g2o/g2o/...):
//   types/types_six_dof_expmap.{h,cpp}   computeError / linearizeOplus of EdgeSE3ProjectXYZ, EdgeStereoSE3ProjectXYZ
//   core/base_binary_edge.hpp:55-120     constructQuadraticForm (Huber weighted)      -> k_build_landmarks / k_build_poses
//   core/block_solver.hpp:354-486        Schur complement + back substitution         -> k_schur / k_chol / k_backsub
//   core/optimization_algorithm_levenberg.cpp:61-189   LM control                     -> k_control_begin / k_control_end
//   types/se3quat.h:223-257, types_sba.h:52-56         oplus updates                  -> k_update_poses / k_backsub
// The LM state machine lives on the device (one record per window); the host only polls one "any window active" word
// per trial, so a batch of windows advances in lock step without host-side per-window logic.
#include <float.h>
#include <stddef.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b2s {

constexpr int PSTRIDE = 8;  // doubles per pose: q(x,y,z,w), t(x,y,z), pad
constexpr int CHOL_BS = 32;

struct BaState {
  int active, round, it, qmax, needBuild, restore, doOutlier, solveOk, stop, nTrials, nBad, robust, finishedRound0;
  int its[2];
  double lambda, ni, currentChi, iniChi, tempChi, rho, chi2Final;
  unsigned long long maxDiagBits;
  int trace[256];
};

struct BaWin {  // per-window sizes (device copy)
  int nKf, nLocal, nMp, nEdges, nFree;
  float fx, fy, cx, cy, bf;
};

struct BaPtrs {  // strided per-window arrays
  int capKf, capMp, capE, ldS, nPartE, nPartM;
  BaWin* win;
  BaState* st;
  double *pose, *poseBak, *pts, *ptsBak;
  int* poseIndex;
  int *eKf, *eMp;
  float *eObs, *eW;
  uint8_t *eStereo, *eLevel, *eOutlier;
  double *err, *chi2, *W;
  int *mpStart, *mpEdges, *kfStart, *kfEdges;
  double *Hpp, *Hll, *b, *x, *Dinv, *S;
  double *db, *Y;      // Dinv*b_l per landmark, W*Dinv per edge
  int *lmEdge, *freeKf;  // [landmark][free pose] -> edge id or -1 ; free pose index -> keyframe index
  double *partChi, *partScale;
  int* anyActive;
};

// ---------------------------------------------------------------- small FP64 helpers
__device__ __forceinline__ void quat_rotate(const double* q, const double* v, double* o) {
  double uvx = q[1] * v[2] - q[2] * v[1], uvy = q[2] * v[0] - q[0] * v[2], uvz = q[0] * v[1] - q[1] * v[0];
  uvx += uvx; uvy += uvy; uvz += uvz;
  o[0] = v[0] + q[3] * uvx + (q[1] * uvz - q[2] * uvy);
  o[1] = v[1] + q[3] * uvy + (q[2] * uvx - q[0] * uvz);
  o[2] = v[2] + q[3] * uvz + (q[0] * uvy - q[1] * uvx);
}
__device__ __forceinline__ void pose_map(const double* P, const double* X, double* o) {
  quat_rotate(P, X, o);
  o[0] += P[4]; o[1] += P[5]; o[2] += P[6];
}
__device__ __forceinline__ void quat_to_R(const double* q, double R[3][3]) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
}
__device__ __forceinline__ void quat_from_R(const double m[3][3], double* q) {  // Eigen::Quaterniond(Matrix3d)
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[2][1] - m[1][2]) * t;
    q[1] = (m[0][2] - m[2][0]) * t;
    q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    v[j] = (m[j][i] + m[i][j]) * t;
    v[k] = (m[k][i] + m[i][k]) * t;
    q[0] = v[0]; q[1] = v[1]; q[2] = v[2];
  }
}
__device__ __forceinline__ void normalize_rot(double* q) {  // SE3Quat::normalizeRotation
  if (q[3] < 0) {
    q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3];
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// SE3Quat::exp(update) * estimate  (se3quat.h:223-257, :103-109)
__device__ void pose_oplus(double* P, const double* upd) {
  const double wx = upd[0], wy = upd[1], wz = upd[2];
  const double theta = sqrt(wx * wx + wy * wy + wz * wz);
  const double O[3][3] = {{0, -wz, wy}, {wz, 0, -wx}, {-wy, wx, 0}};
  double O2[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) O2[i][j] = O[i][0] * O[0][j] + O[i][1] * O[1][j] + O[i][2] * O[2][j];
  double R[3][3], V[3][3];
  if (theta < 0.00001) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        R[i][j] = (i == j ? 1.0 : 0.0) + O[i][j] + O2[i][j];
        V[i][j] = R[i][j];
      }
  } else {
    const double a = sin(theta) / theta;
    const double bb = (1 - cos(theta)) / (theta * theta);
    const double c = (theta - sin(theta)) / (theta * theta * theta);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        R[i][j] = (i == j ? 1.0 : 0.0) + a * O[i][j] + bb * O2[i][j];
        V[i][j] = (i == j ? 1.0 : 0.0) + bb * O[i][j] + c * O2[i][j];
      }
  }
  double eq[4], et[3];
  quat_from_R(R, eq);
  normalize_rot(eq);
  for (int i = 0; i < 3; i++) et[i] = V[i][0] * upd[3] + V[i][1] * upd[4] + V[i][2] * upd[5];
  double rt[3];
  quat_rotate(eq, P + 4, rt);
  const double a0 = eq[0], a1 = eq[1], a2 = eq[2], a3 = eq[3];
  const double b0 = P[0], b1 = P[1], b2 = P[2], b3 = P[3];
  double r[4];
  r[3] = a3 * b3 - a0 * b0 - a1 * b1 - a2 * b2;
  r[0] = a3 * b0 + a0 * b3 + a1 * b2 - a2 * b1;
  r[1] = a3 * b1 + a1 * b3 + a2 * b0 - a0 * b2;
  r[2] = a3 * b2 + a2 * b3 + a0 * b1 - a1 * b0;
  normalize_rot(r);
  P[0] = r[0]; P[1] = r[1]; P[2] = r[2]; P[3] = r[3];
  P[4] = et[0] + rt[0]; P[5] = et[1] + rt[1]; P[6] = et[2] + rt[2];
}

__device__ __forceinline__ void huber(double e, double delta, double& rho0, double& rho1) {
  const double dsqr = delta * delta;
  if (e <= dsqr) {
    rho0 = e;
    rho1 = 1.;
  } else {
    const double s = sqrt(e);
    rho0 = 2 * s * delta - dsqr;
    rho1 = delta / s;
  }
}
__device__ __forceinline__ double delta_of(bool stereo) {
  // const float thHuberMono = sqrt(5.991), thHuberStereo = sqrt(7.815)  (src/Optimizer.cc:764-765)
  return stereo ? (double)(float)2.795532150593156 : (double)(float)2.4476519360399226;
}

__device__ double block_sum(double v, double* sm) {  // deterministic block reduction (blockDim multiple of 32, <=1024)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; i++) r += sm[i];
  }
  return r;  // valid in thread 0
}

// ---------------------------------------------------------------- kernels
// computeActiveErrors + activeRobustChi2 (sparse_optimizer.cpp:61-113).  mode 0: iteration begin, 1: after a trial.
__global__ void __launch_bounds__(256) k_errors(BaPtrs p, int mode) {
  __shared__ double sm[32];
  const int w = blockIdx.y;
  const BaState& st = p.st[w];
  if (!st.active) return;
  if (mode == 0 && !st.needBuild) return;
  const BaWin W = p.win[w];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double contrib = 0;
  if (e < W.nEdges) {
    const size_t eo = (size_t)w * p.capE + e;
    if (!p.eLevel[eo]) {
      const int kf = p.eKf[eo], mp = p.eMp[eo];
      double Xc[3];
      pose_map(p.pose + ((size_t)w * p.capKf + kf) * PSTRIDE, p.pts + ((size_t)w * p.capMp + mp) * 3, Xc);
      const double wgt = (double)p.eW[eo];
      const float* ob = p.eObs + eo * 3;
      double e0, e1, e2 = 0, c2;
      if (p.eStereo[eo]) {
        const float invz = (float)(1.0 / Xc[2]);  // const float invz = 1.0f/trans_xyz[2]  (.cpp:151)
        const double u = Xc[0] * invz * (double)W.fx + (double)W.cx;
        const double v = Xc[1] * invz * (double)W.fy + (double)W.cy;
        const double ur = u - (double)__fmul_rn(W.bf, invz);
        e0 = (double)ob[0] - u;
        e1 = (double)ob[1] - v;
        e2 = (double)ob[2] - ur;
        c2 = e0 * (wgt * e0) + e1 * (wgt * e1) + e2 * (wgt * e2);
      } else {
        const double u = Xc[0] / Xc[2] * (double)W.fx + (double)W.cx;
        const double v = Xc[1] / Xc[2] * (double)W.fy + (double)W.cy;
        e0 = (double)ob[0] - u;
        e1 = (double)ob[1] - v;
        c2 = e0 * (wgt * e0) + e1 * (wgt * e1);
      }
      p.err[eo * 3] = e0; p.err[eo * 3 + 1] = e1; p.err[eo * 3 + 2] = e2;
      p.chi2[eo] = c2;
      if (st.robust) {
        double r0, r1;
        huber(c2, delta_of(p.eStereo[eo]), r0, r1);
        contrib = r0;
      } else
        contrib = c2;
    }
  }
  const double s = block_sum(contrib, sm);
  if (threadIdx.x == 0) p.partChi[(size_t)w * p.nPartE + blockIdx.x] = s;
}

struct EdgeJac {
  double A[3][3], B[3][6];
  int D;
};
// linearizeOplus (types_six_dof_expmap.cpp:103-139 mono, :188-234 stereo)
__device__ __forceinline__ void edge_jacobians(const double* P, const double* X, bool st, double fx, double fy, double bf,
                                               EdgeJac& J) {
  double Xc[3], R[3][3];
  pose_map(P, X, Xc);
  quat_to_R(P, R);
  const double x = Xc[0], y = Xc[1], z = Xc[2], z_2 = z * z;
  J.D = st ? 3 : 2;
  if (st) {
    for (int c = 0; c < 3; c++) {
      J.A[0][c] = -fx * R[0][c] / z + fx * x * R[2][c] / z_2;
      J.A[1][c] = -fy * R[1][c] / z + fy * y * R[2][c] / z_2;
      J.A[2][c] = J.A[0][c] - bf * R[2][c] / z_2;
    }
  } else {
    const double tmp[2][3] = {{fx, 0, -x / z * fx}, {0, fy, -y / z * fy}};
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (-1. / z * tmp[r][k]) * R[k][c];
        J.A[r][c] = s;
      }
    J.A[2][0] = J.A[2][1] = J.A[2][2] = 0;
  }
  J.B[0][0] = x * y / z_2 * fx;
  J.B[0][1] = -(1 + (x * x / z_2)) * fx;
  J.B[0][2] = y / z * fx;
  J.B[0][3] = -1. / z * fx;
  J.B[0][4] = 0;
  J.B[0][5] = x / z_2 * fx;
  J.B[1][0] = (1 + y * y / z_2) * fy;
  J.B[1][1] = -x * y / z_2 * fy;
  J.B[1][2] = -x / z * fy;
  J.B[1][3] = 0;
  J.B[1][4] = -1. / z * fy;
  J.B[1][5] = y / z_2 * fy;
  if (st) {
    J.B[2][0] = J.B[0][0] - bf * y / z_2;
    J.B[2][1] = J.B[0][1] + bf * x / z_2;
    J.B[2][2] = J.B[0][2];
    J.B[2][3] = J.B[0][3];
    J.B[2][4] = 0;
    J.B[2][5] = J.B[0][5] - bf / z_2;
  } else {
    for (int c = 0; c < 6; c++) J.B[2][c] = 0;
  }
}

__device__ __forceinline__ void atomic_max_pos_double(unsigned long long* addr, double v) {
  atomicMax(addr, (unsigned long long)__double_as_longlong(fabs(v)));  // |v| >= 0: IEEE order == integer order
}

// buildSystem, landmark side: one thread per landmark walks its edges in insertion order (deterministic sums)
__global__ void __launch_bounds__(128) k_build_landmarks(BaPtrs p) {
  const int w = blockIdx.y;
  const BaState& st = p.st[w];
  if (!st.active || !st.needBuild) return;
  const BaWin W = p.win[w];
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= W.nMp) return;
  const size_t mo = (size_t)w * p.capMp + l;
  const int* ms = p.mpStart + (size_t)w * (p.capMp + 1);
  const int* me = p.mpEdges + (size_t)w * p.capE;
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
  const double* X = p.pts + mo * 3;
  for (int k = ms[l]; k < ms[l + 1]; k++) {
    const int e = me[k];
    const size_t eo = (size_t)w * p.capE + e;
    if (p.eLevel[eo]) continue;
    const int kf = p.eKf[eo];
    const bool stereo = p.eStereo[eo];
    EdgeJac J;
    edge_jacobians(p.pose + ((size_t)w * p.capKf + kf) * PSTRIDE, X, stereo, W.fx, W.fy, W.bf, J);
    const double w0 = (double)p.eW[eo];
    double rho1 = 1.0;
    if (st.robust) {
      double r0;
      huber(p.chi2[eo], delta_of(stereo), r0, rho1);
    }
    const double* er = p.err + eo * 3;
    double omr[3];
    for (int r = 0; r < 3; r++) omr[r] = -(w0 * er[r]) * rho1;
    const double wq = rho1 * w0;
    for (int i = 0; i < 3; i++) {
      double s = 0;
      for (int r = 0; r < J.D; r++) s += J.A[r][i] * omr[r];
      bl[i] += s;
      for (int j = 0; j < 3; j++) {
        double hh = 0;
        for (int r = 0; r < J.D; r++) hh += J.A[r][i] * wq * J.A[r][j];
        H[i * 3 + j] += hh;
      }
    }
    if (p.poseIndex[(size_t)w * p.capKf + kf] >= 0) {
      double* Wb = p.W + eo * 18;
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 3; j++) {
          double hh = 0;
          for (int r = 0; r < J.D; r++) hh += J.B[r][i] * wq * J.A[r][j];
          Wb[i * 3 + j] = hh;
        }
    }
  }
  for (int k = 0; k < 9; k++) p.Hll[mo * 9 + k] = H[k];
  double* bb = p.b + (size_t)w * (p.capKf * 6 + p.capMp * 3) + (size_t)W.nFree * 6 + (size_t)l * 3;
  bb[0] = bl[0]; bb[1] = bl[1]; bb[2] = bl[2];
  atomic_max_pos_double(&p.st[w].maxDiagBits, fmax(fabs(H[0]), fmax(fabs(H[4]), fabs(H[8]))));
}

// buildSystem, pose side: one CTA per free pose; fixed thread->edge mapping + ordered reduction => deterministic
__global__ void __launch_bounds__(128) k_build_poses(BaPtrs p) {
  __shared__ double sm[32];
  __shared__ double out[27];
  const int w = blockIdx.y;
  const BaState& st = p.st[w];
  if (!st.active || !st.needBuild) return;
  const BaWin W = p.win[w];
  const int kf = blockIdx.x;
  if (kf >= W.nKf) return;
  const int pi = p.poseIndex[(size_t)w * p.capKf + kf];
  if (pi < 0) return;
  const int* ks = p.kfStart + (size_t)w * (p.capKf + 1);
  const int* ke = p.kfEdges + (size_t)w * p.capE;
  double acc[27];
#pragma unroll
  for (int k = 0; k < 27; k++) acc[k] = 0;
  const double* P = p.pose + ((size_t)w * p.capKf + kf) * PSTRIDE;
  for (int k = ks[kf] + threadIdx.x; k < ks[kf + 1]; k += blockDim.x) {
    const int e = ke[k];
    const size_t eo = (size_t)w * p.capE + e;
    if (p.eLevel[eo]) continue;
    const bool stereo = p.eStereo[eo];
    EdgeJac J;
    edge_jacobians(P, p.pts + ((size_t)w * p.capMp + p.eMp[eo]) * 3, stereo, W.fx, W.fy, W.bf, J);
    const double w0 = (double)p.eW[eo];
    double rho1 = 1.0;
    if (st.robust) {
      double r0;
      huber(p.chi2[eo], delta_of(stereo), r0, rho1);
    }
    const double* er = p.err + eo * 3;
    double omr[3];
    for (int r = 0; r < 3; r++) omr[r] = -(w0 * er[r]) * rho1;
    const double wq = rho1 * w0;
    int t = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
      for (int j = i; j < 6; j++) {
        double hh = 0;
        for (int r = 0; r < J.D; r++) hh += J.B[r][i] * wq * J.B[r][j];
        acc[t++] += hh;
      }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      double s = 0;
      for (int r = 0; r < J.D; r++) s += J.B[r][i] * omr[r];
      acc[21 + i] += s;
    }
  }
#pragma unroll
  for (int k = 0; k < 27; k++) {
    const double s = block_sum(acc[k], sm);
    if (threadIdx.x == 0) out[k] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double* Hp = p.Hpp + ((size_t)w * p.capKf + pi) * 36;
    int t = 0;
    double md = 0;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) {
        Hp[i * 6 + j] = out[t];
        Hp[j * 6 + i] = out[t];
        if (i == j) md = fmax(md, fabs(out[t]));
        t++;
      }
    double* bp = p.b + (size_t)w * (p.capKf * 6 + p.capMp * 3) + (size_t)pi * 6;
    for (int i = 0; i < 6; i++) bp[i] = out[21 + i];
    atomic_max_pos_double(&p.st[w].maxDiagBits, md);
  }
}

__global__ void k_control_begin(BaPtrs p, int batch) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= batch) return;
  BaState& st = p.st[w];
  if (!st.active) return;
  if (st.needBuild) {
    const BaWin W = p.win[w];
    const int nb = (W.nEdges + 255) / 256;
    double chi = 0;
    for (int k = 0; k < nb; k++) chi += p.partChi[(size_t)w * p.nPartE + k];
    st.currentChi = chi;
    st.iniChi = chi;
    st.tempChi = chi;
    if (st.it == 0) {  // computeLambdaInit (levenberg.cpp:93-97,166-180)
      st.lambda = 1e-5 * __longlong_as_double((long long)st.maxDiagBits);
      st.ni = 2;
      st.nBad = 0;
    }
    st.rho = 0;
    st.qmax = 0;
    st.needBuild = 0;
  }
  st.maxDiagBits = 0ull;
}

// ---- Schur complement (block_solver.hpp:381-439), atomic-free and deterministic:
//   k_lm_edge      (once per window)  lmEdge[landmark][free pose] = edge id
//   k_dinv         per landmark:  Dinv = (Hll + lambda I)^-1,  db = Dinv b_l
//   k_schur_pose   per free pose: Y_e = W_e Dinv (kept for the block pass), augmented row  b_p - sum_e W_e db
//   k_schur_blocks per lower block (i1 >= i2): S(i1,i2) = [Hpp + lambda I] - sum_{l seen by both} Y_a W_c^T
__global__ void __launch_bounds__(256) k_lm_edge(BaPtrs p) {
  const int w = blockIdx.y;
  const BaWin W = p.win[w];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.nEdges) return;
  const size_t eo = (size_t)w * p.capE + e;
  const int pi = p.poseIndex[(size_t)w * p.capKf + p.eKf[eo]];
  if (pi >= 0) p.lmEdge[((size_t)w * p.capMp + p.eMp[eo]) * p.capKf + pi] = e;
}

__global__ void __launch_bounds__(128) k_dinv(BaPtrs p) {
  const int w = blockIdx.y;
  const BaState& st = p.st[w];
  if (!st.active) return;
  const BaWin W = p.win[w];
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= W.nMp) return;
  const size_t mo = (size_t)w * p.capMp + l;
  double D[9];
#pragma unroll
  for (int k = 0; k < 9; k++) D[k] = p.Hll[mo * 9 + k];
  D[0] += st.lambda; D[4] += st.lambda; D[8] += st.lambda;
  double Di[9];
  const double c00 = D[4] * D[8] - D[5] * D[7], c01 = D[5] * D[6] - D[3] * D[8], c02 = D[3] * D[7] - D[4] * D[6];
  const double id = 1.0 / (D[0] * c00 + D[1] * c01 + D[2] * c02);
  Di[0] = c00 * id; Di[1] = (D[2] * D[7] - D[1] * D[8]) * id; Di[2] = (D[1] * D[5] - D[2] * D[4]) * id;
  Di[3] = c01 * id; Di[4] = (D[0] * D[8] - D[2] * D[6]) * id; Di[5] = (D[2] * D[3] - D[0] * D[5]) * id;
  Di[6] = c02 * id; Di[7] = (D[1] * D[6] - D[0] * D[7]) * id; Di[8] = (D[0] * D[4] - D[1] * D[3]) * id;
#pragma unroll
  for (int k = 0; k < 9; k++) p.Dinv[mo * 9 + k] = Di[k];
  const double* bl = p.b + (size_t)w * (p.capKf * 6 + p.capMp * 3) + (size_t)W.nFree * 6 + (size_t)l * 3;
  p.db[mo * 3 + 0] = Di[0] * bl[0] + Di[1] * bl[1] + Di[2] * bl[2];
  p.db[mo * 3 + 1] = Di[3] * bl[0] + Di[4] * bl[1] + Di[5] * bl[2];
  p.db[mo * 3 + 2] = Di[6] * bl[0] + Di[7] * bl[1] + Di[8] * bl[2];
}

__global__ void __launch_bounds__(128) k_schur_pose(BaPtrs p) {
  __shared__ double sm[32];
  __shared__ double out[6];
  const int w = blockIdx.y;
  const BaState& st = p.st[w];
  if (!st.active) return;
  const BaWin W = p.win[w];
  const int kf = blockIdx.x;
  if (kf >= W.nKf) return;
  const int pi = p.poseIndex[(size_t)w * p.capKf + kf];
  if (pi < 0) return;
  const int* ks = p.kfStart + (size_t)w * (p.capKf + 1);
  const int* ke = p.kfEdges + (size_t)w * p.capE;
  double r[6] = {0, 0, 0, 0, 0, 0};
  for (int k = ks[kf] + threadIdx.x; k < ks[kf + 1]; k += blockDim.x) {
    const int e = ke[k];
    const size_t eo = (size_t)w * p.capE + e;
    if (p.eLevel[eo]) continue;
    const size_t mo = (size_t)w * p.capMp + p.eMp[eo];
    const double* Di = p.Dinv + mo * 9;
    const double* d3 = p.db + mo * 3;
    const double* Wb = p.W + eo * 18;
    double* Yb = p.Y + eo * 18;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const double w0 = Wb[i * 3], w1 = Wb[i * 3 + 1], w2 = Wb[i * 3 + 2];
      Yb[i * 3 + 0] = w0 * Di[0] + w1 * Di[3] + w2 * Di[6];
      Yb[i * 3 + 1] = w0 * Di[1] + w1 * Di[4] + w2 * Di[7];
      Yb[i * 3 + 2] = w0 * Di[2] + w1 * Di[5] + w2 * Di[8];
      r[i] += w0 * d3[0] + w1 * d3[1] + w2 * d3[2];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const double s = block_sum(r[i], sm);
    if (threadIdx.x == 0) out[i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const int n = W.nFree * 6;
    const double* bp = p.b + (size_t)w * (p.capKf * 6 + p.capMp * 3) + (size_t)pi * 6;
    p.S[(size_t)w * p.ldS * p.ldS + (size_t)n * p.ldS + pi * 6 + threadIdx.x] = bp[threadIdx.x] - out[threadIdx.x];
  }
}

__global__ void __launch_bounds__(64) k_schur_blocks(BaPtrs p) {
  __shared__ double red[2][36];
  const int w = blockIdx.y;
  const BaState& st = p.st[w];
  if (!st.active) return;
  const BaWin W = p.win[w];
  // decode lower-triangular block index -> (i1 >= i2)
  const int t = blockIdx.x;
  int i1 = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((i1 + 1) * (i1 + 2) / 2 <= t) i1++;
  while (i1 * (i1 + 1) / 2 > t) i1--;
  const int i2 = t - i1 * (i1 + 1) / 2;
  if (i1 >= W.nFree) return;
  const int kf1 = p.freeKf[(size_t)w * p.capKf + i1];
  const int* ks = p.kfStart + (size_t)w * (p.capKf + 1);
  const int* ke = p.kfEdges + (size_t)w * p.capE;
  const int* lm = p.lmEdge + (size_t)w * p.capMp * p.capKf;
  double acc[36];
#pragma unroll
  for (int k = 0; k < 36; k++) acc[k] = 0;
  for (int k = ks[kf1] + threadIdx.x; k < ks[kf1 + 1]; k += blockDim.x) {
    const int a = ke[k];
    const size_t eoa = (size_t)w * p.capE + a;
    if (p.eLevel[eoa]) continue;
    const int c = lm[(size_t)p.eMp[eoa] * p.capKf + i2];
    if (c < 0) continue;
    const size_t eoc = (size_t)w * p.capE + c;
    if (p.eLevel[eoc]) continue;
    const double* Ya = p.Y + eoa * 18;
    const double* Wc = p.W + eoc * 18;
    double y[18], wc[18];
#pragma unroll
    for (int q = 0; q < 18; q++) {
      y[q] = Ya[q];
      wc[q] = Wc[q];
    }
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int cc = 0; cc < 6; cc++)
        acc[r * 6 + cc] += y[r * 3] * wc[cc * 3] + y[r * 3 + 1] * wc[cc * 3 + 1] + y[r * 3 + 2] * wc[cc * 3 + 2];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 36; k++) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 36) {
    const int r = threadIdx.x / 6, cc = threadIdx.x - r * 6;
    double base = 0;
    if (i1 == i2) {
      base = p.Hpp[((size_t)w * p.capKf + i1) * 36 + r * 6 + cc];
      if (r == cc) base += st.lambda;
    }
    p.S[(size_t)w * p.ldS * p.ldS + (size_t)(i1 * 6 + r) * p.ldS + (i2 * 6 + cc)] = base - (red[0][threadIdx.x] + red[1][threadIdx.x]);
  }
}

// LinearSolver on the reduced camera system: blocked right-looking Cholesky (lower), one CTA per window.
// The augmented last row carries b and ends up holding y = L^-1 b; then L^T x = y by blocked back substitution.
// Diagonal 32x32 blocks are factored by one warp with a row per lane in registers (shuffle broadcast), the panel solve
// keeps each row in registers, the trailing update is 4x4 register tiled out of the shared-memory panel.
constexpr int CHOL_T = 512;
constexpr int CHOL_PP = CHOL_BS + 1;
__global__ void __launch_bounds__(CHOL_T) k_chol(BaPtrs p) {
  extern __shared__ __align__(16) double dsm[];
  const int w = blockIdx.x;
  BaState& st = p.st[w];
  if (!st.active) return;
  const BaWin W = p.win[w];
  const int n = W.nFree * 6, N1 = n + 1, ld = p.ldS;
  double* S = p.S + (size_t)w * ld * ld;
  double* Dblk = dsm;                                // 32 x 33
  double* panel = dsm + CHOL_BS * CHOL_PP;           // ld x 33
  double* xs = panel + (size_t)(ld + 4) * CHOL_PP;   // ld
  __shared__ int fail;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
  if (tid == 0) fail = 0;
  __syncthreads();
  for (int kb = 0; kb < n; kb += CHOL_BS) {
    const int wd = min(CHOL_BS, n - kb);
    for (int idx = tid; idx < CHOL_BS * CHOL_BS; idx += T) {
      const int r = idx >> 5, c = idx & 31;
      double v = (r == c) ? 1.0 : 0.0;  // identity padding for a short last block
      if (r < wd && c < wd) v = (c <= r) ? S[(size_t)(kb + r) * ld + kb + c] : 0.0;
      Dblk[r * CHOL_PP + c] = v;
    }
    __syncthreads();
    if (tid < 32) {
      double row[CHOL_BS];
#pragma unroll
      for (int c = 0; c < CHOL_BS; c++) row[c] = Dblk[lane * CHOL_PP + c];
      bool bad = false;
#pragma unroll
      for (int j = 0; j < CHOL_BS; j++) {
        double djj = __shfl_sync(0xffffffffu, row[j], j);
        if (!(djj > 0.0) || !isfinite(djj)) {
          bad = true;
          djj = 1.0;
        }
        const double ljj = sqrt(djj);
        if (lane == j) row[j] = ljj;
        else if (lane > j) row[j] = row[j] / ljj;
#pragma unroll
        for (int k = j + 1; k < CHOL_BS; k++) {
          const double lkj = __shfl_sync(0xffffffffu, row[j], k);  // L[k][j]
          if (lane >= k) row[k] -= row[j] * lkj;
        }
      }
      if (bad && lane == 0) fail = 1;
#pragma unroll
      for (int c = 0; c < CHOL_BS; c++) Dblk[lane * CHOL_PP + c] = (c <= lane) ? row[c] : 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < wd * wd; idx += T) {
      const int r = idx / wd, c = idx - r * wd;
      if (c <= r) S[(size_t)(kb + r) * ld + kb + c] = Dblk[r * CHOL_PP + c];
    }
    const int m = N1 - (kb + wd);  // rows below the diagonal block (including the augmented row)
    for (int rowi = tid; rowi < m; rowi += T) {
      const int i = kb + wd + rowi;
      double v[CHOL_BS];
#pragma unroll
      for (int c = 0; c < CHOL_BS; c++) v[c] = (c < wd) ? S[(size_t)i * ld + kb + c] : 0.0;
#pragma unroll
      for (int c = 0; c < CHOL_BS; c++) {
        double a = v[c];
#pragma unroll
        for (int k = 0; k < c; k++) a -= v[k] * Dblk[c * CHOL_PP + k];
        v[c] = a / Dblk[c * CHOL_PP + c];
      }
#pragma unroll
      for (int c = 0; c < CHOL_BS; c++) {
        panel[rowi * CHOL_PP + c] = v[c];
        if (c < wd) S[(size_t)i * ld + kb + c] = v[c];
      }
    }
    // zero the padding rows read by the 4x4 tiles
    for (int idx = tid; idx < 4 * CHOL_PP; idx += T) panel[(m + idx / CHOL_PP) * CHOL_PP + idx % CHOL_PP] = 0.0;
    __syncthreads();
    const int mt = (m + 3) >> 2;
    const int ntile = mt * (mt + 1) / 2;
    for (int t = tid; t < ntile; t += T) {
      int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
      while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
      while (ti * (ti + 1) / 2 > t) ti--;
      const int tj = t - ti * (ti + 1) / 2;
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b2 = 0; b2 < 4; b2++) acc[a][b2] = 0;
      const double* pa = panel + (size_t)(4 * ti) * CHOL_PP;
      const double* pb = panel + (size_t)(4 * tj) * CHOL_PP;
#pragma unroll 8
      for (int k = 0; k < CHOL_BS; k++) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
          av[a] = pa[a * CHOL_PP + k];
          bv[a] = pb[a * CHOL_PP + k];
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b2 = 0; b2 < 4; b2++) acc[a][b2] += av[a] * bv[b2];
      }
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int i = 4 * ti + a;
        if (i >= m) continue;
#pragma unroll
        for (int b2 = 0; b2 < 4; b2++) {
          const int j = 4 * tj + b2;
          if (j > i || kb + wd + j >= n) continue;  // lower triangle only; column n is never needed
          S[(size_t)(kb + wd + i) * ld + kb + wd + j] -= acc[a][b2];
        }
      }
    }
    __syncthreads();
  }
  // back substitution L^T x = y
  for (int i = tid; i < n; i += T) xs[i] = S[(size_t)n * ld + i];
  __syncthreads();
  const int nblk = (n + CHOL_BS - 1) / CHOL_BS;
  for (int bk = nblk - 1; bk >= 0; bk--) {
    const int kb = bk * CHOL_BS, wd = min(CHOL_BS, n - kb);
    for (int idx = tid; idx < CHOL_BS * CHOL_BS; idx += T) {
      const int r = idx >> 5, c = idx & 31;
      double v = (r == c) ? 1.0 : 0.0;
      if (r < wd && c < wd) v = (c <= r) ? S[(size_t)(kb + r) * ld + kb + c] : 0.0;
      Dblk[r * CHOL_PP + c] = v;
    }
    __syncthreads();
    if (tid < 32) {
      double col[CHOL_BS];  // column `lane` of L_D == row `lane` of L_D^T
#pragma unroll
      for (int r = 0; r < CHOL_BS; r++) col[r] = Dblk[r * CHOL_PP + lane];
      double tv = (lane < wd) ? xs[kb + lane] : 0.0;
#pragma unroll
      for (int j = CHOL_BS - 1; j >= 0; j--) {
        double xj = tv / col[j];                      // meaningful on lane j (col[j] = L[j][j])
        xj = __shfl_sync(0xffffffffu, xj, j);
        if (lane == j) tv = xj;
        else if (lane < j) tv -= col[j] * xj;         // L[j][lane] * x_j
      }
      if (lane < wd) xs[kb + lane] = tv;
    }
    __syncthreads();
    for (int c = tid; c < kb; c += T) {
      double sacc = xs[c];
      double lv[CHOL_BS];
#pragma unroll
      for (int r = 0; r < CHOL_BS; r++) lv[r] = (r < wd) ? S[(size_t)(kb + r) * ld + c] : 0.0;
#pragma unroll
      for (int r = 0; r < CHOL_BS; r++) sacc -= lv[r] * xs[kb + min(r, wd - 1)] * (r < wd ? 1.0 : 0.0);
      xs[c] = sacc;
    }
    __syncthreads();
  }
  double* x = p.x + (size_t)w * (p.capKf * 6 + p.capMp * 3);
  for (int i = tid; i < n; i += T) x[i] = xs[i];
  if (tid == 0) st.solveOk = fail ? 0 : 1;
}

// landmark back-substitution (block_solver.hpp:461-481) + update (types_sba.h:52-56) + computeScale partials
__global__ void __launch_bounds__(128) k_backsub(BaPtrs p) {
  __shared__ double sm[32];
  const int w = blockIdx.y;
  const BaState& st = p.st[w];
  if (!st.active) return;
  const BaWin W = p.win[w];
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  double sc = 0;
  if (l < W.nMp && st.solveOk) {
    const size_t mo = (size_t)w * p.capMp + l;
    const int* ms = p.mpStart + (size_t)w * (p.capMp + 1);
    const int* me = p.mpEdges + (size_t)w * p.capE;
    const size_t xo = (size_t)w * (p.capKf * 6 + p.capMp * 3);
    const double* bl = p.b + xo + (size_t)W.nFree * 6 + (size_t)l * 3;
    double cl[3] = {bl[0], bl[1], bl[2]};
    for (int k = ms[l]; k < ms[l + 1]; k++) {
      const int e = me[k];
      const size_t eo = (size_t)w * p.capE + e;
      if (p.eLevel[eo]) continue;
      const int pi = p.poseIndex[(size_t)w * p.capKf + p.eKf[eo]];
      if (pi < 0) continue;
      const double* Wb = p.W + eo * 18;
      const double* xp = p.x + xo + (size_t)pi * 6;
      for (int c = 0; c < 3; c++)
        for (int r = 0; r < 6; r++) cl[c] -= Wb[r * 3 + c] * xp[r];
    }
    const double* Di = p.Dinv + mo * 9;
    double xl[3];
    for (int i = 0; i < 3; i++) xl[i] = Di[i * 3] * cl[0] + Di[i * 3 + 1] * cl[1] + Di[i * 3 + 2] * cl[2];
    double* xout = p.x + xo + (size_t)W.nFree * 6 + (size_t)l * 3;
    double* X = p.pts + mo * 3;
    double* Xb = p.ptsBak + mo * 3;
    for (int i = 0; i < 3; i++) {
      xout[i] = xl[i];
      Xb[i] = X[i];
      X[i] += xl[i];
      sc += xl[i] * (st.lambda * xl[i] + bl[i]);
    }
  } else if (l < W.nMp) {
    const size_t mo = (size_t)w * p.capMp + l;
    for (int i = 0; i < 3; i++) p.ptsBak[mo * 3 + i] = p.pts[mo * 3 + i];
  }
  const double s = block_sum(sc, sm);
  if (threadIdx.x == 0) p.partScale[(size_t)w * p.nPartM + blockIdx.x] = s;
}

__global__ void __launch_bounds__(128) k_update_poses(BaPtrs p) {
  __shared__ double sm[32];
  const int w = blockIdx.x;
  BaState& st = p.st[w];
  if (!st.active) return;
  const BaWin W = p.win[w];
  double sc = 0;
  for (int kf = threadIdx.x; kf < W.nKf; kf += blockDim.x) {
    double* P = p.pose + ((size_t)w * p.capKf + kf) * PSTRIDE;
    double* Pb = p.poseBak + ((size_t)w * p.capKf + kf) * PSTRIDE;
    for (int k = 0; k < 7; k++) Pb[k] = P[k];
    const int pi = p.poseIndex[(size_t)w * p.capKf + kf];
    if (pi >= 0 && st.solveOk) {
      const size_t xo = (size_t)w * (p.capKf * 6 + p.capMp * 3);
      const double* xp = p.x + xo + (size_t)pi * 6;
      const double* bp = p.b + xo + (size_t)pi * 6;
      pose_oplus(P, xp);
      for (int i = 0; i < 6; i++) sc += xp[i] * (st.lambda * xp[i] + bp[i]);
    }
  }
  const double s = block_sum(sc, sm);
  if (threadIdx.x == 0) p.partScale[(size_t)w * p.nPartM + (p.nPartM - 1)] = s;  // last slot is reserved for the poses
}

// LM accept/reject + iteration/round bookkeeping (levenberg.cpp:102-161, sparse_optimizer.cpp:376-412,
// src/Optimizer.cc:863-917)
__global__ void k_control_end(BaPtrs p, int batch) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= batch) return;
  BaState& st = p.st[w];
  st.restore = 0;
  st.doOutlier = 0;
  if (!st.active) return;
  const BaWin W = p.win[w];
  const int nbE = (W.nEdges + 255) / 256, nbM = (W.nMp + 127) / 128;
  double tempChi = 0;
  for (int k = 0; k < nbE; k++) tempChi += p.partChi[(size_t)w * p.nPartE + k];
  double scale = p.partScale[(size_t)w * p.nPartM + (p.nPartM - 1)];
  for (int k = 0; k < nbM; k++) scale += p.partScale[(size_t)w * p.nPartM + k];
  st.chi2Final = tempChi;  // activeRobustChi2() of the last evaluated state (diagnostic)
  if (!st.solveOk) {
    tempChi = DBL_MAX;
    scale = 0;
  }
  scale += 1e-3;
  double rho = (st.currentChi - tempChi) / scale;
  if (!st.solveOk) rho = -1;
  const bool good = rho > 0 && isfinite(tempChi);
  if (st.nTrials < 255) st.trace[st.nTrials] = good ? 1 : 0;
  st.nTrials++;
  if (good) {
    double alpha = 1. - pow((2 * rho - 1), 3);
    alpha = fmin(alpha, 2. / 3.);
    st.lambda *= fmax(1. / 3., alpha);
    st.ni = 2;
    st.currentChi = tempChi;
  } else {
    st.lambda *= st.ni;
    st.ni *= 2;
    st.restore = 1;
  }
  st.rho = rho;
  st.tempChi = tempChi;
  st.qmax++;
  if (rho < 0 && st.qmax < 10 && !st.stop) return;  // another trial with the larger lambda
  bool roundEnd = false;
  if (st.qmax == 10 || rho == 0) {
    roundEnd = true;
  } else {
    if ((st.iniChi - st.currentChi) * 1e3 < st.iniChi) st.nBad++;
    else st.nBad = 0;
    if (st.nBad >= 3) roundEnd = true;
  }
  st.it++;
  if (st.it >= st.its[st.round] || st.stop) roundEnd = true;
  if (!roundEnd) {
    st.needBuild = 1;
    return;
  }
  if (st.round == 0 && !st.stop && st.its[1] > 0) {
    st.doOutlier = 1;  // setLevel(1) on outliers, drop the robust kernels, second round
    st.round = 1;
    st.it = 0;
    st.robust = 0;
    st.needBuild = 1;
  } else {
    // its2 == 0 or stop after round 1: the final outlier test still runs (:921-958)
    st.doOutlier = 2;
    st.active = 0;
  }
}

// pop(): restore the state of a rejected trial; then the chi2/depth outlier tests (src/Optimizer.cc:880-958)
__global__ void __launch_bounds__(256) k_restore_outliers(BaPtrs p) {
  const int w = blockIdx.y;
  const BaState& st = p.st[w];
  const BaWin W = p.win[w];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (st.restore) {
    if (i < W.nMp * 3) p.pts[(size_t)w * p.capMp * 3 + i] = p.ptsBak[(size_t)w * p.capMp * 3 + i];
    if (i < W.nKf * PSTRIDE) p.pose[(size_t)w * p.capKf * PSTRIDE + i] = p.poseBak[(size_t)w * p.capKf * PSTRIDE + i];
  }
}
__global__ void __launch_bounds__(256) k_outliers(BaPtrs p) {
  const int w = blockIdx.y;
  const BaState& st = p.st[w];
  if (!st.doOutlier) return;
  const BaWin W = p.win[w];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.nEdges) return;
  const size_t eo = (size_t)w * p.capE + e;
  double Xc[3];
  pose_map(p.pose + ((size_t)w * p.capKf + p.eKf[eo]) * PSTRIDE, p.pts + ((size_t)w * p.capMp + p.eMp[eo]) * 3, Xc);
  const double th = p.eStereo[eo] ? 7.815 : 5.991;
  const bool out = p.chi2[eo] > th || !(Xc[2] > 0.0);
  if (st.doOutlier == 1) p.eLevel[eo] = out ? 1 : p.eLevel[eo];
  else p.eOutlier[eo] = out ? 1 : 0;
}

__global__ void k_any_active(BaPtrs p, int batch) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < batch && p.st[w].active) atomicOr(p.anyActive, 1);
}

// final robust/plain chi2 of the active edges (diagnostic) comes from st.tempChi/currentChi; nothing else to do.

}  // namespace b2s

using namespace b2s;

struct b2s_ba_solver {
  int maxKf, maxMp, maxE, maxBatch, device;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  BaPtrs d;
  std::vector<void*> allocs;
  // pinned staging
  int* hAny = nullptr;
  size_t cholSmem = 0;
};

static void quat_from_R_host(const double m[3][3], double* q) {
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[2][1] - m[1][2]) * t;
    q[1] = (m[0][2] - m[2][0]) * t;
    q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    v[j] = (m[j][i] + m[i][j]) * t;
    v[k] = (m[k][i] + m[i][k]) * t;
    q[0] = v[0]; q[1] = v[1]; q[2] = v[2];
  }
  if (q[3] < 0) {
    q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3];
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

extern "C" int b2s_ba_create(int max_kf, int max_mp, int max_edges, int max_batch, int device, b2s_ba_solver** out) {
  if (!out || max_kf < 1 || max_kf > 128 || max_mp < 1 || max_edges < 1 || max_batch < 1) {
    set_error("b2s_ba_create: bad argument (max_kf must be in [1,128])");
    return B2S_ERR_BAD_ARG;
  }
  *out = nullptr;
  int rc = select_device(device);
  if (rc != B2S_OK) return rc;
  b2s_ba_solver* h = new b2s_ba_solver();
  h->maxKf = max_kf; h->maxMp = max_mp; h->maxE = max_edges; h->maxBatch = max_batch; h->device = device;
  BaPtrs& d = h->d;
  memset(&d, 0, sizeof(d));
  d.capKf = max_kf; d.capMp = max_mp; d.capE = max_edges;
  d.ldS = max_kf * 6 + 1;
  d.nPartE = div_up(max_edges, 256);
  d.nPartM = div_up(max_mp, 128) + 1;
  const size_t B = max_batch;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  auto A = [&](void* pp, size_t bytes) {
    void** p = (void**)pp;
    if (e == cudaSuccess) {
      e = cudaMalloc(p, bytes);
      if (e == cudaSuccess) {
        h->allocs.push_back(*p);
        e = cudaMemset(*p, 0, bytes);
      }
    }
  };
  A(&d.win, B * sizeof(BaWin)); A(&d.st, B * sizeof(BaState));
  A(&d.pose, B * max_kf * PSTRIDE * 8); A(&d.poseBak, B * max_kf * PSTRIDE * 8);
  A(&d.pts, B * max_mp * 3 * 8); A(&d.ptsBak, B * max_mp * 3 * 8);
  A(&d.poseIndex, B * max_kf * 4);
  A(&d.eKf, B * max_edges * 4); A(&d.eMp, B * max_edges * 4);
  A(&d.eObs, B * max_edges * 12); A(&d.eW, B * max_edges * 4);
  A(&d.eStereo, B * max_edges); A(&d.eLevel, B * max_edges); A(&d.eOutlier, B * max_edges);
  A(&d.err, B * max_edges * 24); A(&d.chi2, B * max_edges * 8); A(&d.W, B * max_edges * 18 * 8);
  A(&d.mpStart, B * (max_mp + 1) * 4); A(&d.mpEdges, B * max_edges * 4);
  A(&d.kfStart, B * (max_kf + 1) * 4); A(&d.kfEdges, B * max_edges * 4);
  A(&d.Hpp, B * max_kf * 36 * 8); A(&d.Hll, B * max_mp * 9 * 8);
  A(&d.b, B * ((size_t)max_kf * 6 + (size_t)max_mp * 3) * 8); A(&d.x, B * ((size_t)max_kf * 6 + (size_t)max_mp * 3) * 8);
  A(&d.Dinv, B * max_mp * 9 * 8); A(&d.S, B * (size_t)d.ldS * d.ldS * 8);
  A(&d.db, B * max_mp * 3 * 8); A(&d.Y, B * max_edges * 18 * 8);
  A(&d.lmEdge, B * (size_t)max_mp * max_kf * 4); A(&d.freeKf, B * max_kf * 4);
  A(&d.partChi, B * d.nPartE * 8); A(&d.partScale, B * d.nPartM * 8);
  A(&d.anyActive, 4);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&h->hAny, 4);
  h->cholSmem = (size_t)(CHOL_BS * CHOL_PP + (size_t)(d.ldS + 4) * CHOL_PP + d.ldS) * 8;
  if (e == cudaSuccess && h->cholSmem > 48 * 1024)
    e = cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->cholSmem);
  if (e != cudaSuccess) {
    set_error("b2s_ba_create: %s", cudaGetErrorString(e));
    b2s_ba_destroy(h);
    return B2S_ERR_CUDA;
  }
  *out = h;
  return B2S_OK;
}

extern "C" void b2s_ba_destroy(b2s_ba_solver* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (void* p : h->allocs) cudaFree(p);
  if (h->hAny) cudaFreeHost(h->hAny);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}
extern "C" long long b2s_ba_launch_count(const b2s_ba_solver* h) { return h ? h->launches : 0; }

static int ba_run(b2s_ba_solver* h, int batch, const b2s_ba_problem* probs, const volatile uint8_t* stop,
                  b2s_ba_result* res) {
  if (!h || !probs || !res || batch < 1 || batch > h->maxBatch) {
    set_error("b2s_local_ba: bad argument");
    return B2S_ERR_BAD_ARG;
  }
  BaPtrs& d = h->d;
  B2S_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  if (stop && *stop) return B2S_ERR_ABORTED;  // src/Optimizer.cc:858-860: return without write-back
  int maxE = 0, maxMp = 0, maxKf = 0, maxN = 0, maxFree = 0;
  std::vector<BaWin> wins(batch);
  std::vector<BaState> states(batch);
  // host staging (pageable -> device); a window is a few MB at most
  for (int w = 0; w < batch; w++) {
    const b2s_ba_problem& P = probs[w];
    if (P.its1 < 1 || P.its2 < 0) {
      set_error("b2s_local_ba: its1 must be >= 1 and its2 >= 0");
      return B2S_ERR_BAD_ARG;
    }
    if (P.n_kf < 1 || P.n_kf > h->maxKf || P.n_mp < 0 || P.n_mp > h->maxMp || P.n_edges < 0 || P.n_edges > h->maxE ||
        P.n_local < 0 || P.n_local > P.n_kf || !P.Tcw || !P.fixed || (P.n_mp && !P.points) || (P.n_edges && !P.edges)) {
      set_error("b2s_local_ba: window %d exceeds the solver's capacity or has null arrays", w);
      return B2S_ERR_BAD_ARG;
    }
    std::vector<double> pose((size_t)P.n_kf * PSTRIDE, 0.0);
    std::vector<int> pidx(P.n_kf, -1), freeKf;
    int nFree = 0;
    for (int k = 0; k < P.n_kf; k++) {
      const float* T = P.Tcw + (size_t)k * 16;
      double R[3][3];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i][j] = T[i * 4 + j];  // Converter::toSE3Quat (src/Converter.cc:57-66)
      quat_from_R_host(R, &pose[(size_t)k * PSTRIDE]);
      pose[(size_t)k * PSTRIDE + 4] = T[3];
      pose[(size_t)k * PSTRIDE + 5] = T[7];
      pose[(size_t)k * PSTRIDE + 6] = T[11];
      if (!P.fixed[k]) {
        pidx[k] = nFree++;
        freeKf.push_back(k);
      }
    }
    std::vector<double> pts((size_t)P.n_mp * 3);
    for (size_t i = 0; i < pts.size(); i++) pts[i] = P.points[i];  // Converter::toVector3d
    std::vector<int> eKf(P.n_edges), eMp(P.n_edges), mpStart(P.n_mp + 1, 0), kfStart(P.n_kf + 1, 0), mpEdges(P.n_edges),
        kfEdges(P.n_edges);
    std::vector<float> eObs((size_t)P.n_edges * 3), eW(P.n_edges);
    std::vector<uint8_t> eSt(P.n_edges);
    for (int e = 0; e < P.n_edges; e++) {
      const b2s_ba_edge& E = P.edges[e];
      if (E.kf < 0 || E.kf >= P.n_kf || E.mp < 0 || E.mp >= P.n_mp) {
        set_error("b2s_local_ba: edge %d references a vertex out of range", e);
        return B2S_ERR_BAD_ARG;
      }
      eKf[e] = E.kf; eMp[e] = E.mp;
      eObs[(size_t)e * 3] = E.obs[0]; eObs[(size_t)e * 3 + 1] = E.obs[1]; eObs[(size_t)e * 3 + 2] = E.obs[2];
      eW[e] = E.inv_sigma2;
      eSt[e] = !(E.obs[2] < 0);  // mvuRight<0 -> monocular edge (src/Optimizer.cc:794)
      mpStart[E.mp + 1]++;
      kfStart[E.kf + 1]++;
    }
    for (int i = 0; i < P.n_mp; i++) mpStart[i + 1] += mpStart[i];
    for (int i = 0; i < P.n_kf; i++) kfStart[i + 1] += kfStart[i];
    {
      std::vector<int> cm(mpStart.begin(), mpStart.end() - 1), ck(kfStart.begin(), kfStart.end() - 1);
      for (int e = 0; e < P.n_edges; e++) {
        mpEdges[cm[eMp[e]]++] = e;
        kfEdges[ck[eKf[e]]++] = e;
      }
    }
    BaWin& Wn = wins[w];
    Wn.nKf = P.n_kf; Wn.nLocal = P.n_local; Wn.nMp = P.n_mp; Wn.nEdges = P.n_edges; Wn.nFree = nFree;
    Wn.fx = P.fx; Wn.fy = P.fy; Wn.cx = P.cx; Wn.cy = P.cy; Wn.bf = P.bf;
    BaState& S0 = states[w];
    memset(&S0, 0, sizeof(S0));
    S0.active = (P.its1 > 0) ? 1 : 0;
    S0.needBuild = 1;
    S0.robust = 1;
    S0.its[0] = P.its1; S0.its[1] = P.its2;
    S0.ni = 2;
    maxE = std::max(maxE, P.n_edges); maxMp = std::max(maxMp, P.n_mp); maxKf = std::max(maxKf, P.n_kf);
    maxN = std::max(maxN, nFree * 6 + 1);
    maxFree = std::max(maxFree, nFree);
#define UP(dst, vec, stride) \
  if (!(vec).empty()) B2S_CUDA(cudaMemcpyAsync((dst) + (size_t)w * (stride), (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice, st))
    UP(d.pose, pose, (size_t)d.capKf * PSTRIDE);
    UP(d.pts, pts, (size_t)d.capMp * 3);
    UP(d.poseIndex, pidx, d.capKf);
    UP(d.freeKf, freeKf, d.capKf);
    UP(d.eKf, eKf, d.capE); UP(d.eMp, eMp, d.capE);
    UP(d.eObs, eObs, (size_t)d.capE * 3); UP(d.eW, eW, d.capE);
    UP(d.eStereo, eSt, d.capE);
    UP(d.mpStart, mpStart, d.capMp + 1); UP(d.mpEdges, mpEdges, d.capE);
    UP(d.kfStart, kfStart, d.capKf + 1); UP(d.kfEdges, kfEdges, d.capE);
#undef UP
    B2S_CUDA(cudaMemsetAsync(d.eLevel + (size_t)w * d.capE, 0, d.capE, st));
    B2S_CUDA(cudaMemsetAsync(d.lmEdge + (size_t)w * d.capMp * d.capKf, 0xFF, (size_t)P.n_mp * d.capKf * 4, st));
    B2S_CUDA(cudaMemsetAsync(d.chi2 + (size_t)w * d.capE, 0, (size_t)d.capE * 8, st));
    B2S_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
  }
  B2S_CUDA(cudaMemcpyAsync(d.win, wins.data(), batch * sizeof(BaWin), cudaMemcpyHostToDevice, st));
  B2S_CUDA(cudaMemcpyAsync(d.st, states.data(), batch * sizeof(BaState), cudaMemcpyHostToDevice, st));
  const int gE = std::max(1, div_up(maxE, 256)), gM = std::max(1, div_up(maxMp, 128));
  const int gRestore = std::max(1, div_up(std::max(maxMp * 3, maxKf * PSTRIDE), 256));
  k_lm_edge<<<dim3(gE, batch), 256, 0, st>>>(d);
  h->launches++;
  bool stopSent = false;
  for (int step = 0; step < 400; step++) {
    if (stop && *stop && !stopSent) {  // asynchronous abort (LocalMapping::InsertKeyFrame sets mbAbortBA)
      for (int w = 0; w < batch; w++) {
        const int one = 1;
        B2S_CUDA(cudaMemcpyAsync((char*)(d.st + w) + offsetof(BaState, stop), &one, 4, cudaMemcpyHostToDevice, st));
      }
      stopSent = true;
    }
    k_errors<<<dim3(gE, batch), 256, 0, st>>>(d, 0);
    k_build_landmarks<<<dim3(gM, batch), 128, 0, st>>>(d);
    k_build_poses<<<dim3(maxKf, batch), 128, 0, st>>>(d);
    k_control_begin<<<div_up(batch, 64), 64, 0, st>>>(d, batch);
    k_dinv<<<dim3(gM, batch), 128, 0, st>>>(d);
    k_schur_pose<<<dim3(maxKf, batch), 128, 0, st>>>(d);
    k_schur_blocks<<<dim3(std::max(1, maxFree * (maxFree + 1) / 2), batch), 64, 0, st>>>(d);
    k_chol<<<batch, CHOL_T, h->cholSmem, st>>>(d);
    k_backsub<<<dim3(gM, batch), 128, 0, st>>>(d);
    k_update_poses<<<batch, 128, 0, st>>>(d);
    k_errors<<<dim3(gE, batch), 256, 0, st>>>(d, 1);
    k_control_end<<<div_up(batch, 64), 64, 0, st>>>(d, batch);
    k_restore_outliers<<<dim3(gRestore, batch), 256, 0, st>>>(d);
    k_outliers<<<dim3(gE, batch), 256, 0, st>>>(d);
    B2S_CUDA(cudaMemsetAsync(d.anyActive, 0, 4, st));
    k_any_active<<<div_up(batch, 64), 64, 0, st>>>(d, batch);
    h->launches += 15;
    B2S_CUDA(cudaMemcpyAsync(h->hAny, d.anyActive, 4, cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    if (!*h->hAny) break;
  }
  B2S_CUDA(cudaGetLastError());
  // write-back values (src/Optimizer.cc:961-996): SetPose(toCvMat(SE3quat)), SetWorldPos(toCvMat(estimate))
  B2S_CUDA(cudaMemcpyAsync(states.data(), d.st, batch * sizeof(BaState), cudaMemcpyDeviceToHost, st));
  for (int w = 0; w < batch; w++) {
    const b2s_ba_problem& P = probs[w];
    b2s_ba_result& R = res[w];
    std::vector<double> pose((size_t)P.n_kf * PSTRIDE), pts((size_t)P.n_mp * 3);
    B2S_CUDA(cudaMemcpyAsync(pose.data(), d.pose + (size_t)w * d.capKf * PSTRIDE, pose.size() * 8, cudaMemcpyDeviceToHost, st));
    if (P.n_mp) B2S_CUDA(cudaMemcpyAsync(pts.data(), d.pts + (size_t)w * d.capMp * 3, pts.size() * 8, cudaMemcpyDeviceToHost, st));
    if (R.edge_outlier && P.n_edges)
      B2S_CUDA(cudaMemcpyAsync(R.edge_outlier, d.eOutlier + (size_t)w * d.capE, P.n_edges, cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    if (R.Tcw_out) {
      for (int k = 0; k < P.n_local; k++) {
        const double* q = &pose[(size_t)k * PSTRIDE];
        const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
        const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
        const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
        const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
        float* T = R.Tcw_out + (size_t)k * 16;
        T[0] = (float)(1 - (tyy + tzz)); T[1] = (float)(txy - twz); T[2] = (float)(txz + twy); T[3] = (float)q[4];
        T[4] = (float)(txy + twz); T[5] = (float)(1 - (txx + tzz)); T[6] = (float)(tyz - twx); T[7] = (float)q[5];
        T[8] = (float)(txz - twy); T[9] = (float)(tyz + twx); T[10] = (float)(1 - (txx + tyy)); T[11] = (float)q[6];
        T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
      }
    }
    if (R.points_out)
      for (size_t i = 0; i < pts.size(); i++) R.points_out[i] = (float)pts[i];
    R.chi2_final = states[w].chi2Final;
    R.n_trials = states[w].nTrials;
    if (R.trace) {
      const int n = std::min(states[w].nTrials, 255);
      for (int i = 0; i < n; i++) R.trace[i] = states[w].trace[i];
      R.trace[n] = -1;
    }
  }
  return B2S_OK;
}

extern "C" int b2s_local_ba(b2s_ba_solver* h, const b2s_ba_problem* p, const volatile uint8_t* stop, b2s_ba_result* r) {
  return ba_run(h, 1, p, stop, r);
}
extern "C" int b2s_local_ba_batch(b2s_ba_solver* h, int batch, const b2s_ba_problem* p, b2s_ba_result* r) {
  return ba_run(h, batch, p, nullptr, r);
}
