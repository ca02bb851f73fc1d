/*
 * oracle/local_ba.cpp — CPU restatement of Optimizer::LocalBundleAdjustment from the point where the
 * graph is built (TEST INFRASTRUCTURE, see orb_oracle.h).
 *
 * Follows /root/reference/src/Optimizer.cc:698-996 and the vendored g2o functions it executes
 * (/root/reference/Thirdparty/g2o/g2o/...):
 *   types/types_six_dof_expmap.{h,cpp}  error, projection (stereo uses a float 1/z), Jacobians
 *   types/se3quat.h, types/se3_ops.hpp  SE3 exp map, quaternion pose composition
 *   core/base_binary_edge.hpp:55-120    constructQuadraticForm (Huber-weighted J^T W J blocks)
 *   core/robust_kernel_impl.cpp:65-91   Huber
 *   core/block_solver.hpp:354-604       lambda on all diagonals, Schur complement, back-substitution
 *   core/optimization_algorithm_levenberg.cpp:61-189  LM control
 *   core/sparse_optimizer.cpp:354-435   optimize loop, update
 * Eigen (absent here) is replaced by hand-written fixed-size FP64 code; the reduced camera system is
 * solved by dense Cholesky instead of SimplicialLDLT (same solution up to rounding << 1e-5).
 * Deviation: a failed factorisation is treated as a rejected LM step (the reference applies a stale
 * solution vector and relies on tempChi = DBL_MAX); never triggered with lambda > 0 on SPD systems.
 */
#include "orb_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct Quat {
  double x, y, z, w;
};
struct Pose {
  Quat r;
  double t[3];
};

// Eigen::Quaterniond(const Matrix3d&) — Eigen/src/Geometry/Quaternion.h (quaternionbase_assign_impl<..,3,3>)
Quat quat_from_R(const double m[3][3]) {
  Quat q;
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[2][1] - m[1][2]) * t;
    q.y = (m[0][2] - m[2][0]) * t;
    q.z = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m[k][j] - m[j][k]) * t;
    v[j] = (m[j][i] + m[i][j]) * t;
    v[k] = (m[k][i] + m[i][k]) * t;
    q.x = v[0];
    q.y = v[1];
    q.z = v[2];
  }
  return q;
}

// SE3Quat::normalizeRotation — se3quat.h:266-271
void normalize_rot(Quat& q) {
  if (q.w < 0) {
    q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w;
  }
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

// Eigen quaternion * vector: v + w*uv + vec x uv, uv = 2*(vec x v)
void quat_rotate(const Quat& q, const double v[3], double out[3]) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
  out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
  out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}

Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

void quat_to_R(const Quat& q, double R[3][3]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
}

// SE3Quat::map — se3quat.h:217-220
inline void pose_map(const Pose& T, const double X[3], double out[3]) {
  quat_rotate(T.r, X, out);
  out[0] += T.t[0]; out[1] += T.t[1]; out[2] += T.t[2];
}

// Converter::toSE3Quat — src/Converter.cc:57-66 (float R,t -> double; SE3Quat(R,t))
Pose pose_from_Tcw(const float* T) {
  double R[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = T[i * 4 + j];
  Pose p;
  p.r = quat_from_R(R);
  normalize_rot(p.r);
  p.t[0] = T[3]; p.t[1] = T[7]; p.t[2] = T[11];
  return p;
}

// Converter::toCvMat(SE3Quat) — src/Converter.cc:73-80,96-107 (to_homogeneous_matrix -> float)
void pose_to_Tcw(const Pose& p, float* T) {
  double R[3][3];
  quat_to_R(p.r, R);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[i * 4 + j] = (float)R[i][j];
    T[i * 4 + 3] = (float)p.t[i];
  }
  T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}

// SE3Quat::exp(update) * estimate — se3quat.h:223-257, :103-109; VertexSE3Expmap::oplusImpl types_six_dof_expmap.h:73-76
void pose_oplus(Pose& T, const double* upd) {
  const double omega[3] = {upd[0], upd[1], upd[2]};
  const double ups[3] = {upd[3], upd[4], upd[5]};
  const double theta = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  double O[3][3] = {{0, -omega[2], omega[1]}, {omega[2], 0, -omega[0]}, {-omega[1], omega[0], 0}};
  double O2[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += O[i][k] * O[k][j];
      O2[i][j] = s;
    }
  double R[3][3], V[3][3];
  if (theta < 0.00001) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        R[i][j] = (i == j ? 1.0 : 0.0) + O[i][j] + O2[i][j];
        V[i][j] = R[i][j];
      }
  } else {
    const double a = std::sin(theta) / theta;
    const double b = (1 - std::cos(theta)) / (theta * theta);
    const double c = (theta - std::sin(theta)) / (std::pow(theta, 3));
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        R[i][j] = (i == j ? 1.0 : 0.0) + a * O[i][j] + b * O2[i][j];
        V[i][j] = (i == j ? 1.0 : 0.0) + b * O[i][j] + c * O2[i][j];
      }
  }
  Pose E;
  E.r = quat_from_R(R);
  normalize_rot(E.r);  // SE3Quat(const Quaterniond&, const Vector3d&) normalises
  for (int i = 0; i < 3; i++) E.t[i] = V[i][0] * ups[0] + V[i][1] * ups[1] + V[i][2] * ups[2];
  // operator*: result._t += _r*tr2._t ; result._r *= tr2._r ; normalizeRotation
  double rt[3];
  quat_rotate(E.r, T.t, rt);
  Pose Rz;
  Rz.t[0] = E.t[0] + rt[0]; Rz.t[1] = E.t[1] + rt[1]; Rz.t[2] = E.t[2] + rt[2];
  Rz.r = quat_mul(E.r, T.r);
  normalize_rot(Rz.r);
  T = Rz;
}

bool inv3(const double* D, double* Di) {  // Eigen 3x3 inverse by cofactors
  const double c00 = D[4] * D[8] - D[5] * D[7];
  const double c01 = D[5] * D[6] - D[3] * D[8];
  const double c02 = D[3] * D[7] - D[4] * D[6];
  const double det = D[0] * c00 + D[1] * c01 + D[2] * c02;
  const double id = 1.0 / det;
  Di[0] = c00 * id; Di[1] = (D[2] * D[7] - D[1] * D[8]) * id; Di[2] = (D[1] * D[5] - D[2] * D[4]) * id;
  Di[3] = c01 * id; Di[4] = (D[0] * D[8] - D[2] * D[6]) * id; Di[5] = (D[2] * D[3] - D[0] * D[5]) * id;
  Di[6] = c02 * id; Di[7] = (D[1] * D[6] - D[0] * D[7]) * id; Di[8] = (D[0] * D[4] - D[1] * D[3]) * id;
  return std::isfinite(id);
}

// dense Cholesky A = L L^T (lower), in place; solves A x = b. n x n row-major.
bool chol_solve(std::vector<double>& A, int n, const double* b, double* x) {
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * y[k];
    y[i] = s / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < n; k++) s -= A[(size_t)k * n + i] * x[k];
    x[i] = s / A[(size_t)i * n + i];
  }
  return true;
}

struct BA {
  const orc_ba_problem* P;
  const volatile uint8_t* stop;
  std::vector<Pose> poses;
  std::vector<double> pts;        // n_mp x 3
  std::vector<int> pose_index;    // free-pose index or -1
  int n_free = 0;
  std::vector<uint8_t> level;     // per edge
  std::vector<double> err;        // per edge x3
  std::vector<double> chi2;       // per edge (last computed while active)
  bool robust = true;
  double delta_mono, delta_stereo;
  // system
  std::vector<double> Hpp, Hll, Hpl, b, x;  // Hpp: n_free x36 (block-diag), Hll: n_mp x 9, Hpl: per edge 18 (6x3)
  double lambda = -1, ni = 2;
  int nBad = 0;
  std::vector<int32_t> trace;
  int n_trials = 0;

  bool terminate() const { return stop && *stop; }
  bool is_stereo(int e) const { return !(P->edges[e].obs[2] < 0); }  // src/Optimizer.cc:794 mvuRight<0 => mono

  // computeError — types_six_dof_expmap.h:90-95 (mono), :122-127 (stereo); cam_project .cpp:141-157
  void compute_errors() {
    for (int e = 0; e < P->n_edges; e++) {
      if (level[e]) continue;
      const orc_ba_edge& ed = P->edges[e];
      double Xc[3];
      pose_map(poses[ed.kf], &pts[(size_t)ed.mp * 3], Xc);
      double* er = &err[(size_t)e * 3];
      const double w = (double)ed.inv_sigma2;
      if (is_stereo(e)) {
        const float invz = (float)(1.0f / Xc[2]);
        const double u = Xc[0] * invz * (double)P->fx + (double)P->cx;
        const double v = Xc[1] * invz * (double)P->fy + (double)P->cy;
        // res[2] = res[0] - bf*invz with `const float& bf` and float invz: float product, double subtraction
        const double ur_f = u - (double)(P->bf * invz);
        er[0] = (double)ed.obs[0] - u;
        er[1] = (double)ed.obs[1] - v;
        er[2] = (double)ed.obs[2] - ur_f;
        chi2[e] = er[0] * (w * er[0]) + er[1] * (w * er[1]) + er[2] * (w * er[2]);
      } else {
        const double u = Xc[0] / Xc[2] * (double)P->fx + (double)P->cx;
        const double v = Xc[1] / Xc[2] * (double)P->fy + (double)P->cy;
        er[0] = (double)ed.obs[0] - u;
        er[1] = (double)ed.obs[1] - v;
        er[2] = 0;
        chi2[e] = er[0] * (w * er[0]) + er[1] * (w * er[1]);
      }
    }
  }

  // RobustKernelHuber::robustify — robust_kernel_impl.cpp:75-91
  inline void huber(double e, double delta, double& rho0, double& rho1) const {
    const double dsqr = delta * delta;
    if (e <= dsqr) {
      rho0 = e;
      rho1 = 1.;
    } else {
      const double sqrte = std::sqrt(e);
      rho0 = 2 * sqrte * delta - dsqr;
      rho1 = delta / sqrte;
    }
  }

  // activeRobustChi2 — sparse_optimizer.cpp:100-113
  double active_robust_chi2() const {
    double chi = 0;
    for (int e = 0; e < P->n_edges; e++) {
      if (level[e]) continue;
      if (robust) {
        double r0, r1;
        huber(chi2[e], is_stereo(e) ? delta_stereo : delta_mono, r0, r1);
        chi += r0;
      } else
        chi += chi2[e];
    }
    return chi;
  }

  // buildSystem — block_solver.hpp:502-560: linearizeOplus (.cpp:103-139, 188-234) + constructQuadraticForm
  void build_system() {
    std::fill(Hpp.begin(), Hpp.end(), 0.0);
    std::fill(Hll.begin(), Hll.end(), 0.0);
    std::fill(Hpl.begin(), Hpl.end(), 0.0);
    std::fill(b.begin(), b.end(), 0.0);
    const double fx = P->fx, fy = P->fy, bf = P->bf;
    for (int e = 0; e < P->n_edges; e++) {
      if (level[e]) continue;
      const orc_ba_edge& ed = P->edges[e];
      const Pose& T = poses[ed.kf];
      double Xc[3], R[3][3];
      pose_map(T, &pts[(size_t)ed.mp * 3], Xc);
      quat_to_R(T.r, R);
      const double x = Xc[0], y = Xc[1], z = Xc[2], z_2 = z * z;
      const bool st = is_stereo(e);
      const int D = st ? 3 : 2;
      double A[3][3], B[3][6];  // A = d e / d point, B = d e / d pose
      if (st) {
        for (int c = 0; c < 3; c++) {
          A[0][c] = -fx * R[0][c] / z + fx * x * R[2][c] / z_2;
          A[1][c] = -fy * R[1][c] / z + fy * y * R[2][c] / z_2;
          A[2][c] = A[0][c] - bf * R[2][c] / z_2;
        }
      } else {
        // _jacobianOplusXi = -1./z * tmp * R
        const double tmp[2][3] = {{fx, 0, -x / z * fx}, {0, fy, -y / z * fy}};
        for (int r = 0; r < 2; r++)
          for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += (-1. / z * tmp[r][k]) * R[k][c];
            A[r][c] = s;
          }
        A[2][0] = A[2][1] = A[2][2] = 0;
      }
      B[0][0] = x * y / z_2 * fx;
      B[0][1] = -(1 + (x * x / z_2)) * fx;
      B[0][2] = y / z * fx;
      B[0][3] = -1. / z * fx;
      B[0][4] = 0;
      B[0][5] = x / z_2 * fx;
      B[1][0] = (1 + y * y / z_2) * fy;
      B[1][1] = -x * y / z_2 * fy;
      B[1][2] = -x / z * fy;
      B[1][3] = 0;
      B[1][4] = -1. / z * fy;
      B[1][5] = y / z_2 * fy;
      if (st) {
        B[2][0] = B[0][0] - bf * y / z_2;
        B[2][1] = B[0][1] + bf * x / z_2;
        B[2][2] = B[0][2];
        B[2][3] = B[0][3];
        B[2][4] = 0;
        B[2][5] = B[0][5] - bf / z_2;
      } else {
        for (int c = 0; c < 6; c++) B[2][c] = 0;
      }
      const double w0 = (double)ed.inv_sigma2;
      double rho1 = 1.0;
      if (robust) {
        double r0;
        huber(chi2[e], st ? delta_stereo : delta_mono, r0, rho1);
      }
      const double* er = &err[(size_t)e * 3];
      double omega_r[3];
      for (int r = 0; r < 3; r++) omega_r[r] = -(w0 * er[r]) * rho1;
      const double w = rho1 * w0;
      // point (always free)
      double* bl = &b[(size_t)n_free * 6 + (size_t)ed.mp * 3];
      double* Hl = &Hll[(size_t)ed.mp * 9];
      for (int i = 0; i < 3; i++) {
        double s = 0;
        for (int r = 0; r < D; r++) s += A[r][i] * omega_r[r];
        bl[i] += s;
        for (int j = 0; j < 3; j++) {
          double h = 0;
          for (int r = 0; r < D; r++) h += A[r][i] * w * A[r][j];
          Hl[i * 3 + j] += h;
        }
      }
      const int pi = pose_index[ed.kf];
      if (pi >= 0) {
        double* bp = &b[(size_t)pi * 6];
        double* Hp = &Hpp[(size_t)pi * 36];
        double* W = &Hpl[(size_t)e * 18];
        for (int i = 0; i < 6; i++) {
          double s = 0;
          for (int r = 0; r < D; r++) s += B[r][i] * omega_r[r];
          bp[i] += s;
          for (int j = 0; j < 6; j++) {
            double h = 0;
            for (int r = 0; r < D; r++) h += B[r][i] * w * B[r][j];
            Hp[i * 6 + j] += h;
          }
          for (int j = 0; j < 3; j++) {
            double h = 0;
            for (int r = 0; r < D; r++) h += B[r][i] * w * A[r][j];
            W[i * 3 + j] += h;
          }
        }
      }
    }
  }

  // computeLambdaInit — levenberg.cpp:166-180
  double lambda_init() const {
    double mx = 0;
    for (int i = 0; i < n_free; i++)
      for (int j = 0; j < 6; j++) mx = std::max(std::fabs(Hpp[(size_t)i * 36 + j * 7]), mx);
    for (int l = 0; l < P->n_mp; l++)
      for (int j = 0; j < 3; j++) mx = std::max(std::fabs(Hll[(size_t)l * 9 + j * 4]), mx);
    return 1e-5 * mx;
  }

  // setLambda + BlockSolver::solve (Schur) — block_solver.hpp:564-604, 354-486
  std::vector<std::vector<int>> mp_edges;  // per landmark: active edges with a free pose
  bool solve(double lam) {
    const int np = n_free * 6;
    std::vector<double> S((size_t)np * np, 0.0), bs(np);
    for (int i = 0; i < n_free; i++)
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++)
          S[(size_t)(i * 6 + r) * np + (i * 6 + c)] = Hpp[(size_t)i * 36 + r * 6 + c] + (r == c ? lam : 0.0);
    for (int i = 0; i < np; i++) bs[i] = b[i];
    std::vector<double> Dinv((size_t)P->n_mp * 9);
    for (int l = 0; l < P->n_mp; l++) {
      double D[9];
      for (int k = 0; k < 9; k++) D[k] = Hll[(size_t)l * 9 + k];
      D[0] += lam; D[4] += lam; D[8] += lam;
      double* Di = &Dinv[(size_t)l * 9];
      inv3(D, Di);
      const double* bl = &b[(size_t)np + (size_t)l * 3];
      double db[3];
      for (int i = 0; i < 3; i++) db[i] = Di[i * 3] * bl[0] + Di[i * 3 + 1] * bl[1] + Di[i * 3 + 2] * bl[2];
      const std::vector<int>& el = mp_edges[l];
      for (size_t a = 0; a < el.size(); a++) {
        const double* Bi = &Hpl[(size_t)el[a] * 18];
        const int i1 = pose_index[P->edges[el[a]].kf];
        double BD[18];
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 3; c++)
            BD[r * 3 + c] = Bi[r * 3] * Di[c] + Bi[r * 3 + 1] * Di[3 + c] + Bi[r * 3 + 2] * Di[6 + c];
        for (int r = 0; r < 6; r++) bs[i1 * 6 + r] -= Bi[r * 3] * db[0] + Bi[r * 3 + 1] * db[1] + Bi[r * 3 + 2] * db[2];
        for (size_t c2 = 0; c2 < el.size(); c2++) {
          const double* Bj = &Hpl[(size_t)el[c2] * 18];
          const int i2 = pose_index[P->edges[el[c2]].kf];
          for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++)
              S[(size_t)(i1 * 6 + r) * np + (i2 * 6 + c)] -=
                  BD[r * 3] * Bj[c * 3] + BD[r * 3 + 1] * Bj[c * 3 + 1] + BD[r * 3 + 2] * Bj[c * 3 + 2];
        }
      }
    }
    if (np > 0) {
      if (!chol_solve(S, np, bs.data(), x.data())) return false;
    }
    // xl = Dinv * (bl - Bt * xp)
    for (int l = 0; l < P->n_mp; l++) {
      double cl[3] = {b[(size_t)np + l * 3], b[(size_t)np + l * 3 + 1], b[(size_t)np + l * 3 + 2]};
      const std::vector<int>& el = mp_edges[l];
      for (size_t a = 0; a < el.size(); a++) {
        const double* Bi = &Hpl[(size_t)el[a] * 18];
        const double* xp = &x[(size_t)pose_index[P->edges[el[a]].kf] * 6];
        for (int c = 0; c < 3; c++)
          for (int r = 0; r < 6; r++) cl[c] -= Bi[r * 3 + c] * xp[r];
      }
      const double* Di = &Dinv[(size_t)l * 9];
      for (int i = 0; i < 3; i++) x[(size_t)np + l * 3 + i] = Di[i * 3] * cl[0] + Di[i * 3 + 1] * cl[1] + Di[i * 3 + 2] * cl[2];
    }
    return true;
  }

  void rebuild_structure() {
    mp_edges.assign(P->n_mp, std::vector<int>());
    for (int e = 0; e < P->n_edges; e++)
      if (!level[e] && pose_index[P->edges[e].kf] >= 0) mp_edges[P->edges[e].mp].push_back(e);
  }

  // SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve
  void optimize(int iterations) {
    // initializeOptimization(level) activates only the edges of that level and the vertices they touch
    // (sparse_optimizer.cpp:292-358); with no active edge the index mapping is empty and optimize() returns at once
    // ("0 vertices to optimize", sparse_optimizer.cpp:356-359).  Every edge has a free landmark, so "no active free vertex"
    // is "no level-0 edge".  Pinned by tests/test_oracle_reference_optimizer.py::rejections_2 (round 1 flags everything).
    bool any_active = false;
    for (uint8_t l : level) any_active |= (l == 0);
    if (!any_active) return;
    rebuild_structure();
    const int nx = n_free * 6 + P->n_mp * 3;
    bool ok = true;
    for (int it = 0; it < iterations && !terminate() && ok; it++) {
      compute_errors();
      double currentChi = active_robust_chi2();
      double tempChi = currentChi;
      const double iniChi = currentChi;
      build_system();
      if (it == 0) {
        lambda = lambda_init();
        ni = 2;
        nBad = 0;
      }
      double rho = 0;
      int qmax = 0;
      do {
        std::vector<Pose> backupP = poses;  // push()
        std::vector<double> backupX = pts;
        const bool ok2 = solve(lambda);
        if (ok2) {
          // SparseOptimizer::update — sparse_optimizer.cpp:422-435
          for (int k = 0; k < P->n_kf; k++)
            if (pose_index[k] >= 0) pose_oplus(poses[k], &x[(size_t)pose_index[k] * 6]);
          for (int l = 0; l < P->n_mp; l++)
            for (int c = 0; c < 3; c++) pts[(size_t)l * 3 + c] += x[(size_t)n_free * 6 + l * 3 + c];  // types_sba.h:52-56
        }
        compute_errors();
        tempChi = active_robust_chi2();
        if (!ok2) tempChi = DBL_MAX;
        rho = (currentChi - tempChi);
        double scale = 0;
        if (ok2)
          for (int j = 0; j < nx; j++) scale += x[j] * (lambda * x[j] + b[j]);  // computeScale :182-189
        scale += 1e-3;
        rho /= scale;
        if (!ok2) rho = -1;
        const bool good = rho > 0 && std::isfinite(tempChi);
        if (n_trials < 250) trace.push_back(good ? 1 : 0);
        n_trials++;
        if (good) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, 2. / 3.);
          const double scaleFactor = std::max(1. / 3., alpha);
          lambda *= scaleFactor;
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni;
          ni *= 2;
          poses = backupP;  // pop()
          pts = backupX;
        }
        qmax++;
      } while (rho < 0 && qmax < 10 && !terminate());
      if (qmax == 10 || rho == 0) {
        ok = false;
        continue;
      }
      if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
      else nBad = 0;
      if (nBad >= 3) ok = false;
    }
  }

  bool depth_positive(int e) const {
    double Xc[3];
    pose_map(poses[P->edges[e].kf], &pts[(size_t)P->edges[e].mp * 3], Xc);
    return Xc[2] > 0.0;
  }

  void init() {
    const int ne = P->n_edges;
    poses.resize(P->n_kf);
    pose_index.assign(P->n_kf, -1);
    n_free = 0;
    for (int k = 0; k < P->n_kf; k++) {
      poses[k] = pose_from_Tcw(P->Tcw + (size_t)k * 16);
      if (!P->fixed[k]) pose_index[k] = n_free++;
    }
    pts.resize((size_t)P->n_mp * 3);
    for (size_t i = 0; i < pts.size(); i++) pts[i] = P->points[i];  // Converter::toVector3d float->double
    level.assign(ne, 0);
    err.assign((size_t)ne * 3, 0.0);
    chi2.assign(ne, 0.0);
    delta_mono = (double)(float)std::sqrt(5.991);    // const float thHuberMono = sqrt(5.991)  (src/Optimizer.cc:764)
    delta_stereo = (double)(float)std::sqrt(7.815);  // :765
    Hpp.assign((size_t)n_free * 36, 0.0);
    Hll.assign((size_t)P->n_mp * 9, 0.0);
    Hpl.assign((size_t)ne * 18, 0.0);
    b.assign((size_t)n_free * 6 + (size_t)P->n_mp * 3, 0.0);
    x.assign(b.size(), 0.0);
  }

  int run(orc_ba_result* r) {
    const int ne = P->n_edges;
    init();
    if (terminate()) return 1;  // src/Optimizer.cc:858-860
    robust = true;
    optimize(P->its1);
    bool doMore = !terminate();
    if (doMore) {
      for (int e = 0; e < ne; e++) {
        const double th = is_stereo(e) ? 7.815 : 5.991;
        if (chi2[e] > th || !depth_positive(e)) level[e] = 1;
      }
      robust = false;
      optimize(P->its2);
    }
    for (int e = 0; e < ne; e++) {
      const double th = is_stereo(e) ? 7.815 : 5.991;
      r->edge_outlier[e] = (chi2[e] > th || !depth_positive(e)) ? 1 : 0;
    }
    for (int k = 0; k < P->n_local; k++) pose_to_Tcw(poses[k], r->Tcw_out + (size_t)k * 16);
    for (size_t i = 0; i < pts.size(); i++) r->points_out[i] = (float)pts[i];
    // final robust/plain chi2 over active edges at the last evaluated state
    r->chi2_final = active_robust_chi2();
    r->n_trials = n_trials;
    if (r->trace) {
      size_t n = std::min<size_t>(trace.size(), 255);
      for (size_t i = 0; i < n; i++) r->trace[i] = trace[i];
      r->trace[n] = -1;
    }
    return 0;
  }
};

// ------------------------------------------------------------------------------------------------
// Optimizer::PoseOptimization (src/Optimizer.cc:363-605): one free SE3 vertex, unary edges
// EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose (types_six_dof_expmap.h:143-205, .cpp:266-364) with fixed
// map points, BlockSolver_6_3 + LinearSolverDense (6x6 LDLT), Levenberg; 4 rounds of 10 iterations from the SAME
// initial pose, inlier/outlier re-classification after every round, Huber kernels dropped after round 2.
// ------------------------------------------------------------------------------------------------
struct PO {
  const orc_pose_problem* P;
  Pose pose;
  std::vector<int> edges;      // feature index per edge (insertion order = ascending feature index, :414-487)
  std::vector<uint8_t> level;  // per edge
  std::vector<double> err, chi2;
  bool robust = true;
  double delta_mono, delta_stereo;
  double H[36], b[6], x[6];
  double lambda = -1, ni = 2;
  int nBadLM = 0;
  std::vector<int32_t> trace;
  int n_trials = 0;

  bool is_stereo(int f) const { return !(P->uright[f] < 0); }  // :418 mvuRight<0 => monocular edge

  void compute_error(int e) {  // computeError, types_six_dof_expmap.h:153-157 / :184-188; cam_project .cpp:295-312
    const int f = edges[e];
    const double Xw[3] = {(double)P->Xw[f * 3], (double)P->Xw[f * 3 + 1], (double)P->Xw[f * 3 + 2]};
    double Xc[3];
    pose_map(pose, Xw, Xc);
    double* er = &err[(size_t)e * 3];
    const double w = (double)P->inv_sigma2[f];
    if (is_stereo(f)) {
      const float invz = (float)(1.0f / Xc[2]);
      const double u = Xc[0] * invz * (double)P->fx + (double)P->cx;
      const double v = Xc[1] * invz * (double)P->fy + (double)P->cy;
      const double ur = u - (double)P->bf * invz;  // `double bf` member times float invz (.cpp:310)
      er[0] = (double)P->kpx[f] - u;
      er[1] = (double)P->kpy[f] - v;
      er[2] = (double)P->uright[f] - ur;
      chi2[e] = er[0] * (w * er[0]) + er[1] * (w * er[1]) + er[2] * (w * er[2]);
    } else {
      const double u = Xc[0] / Xc[2] * (double)P->fx + (double)P->cx;
      const double v = Xc[1] / Xc[2] * (double)P->fy + (double)P->cy;
      er[0] = (double)P->kpx[f] - u;
      er[1] = (double)P->kpy[f] - v;
      er[2] = 0;
      chi2[e] = er[0] * (w * er[0]) + er[1] * (w * er[1]);
    }
  }
  void compute_active_errors() {
    for (size_t e = 0; e < edges.size(); e++)
      if (!level[e]) compute_error((int)e);
  }
  static void huber(double e, double delta, double& rho0, double& rho1) {
    const double dsqr = delta * delta;
    if (e <= dsqr) {
      rho0 = e;
      rho1 = 1.;
    } else {
      const double sqrte = std::sqrt(e);
      rho0 = 2 * sqrte * delta - dsqr;
      rho1 = delta / sqrte;
    }
  }
  double active_robust_chi2() const {
    double chi = 0;
    for (size_t e = 0; e < edges.size(); e++) {
      if (level[e]) continue;
      if (robust) {
        double r0, r1;
        huber(chi2[e], is_stereo(edges[e]) ? delta_stereo : delta_mono, r0, r1);
        chi += r0;
      } else
        chi += chi2[e];
    }
    return chi;
  }
  void build_system() {  // linearizeOplus (.cpp:266-290, :335-364) + BaseUnaryEdge::constructQuadraticForm
    for (double& v : H) v = 0;
    for (double& v : b) v = 0;
    const double fx = P->fx, fy = P->fy, bf = P->bf;
    for (size_t e = 0; e < edges.size(); e++) {
      if (level[e]) continue;
      const int f = edges[e];
      const double Xw[3] = {(double)P->Xw[f * 3], (double)P->Xw[f * 3 + 1], (double)P->Xw[f * 3 + 2]};
      double Xc[3];
      pose_map(pose, Xw, Xc);
      const double x = Xc[0], y = Xc[1];
      const double invz = 1.0 / Xc[2], invz_2 = invz * invz;
      const bool st = is_stereo(f);
      const int D = st ? 3 : 2;
      double B[3][6];
      B[0][0] = x * y * invz_2 * fx;
      B[0][1] = -(1 + (x * x * invz_2)) * fx;
      B[0][2] = y * invz * fx;
      B[0][3] = -invz * fx;
      B[0][4] = 0;
      B[0][5] = x * invz_2 * fx;
      B[1][0] = (1 + y * y * invz_2) * fy;
      B[1][1] = -x * y * invz_2 * fy;
      B[1][2] = -x * invz * fy;
      B[1][3] = 0;
      B[1][4] = -invz * fy;
      B[1][5] = y * invz_2 * fy;
      if (st) {
        B[2][0] = B[0][0] - bf * y * invz_2;
        B[2][1] = B[0][1] + bf * x * invz_2;
        B[2][2] = B[0][2];
        B[2][3] = B[0][3];
        B[2][4] = 0;
        B[2][5] = B[0][5] - bf * invz_2;
      } else {
        for (int c = 0; c < 6; c++) B[2][c] = 0;
      }
      const double w0 = (double)P->inv_sigma2[f];
      double rho1 = 1.0;
      if (robust) {
        double r0;
        huber(chi2[e], st ? delta_stereo : delta_mono, r0, rho1);
      }
      const double* er = &err[e * 3];
      double omega_r[3];
      for (int r = 0; r < 3; r++) omega_r[r] = -(w0 * er[r]) * rho1;
      const double w = rho1 * w0;
      for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int r = 0; r < D; r++) s += B[r][i] * omega_r[r];
        b[i] += s;
        for (int j = 0; j < 6; j++) {
          double h = 0;
          for (int r = 0; r < D; r++) h += B[r][i] * w * B[r][j];
          H[i * 6 + j] += h;
        }
      }
    }
  }
  bool solve(double lam) {
    std::vector<double> S(36);
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 6; c++) S[r * 6 + c] = H[r * 6 + c] + (r == c ? lam : 0.0);
    return chol_solve(S, 6, b, x);
  }
  int n_active() const {
    int n = 0;
    for (uint8_t l : level) n += !l;
    return n;
  }
  void optimize(int iterations) {  // SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve (as in BA::optimize)
    if (n_active() == 0) return;
    bool ok = true;
    for (int it = 0; it < iterations && ok; it++) {
      compute_active_errors();
      double currentChi = active_robust_chi2();
      double tempChi = currentChi;
      const double iniChi = currentChi;
      build_system();
      if (it == 0) {
        double mx = 0;
        for (int j = 0; j < 6; j++) mx = std::max(std::fabs(H[j * 7]), mx);
        lambda = 1e-5 * mx;
        ni = 2;
        nBadLM = 0;
      }
      double rho = 0;
      int qmax = 0;
      do {
        const Pose backup = pose;
        const bool ok2 = solve(lambda);
        if (ok2) pose_oplus(pose, x);
        compute_active_errors();
        tempChi = active_robust_chi2();
        if (!ok2) tempChi = DBL_MAX;
        rho = (currentChi - tempChi);
        double scale = 0;
        if (ok2)
          for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
        scale += 1e-3;
        rho /= scale;
        if (!ok2) rho = -1;
        const bool good = rho > 0 && std::isfinite(tempChi);
        if (n_trials < 250) trace.push_back(good ? 1 : 0);
        n_trials++;
        if (good) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, 2. / 3.);
          lambda *= std::max(1. / 3., alpha);
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni;
          ni *= 2;
          pose = backup;
        }
        qmax++;
      } while (rho < 0 && qmax < 10);
      if (qmax == 10 || rho == 0) {
        ok = false;
        continue;
      }
      if ((iniChi - currentChi) * 1e3 < iniChi) nBadLM++;
      else nBadLM = 0;
      if (nBadLM >= 3) ok = false;
    }
  }
  int run(orc_pose_result* r) {
    const int N = P->n;
    for (int i = 0; i < N; i++) {
      r->outlier[i] = 0;  // pFrame->mvbOutlier[i] = false for every feature with a MapPoint (:421,457)
      if (P->has_mp[i]) edges.push_back(i);
    }
    const int nInitialCorrespondences = (int)edges.size();
    const Pose initial = pose_from_Tcw(P->Tcw);
    pose = initial;
    r->n_trials = 0;
    if (r->trace) r->trace[0] = -1;
    if (nInitialCorrespondences < 3) {  // :492-493 (the frame pose is left untouched)
      for (int k = 0; k < 16; k++) r->Tcw_out[k] = P->Tcw[k];
      return 0;
    }
    level.assign(edges.size(), 0);
    err.assign(edges.size() * 3, 0.0);
    chi2.assign(edges.size(), 0.0);
    delta_mono = (double)(float)std::sqrt(5.991);    // const float deltaMono = sqrt(5.991) (:406)
    delta_stereo = (double)(float)std::sqrt(7.815);  // :407
    const float chi2Mono = 5.991f, chi2Stereo = 7.815f;  // :496-497 (float arrays)
    int nBad = 0;
    robust = true;
    for (int it = 0; it < 4; it++) {
      pose = initial;  // vSE3->setEstimate(Converter::toSE3Quat(pFrame->mTcw)) (:505)
      optimize(10);
      nBad = 0;
      for (size_t e = 0; e < edges.size(); e++) {  // :510-567 (mono and stereo lists; per-edge logic is identical)
        const int f = edges[e];
        if (r->outlier[f]) compute_error((int)e);  // inactive edges are not refreshed by the optimizer
        const float c = (float)chi2[e];
        if (c > (is_stereo(f) ? chi2Stereo : chi2Mono)) {
          r->outlier[f] = 1;
          level[e] = 1;
          nBad++;
        } else {
          r->outlier[f] = 0;
          level[e] = 0;
        }
      }
      if (it == 2) robust = false;  // e->setRobustKernel(0)
      if (edges.size() < 10) break;  // optimizer.edges().size()<10 (:569-570)
    }
    pose_to_Tcw(pose, r->Tcw_out);
    r->n_trials = n_trials;
    if (r->trace) {
      size_t n = std::min<size_t>(trace.size(), 255);
      for (size_t i = 0; i < n; i++) r->trace[i] = trace[i];
      r->trace[n] = -1;
    }
    return nInitialCorrespondences - nBad;
  }
};
}  // namespace

extern "C" int orc_pose_optimization(const orc_pose_problem* p, orc_pose_result* r) {
  PO po;
  po.P = p;
  return po.run(r);
}

// ---- test hooks (tests/test_oracle_ba_math.py): the linear system of the first LM iteration and one damped solve, and
// the pose retraction, exposed so that they can be checked against first principles (numeric differentiation of the
// projection, a dense solve of the full normal equations, the matrix exponential of the twist).
extern "C" int orc_ba_debug_linear_system(const orc_ba_problem* p, int robust, double lambda, double* Hpp, double* Hll,
                                          double* Hpl, double* b, double* err, double* chi2, int32_t* pose_index, double* x) {
  BA ba;
  ba.P = p;
  ba.stop = nullptr;
  ba.init();
  ba.robust = robust != 0;
  ba.rebuild_structure();
  ba.compute_errors();
  ba.build_system();
  std::copy(ba.Hpp.begin(), ba.Hpp.end(), Hpp);
  std::copy(ba.Hll.begin(), ba.Hll.end(), Hll);
  std::copy(ba.Hpl.begin(), ba.Hpl.end(), Hpl);
  std::copy(ba.b.begin(), ba.b.end(), b);
  std::copy(ba.err.begin(), ba.err.end(), err);
  std::copy(ba.chi2.begin(), ba.chi2.end(), chi2);
  std::copy(ba.pose_index.begin(), ba.pose_index.end(), pose_index);
  if (x) {
    if (!ba.solve(lambda)) return -1;
    std::copy(ba.x.begin(), ba.x.end(), x);
  }
  return ba.n_free;
}

// PoseOptimization: H (6x6), b, the damped solve x, per-edge error / chi2 (in edge order = features with a map point) at the
// initial pose of the first iteration.  Returns the number of edges.
extern "C" int orc_po_debug_linear_system(const orc_pose_problem* p, int robust, double lambda, double* H36, double* b6,
                                          double* x6, double* err, double* chi2) {
  PO po;
  po.P = p;
  for (int i = 0; i < p->n; i++)
    if (p->has_mp[i]) po.edges.push_back(i);
  po.pose = pose_from_Tcw(p->Tcw);
  po.level.assign(po.edges.size(), 0);
  po.err.assign(po.edges.size() * 3, 0.0);
  po.chi2.assign(po.edges.size(), 0.0);
  po.delta_mono = (double)(float)std::sqrt(5.991);
  po.delta_stereo = (double)(float)std::sqrt(7.815);
  po.robust = robust != 0;
  po.compute_active_errors();
  po.build_system();
  std::copy(po.H, po.H + 36, H36);
  std::copy(po.b, po.b + 6, b6);
  std::copy(po.err.begin(), po.err.end(), err);
  std::copy(po.chi2.begin(), po.chi2.end(), chi2);
  if (x6) {
    if (!po.solve(lambda)) return -1;
    std::copy(po.x, po.x + 6, x6);
  }
  return (int)po.edges.size();
}

extern "C" void orc_se3_oplus(const float* Tcw16, const double* upd6, double* R9, double* t3) {
  Pose T = pose_from_Tcw(Tcw16);
  pose_oplus(T, upd6);
  double R[3][3];
  quat_to_R(T.r, R);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) R9[i * 3 + j] = R[i][j];
    t3[i] = T.t[i];
  }
}

extern "C" int orc_local_ba(const orc_ba_problem* p, const volatile uint8_t* stop, orc_ba_result* r) {
  BA ba;
  ba.P = p;
  ba.stop = stop;
  return ba.run(r);
}
