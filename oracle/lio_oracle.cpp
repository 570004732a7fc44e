// =====================================================================================
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// CPU oracle: a plain C++ restatement of the S-FAST_LIO hot path of zhan994/agi_lidar_slam.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product (agi_lidar_slam_b200/csrc) never links or calls it.
//
// PARITY STATUS: the reference ships no tests, fixtures or golden vectors for this path
// (SURVEY.md §4), and Eigen / PCL / Sophus@a621ff are absent from this image, so every
// function below that restates third-party arithmetic is "parity unpinned" against a real
// Eigen/PCL/Sophus build.  What IS pinned: the kNN / map functions are checked against the
// reference's own ikd-Tree compiled in place (oracle/_ref/libikd_ref.so, tests/test_oracle_map.py),
// and the restated math is cross-checked against numpy/scipy (tests/test_oracle_math.py).
//
// Build: see oracle/Makefile  (-O2 -ffp-contract=off -fopenmp; no fast-math, so FP32/FP64
// operations are individually IEEE-rounded in the order written here).
//
// Reference locations restated (paths relative to src/S-FAST_LIO/):
//   boxplus / boxminus            include/esekfom.hpp:59-73, 236-258
//   predict                       include/esekfom.hpp:82-95 ; include/use-ikfom.hpp:40-123
//   h_share_model                 include/esekfom.hpp:106-227
//   update_iterated_dyn_share_... include/esekfom.hpp:270-346
//   esti_plane                    include/common_lib.h:102-134
//   IMU_init                      src/IMU_Processing.hpp:180-244
//   UndistortPcl                  src/IMU_Processing.hpp:253-402
//   VoxelGrid filter (PCL, ext)   src/laserMapping.cpp:683,737-738 ; SURVEY.md App. B.4
//   map_incremental               src/laserMapping.cpp:382-433
//   pointBodyToWorld              src/laserMapping.cpp:277-288
//   ikd-Tree Build/Search/Add     include/ikd-Tree/ikd_Tree.cpp:355-512, 960-1101 (semantics only;
//                                 the data structure here is a plain hashed voxel grid)
// =====================================================================================
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ------------------------------------------------------------------------------------
// Small fixed-size FP64 helpers
// ------------------------------------------------------------------------------------
struct Quat {
  double w, x, y, z;
};

static inline void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double norm3(const double a[3]) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// 3x3 row-major times vector:  (m0*v0 + m1*v1) + m2*v2   — this op order is the oracle's definition.
static inline void mat3_vec(const double m[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; ++i) o[i] = (m[3 * i] * v[0] + m[3 * i + 1] * v[1]) + m[3 * i + 2] * v[2];
}
static inline void mat3T_vec(const double m[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; ++i) o[i] = (m[i] * v[0] + m[3 + i] * v[1]) + m[6 + i] * v[2];
}
static inline void mat3_mul(const double a[9], const double b[9], double o[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * i + j] = (a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j];
}
static inline void mat3T_mul(const double a[9], const double b[9], double o[9]) {  // a^T * b
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * i + j] = (a[i] * b[j] + a[3 + i] * b[3 + j]) + a[6 + i] * b[6 + j];
}

// Old (non-templated) Sophus::SO3 stores a unit quaternion ([ext]).
static inline Quat quat_normalized(Quat q) {
  double n = std::sqrt(((q.w * q.w + q.x * q.x) + q.y * q.y) + q.z * q.z);
  return Quat{q.w / n, q.x / n, q.y / n, q.z / n};
}
// Eigen quaternion product followed by Sophus' renormalisation (SO3::operator*).
static inline Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return quat_normalized(r);
}
// Eigen Quaternion::toRotationMatrix (row-major out).
static inline void quat_to_mat(const Quat& q, double m[9]) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  m[0] = 1.0 - (tyy + tzz);
  m[1] = txy - twz;
  m[2] = txz + twy;
  m[3] = txy + twz;
  m[4] = 1.0 - (txx + tzz);
  m[5] = tyz - twx;
  m[6] = txz - twy;
  m[7] = tyz + twx;
  m[8] = 1.0 - (txx + tyy);
}
// Eigen Quaternion::_transformVector: v + w*uv + qv x uv, uv = 2 (qv x v).   (SO3 * Vector3d)
static inline void quat_rotate(const Quat& q, const double v[3], double o[3]) {
  const double qv[3] = {q.x, q.y, q.z};
  double uv[3], c2[3];
  cross3(qv, v, uv);
  uv[0] += uv[0];
  uv[1] += uv[1];
  uv[2] += uv[2];
  cross3(qv, uv, c2);
  for (int i = 0; i < 3; ++i) o[i] = (v[i] + q.w * uv[i]) + c2[i];
}
// Eigen rotation-matrix -> quaternion (trace method), then Sophus normalises.   (SO3(Matrix3d))
static inline Quat mat_to_quat(const double m[9]) {
  Quat q;
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[7] - m[5]) * t;
    q.y = (m[2] - m[6]) * t;
    q.z = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m[3 * k + j] - m[3 * j + k]) * t;
    v[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    v[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    q.x = v[0];
    q.y = v[1];
    q.z = v[2];
  }
  return quat_normalized(q);
}
// Sophus::SO3::exp  (esekfom.hpp:63-64, IMU_Processing.hpp:382)
static inline Quat so3_exp(const double w[3]) {
  const double theta = norm3(w);
  const double half = 0.5 * theta;
  double imag, real;
  if (theta < 1e-10) {
    const double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
    real = 1.0 - 0.125 * t2 + 0.00260417 * t4;
  } else {
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  return quat_normalized(Quat{real, imag * w[0], imag * w[1], imag * w[2]});
}
// Sophus::SO3::log  (esekfom.hpp:243-246)
static inline void so3_log(const Quat& q, double o[3]) {
  const double v[3] = {q.x, q.y, q.z};
  const double n = norm3(v);
  const double w = q.w;
  double f;
  if (n < 1e-10) {
    f = 2.0 / w - 2.0 * (n * n) / (w * (w * w));
  } else if (std::fabs(w) < 1e-10) {
    f = (w > 0 ? M_PI : -M_PI) / n;
  } else {
    f = 2.0 * std::atan(n / w) / n;
  }
  o[0] = f * v[0];
  o[1] = f * v[1];
  o[2] = f * v[2];
}

// ------------------------------------------------------------------------------------
// State (use-ikfom.hpp:18-27), flat layout shared with the tests:
//   [0:3) pos  [3:7) rot (w,x,y,z)  [7:11) R_LI (w,x,y,z)  [11:14) t_LI  [14:17) vel
//   [17:20) bg  [20:23) ba  [23:26) grav
// Error-state order (24): pos, rot, R_LI, t_LI, vel, bg, ba, grav.
// ------------------------------------------------------------------------------------
struct State {
  double pos[3];
  Quat rot;
  Quat rli;
  double tli[3];
  double vel[3];
  double bg[3];
  double ba[3];
  double grav[3];
};
static_assert(sizeof(State) == 26 * 8, "State must be 26 packed doubles");

static State boxplus(const State& x, const double f[24]) {  // esekfom.hpp:59-73
  State r = x;
  for (int i = 0; i < 3; ++i) r.pos[i] = x.pos[i] + f[i];
  r.rot = quat_mul(x.rot, so3_exp(f + 3));
  r.rli = quat_mul(x.rli, so3_exp(f + 6));
  for (int i = 0; i < 3; ++i) {
    r.tli[i] = x.tli[i] + f[9 + i];
    r.vel[i] = x.vel[i] + f[12 + i];
    r.bg[i] = x.bg[i] + f[15 + i];
    r.ba[i] = x.ba[i] + f[18 + i];
    r.grav[i] = x.grav[i] + f[21 + i];
  }
  return r;
}
static void boxminus(const State& x1, const State& x2, double o[24]) {  // esekfom.hpp:236-258
  double R1[9], R2[9], D[9];
  for (int i = 0; i < 3; ++i) o[i] = x1.pos[i] - x2.pos[i];
  quat_to_mat(x1.rot, R1);
  quat_to_mat(x2.rot, R2);
  mat3T_mul(R2, R1, D);
  so3_log(mat_to_quat(D), o + 3);
  quat_to_mat(x1.rli, R1);
  quat_to_mat(x2.rli, R2);
  mat3T_mul(R2, R1, D);
  so3_log(mat_to_quat(D), o + 6);
  for (int i = 0; i < 3; ++i) {
    o[9 + i] = x1.tli[i] - x2.tli[i];
    o[12 + i] = x1.vel[i] - x2.vel[i];
    o[15 + i] = x1.bg[i] - x2.bg[i];
    o[18 + i] = x1.ba[i] - x2.ba[i];
    o[21 + i] = x1.grav[i] - x2.grav[i];
  }
}

// ------------------------------------------------------------------------------------
// Dense n x n FP64: partial-pivot LU inverse (Eigen's inverse() for n > 4)
// ------------------------------------------------------------------------------------
static bool inverse_lu(const double* A, double* Ainv, int n) {
  std::vector<double> lu(A, A + n * n);
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(lu[k * n + k]);
    for (int i = k + 1; i < n; ++i) {
      double v = std::fabs(lu[i * n + k]);
      if (v > best) {
        best = v;
        piv = i;
      }
    }
    if (best == 0.0) ok = false;
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(lu[k * n + j], lu[piv * n + j]);
      std::swap(perm[k], perm[piv]);
    }
    const double d = lu[k * n + k];
    for (int i = k + 1; i < n; ++i) {
      lu[i * n + k] /= d;
      const double l = lu[i * n + k];
      for (int j = k + 1; j < n; ++j) lu[i * n + j] -= l * lu[k * n + j];
    }
  }
  // Solve L U X = P I, column by column.
  std::vector<double> col(n);
  for (int c = 0; c < n; ++c) {
    for (int i = 0; i < n; ++i) col[i] = (perm[i] == c) ? 1.0 : 0.0;
    for (int i = 0; i < n; ++i) {
      double s = col[i];
      for (int j = 0; j < i; ++j) s -= lu[i * n + j] * col[j];
      col[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = col[i];
      for (int j = n - 1; j > i; --j) s -= lu[i * n + j] * col[j];  // j descending (the oracle's definition)
      col[i] = s / lu[i * n + i];
    }
    for (int i = 0; i < n; ++i) Ainv[i * n + c] = col[i];
  }
  return ok;
}

// ------------------------------------------------------------------------------------
// esti_plane<float>  (common_lib.h:102-134): 5x3 FP32 column-pivoted Householder QR solve of
// A n = -1 (Eigen ColPivHouseholderQR recipe, SURVEY App. B.1, [ext]), normalise, 0.1 m test.
// Every FP32 operation is written out one rounding at a time; the CUDA kernel mirrors this
// order so that the two discontinuous tests select the same points.
// pts: 5 x 3 (row-major).  Returns 1 if plane accepted.
// ------------------------------------------------------------------------------------
static inline float col_norm(const float A[5][3], int j, int r0) {
  float s = 0.f;
  for (int i = r0; i < 5; ++i) s = s + A[i][j] * A[i][j];
  return std::sqrt(s);
}
static bool qr_solve_5x3(const float pts[15], float x[3]) {
  float A[5][3];
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = pts[3 * i + j];
  float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
  float nd[3], nu[3], tau[3];
  int trans[3];
  for (int j = 0; j < 3; ++j) nd[j] = nu[j] = col_norm(A, j, 0);
  float maxn = std::max(nu[0], std::max(nu[1], nu[2]));
  const float eps = FLT_EPSILON;
  const float thr_helper = ((maxn * eps) * (maxn * eps)) / 5.0f;
  const float downdate_thr = std::sqrt(eps);
  int nonzero = 3;
  for (int k = 0; k < 3; ++k) {
    int jb = k;
    float best = nu[k];
    for (int j = k + 1; j < 3; ++j)
      if (nu[j] > best) {
        best = nu[j];
        jb = j;
      }
    if (nonzero == 3 && best * best < thr_helper * (float)(5 - k)) nonzero = k;
    trans[k] = jb;
    if (jb != k) {
      for (int i = 0; i < 5; ++i) std::swap(A[i][k], A[i][jb]);
      std::swap(nu[k], nu[jb]);
      std::swap(nd[k], nd[jb]);
    }
    // Householder on A[k:,k]
    float tailsq = 0.f;
    for (int i = k + 1; i < 5; ++i) tailsq = tailsq + A[i][k] * A[i][k];
    const float c0 = A[k][k];
    float beta, t;
    float ess[5] = {0, 0, 0, 0, 0};
    if (tailsq <= FLT_MIN) {
      t = 0.f;
      beta = c0;
    } else {
      beta = std::sqrt(c0 * c0 + tailsq);
      if (c0 >= 0.f) beta = -beta;
      const float den = c0 - beta;
      for (int i = k + 1; i < 5; ++i) ess[i] = A[i][k] / den;
      t = (beta - c0) / beta;
    }
    tau[k] = t;
    A[k][k] = beta;
    for (int i = k + 1; i < 5; ++i) A[i][k] = ess[i];
    // apply H = I - tau [1;ess][1;ess]^T to the trailing columns
    for (int j = k + 1; j < 3; ++j) {
      float tmp = 0.f;
      for (int i = k + 1; i < 5; ++i) tmp = tmp + ess[i] * A[i][j];
      tmp = tmp + A[k][j];
      A[k][j] = A[k][j] - t * tmp;
      for (int i = k + 1; i < 5; ++i) A[i][j] = A[i][j] - (t * tmp) * ess[i];
    }
    // norm down-dating (LAPACK WN-176)
    for (int j = k + 1; j < 3; ++j) {
      if (nu[j] != 0.f) {
        float tt = std::fabs(A[k][j]) / nu[j];
        tt = (1.f + tt) * (1.f - tt);
        if (tt < 0.f) tt = 0.f;
        const float ratio = nu[j] / nd[j];
        const float t2 = tt * (ratio * ratio);
        if (t2 <= downdate_thr) {
          nd[j] = col_norm(A, j, k + 1);
          nu[j] = nd[j];
        } else {
          nu[j] = nu[j] * std::sqrt(tt);
        }
      }
    }
  }
  // c = Q^T b
  for (int k = 0; k < nonzero; ++k) {
    float tmp = 0.f;
    for (int i = k + 1; i < 5; ++i) tmp = tmp + A[i][k] * c[i];
    tmp = tmp + c[k];
    c[k] = c[k] - tau[k] * tmp;
    for (int i = k + 1; i < 5; ++i) c[i] = c[i] - (tau[k] * tmp) * A[i][k];
  }
  // back substitution on the leading nonzero x nonzero block
  float y[3] = {0.f, 0.f, 0.f};
  for (int i = nonzero - 1; i >= 0; --i) {
    float s = c[i];
    for (int j = i + 1; j < nonzero; ++j) s = s - A[i][j] * y[j];
    y[i] = s / A[i][i];
  }
  // un-permute: apply the column transpositions in reverse
  for (int k = 2; k >= 0; --k)
    if (trans[k] != k) std::swap(y[k], y[trans[k]]);
  x[0] = y[0];
  x[1] = y[1];
  x[2] = y[2];
  return nonzero == 3;
}

static bool esti_plane(float pabcd[4], const float pts[15], float threshold) {
  float nv[3];
  qr_solve_5x3(pts, nv);
  const float n = std::sqrt((nv[0] * nv[0] + nv[1] * nv[1]) + nv[2] * nv[2]);
  pabcd[0] = nv[0] / n;
  pabcd[1] = nv[1] / n;
  pabcd[2] = nv[2] / n;
  pabcd[3] = (float)(1.0 / (double)n);  // common_lib.h:124  "1.0 / n" is double then narrowed
  for (int j = 0; j < 5; ++j) {
    const float d = ((pabcd[0] * pts[3 * j] + pabcd[1] * pts[3 * j + 1]) + pabcd[2] * pts[3 * j + 2]) + pabcd[3];
    if (std::fabs(d) > threshold) return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------
// Forward propagation  (esekfom.hpp:82-95, use-ikfom.hpp:57-123)
// ------------------------------------------------------------------------------------
static void predict(State& x, double* P /*24x24 row-major*/, double dt, const double* Q /*12x12*/,
                    const double acc[3], const double gyro[3]) {
  double R[9];
  quat_to_mat(x.rot, R);
  double f[24] = {0};
  double am[3] = {acc[0] - x.ba[0], acc[1] - x.ba[1], acc[2] - x.ba[2]};
  double a_in[3];
  mat3_vec(R, am, a_in);
  for (int i = 0; i < 3; ++i) {
    f[i] = x.vel[i];
    f[3 + i] = gyro[i] - x.bg[i];
    f[12 + i] = a_in[i] + x.grav[i];
  }
  std::vector<double> Fx(24 * 24, 0.0), Fw(24 * 12, 0.0);
  for (int i = 0; i < 3; ++i) Fx[i * 24 + 12 + i] = 1.0;
  const double hat[9] = {0, -am[2], am[1], am[2], 0, -am[0], -am[1], am[0], 0};
  double Rh[9];
  mat3_mul(R, hat, Rh);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Fx[(12 + i) * 24 + 3 + j] = -Rh[3 * i + j];
      Fx[(12 + i) * 24 + 18 + j] = -R[3 * i + j];
      Fw[(12 + i) * 12 + 3 + j] = -R[3 * i + j];
    }
  for (int i = 0; i < 3; ++i) {
    Fx[(12 + i) * 24 + 21 + i] = 1.0;
    Fx[(3 + i) * 24 + 15 + i] = -1.0;
    Fw[(3 + i) * 12 + i] = -1.0;
    Fw[(15 + i) * 12 + 6 + i] = 1.0;
    Fw[(18 + i) * 12 + 9 + i] = 1.0;
  }
  double fdt[24];
  for (int i = 0; i < 24; ++i) fdt[i] = f[i] * dt;
  x = boxplus(x, fdt);
  for (int i = 0; i < 24 * 24; ++i) Fx[i] *= dt;
  for (int i = 0; i < 24; ++i) Fx[i * 24 + i] += 1.0;
  for (int i = 0; i < 24 * 12; ++i) Fw[i] *= dt;
  std::vector<double> T(24 * 24), Pn(24 * 24), W(24 * 12);
  for (int i = 0; i < 24; ++i)
    for (int j = 0; j < 24; ++j) {
      double s = 0;
      for (int k = 0; k < 24; ++k) s += Fx[i * 24 + k] * P[k * 24 + j];
      T[i * 24 + j] = s;
    }
  for (int i = 0; i < 24; ++i)
    for (int j = 0; j < 12; ++j) {
      double s = 0;
      for (int k = 0; k < 12; ++k) s += Fw[i * 12 + k] * Q[k * 12 + j];
      W[i * 12 + j] = s;
    }
  for (int i = 0; i < 24; ++i)
    for (int j = 0; j < 24; ++j) {
      double s = 0, s2 = 0;
      for (int k = 0; k < 24; ++k) s += T[i * 24 + k] * Fx[j * 24 + k];
      for (int k = 0; k < 12; ++k) s2 += W[i * 12 + k] * Fw[j * 12 + k];
      Pn[i * 24 + j] = s + s2;
    }
  std::memcpy(P, Pn.data(), sizeof(double) * 576);
}

// Pose6D.msg restated as a POD: offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9]  (22 doubles)
struct Pose6D {
  double offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9];
};
static_assert(sizeof(Pose6D) == 22 * 8, "Pose6D must be 22 packed doubles");

struct ImuSample {
  double t, acc[3], gyr[3];
};
static_assert(sizeof(ImuSample) == 7 * 8, "ImuSample must be 7 packed doubles");

// Persistent ImuProcess members the forward pass reads/writes (IMU_Processing.hpp:95-136).
struct ImuCarry {
  double cov_gyr[3], cov_acc[3], cov_bias_gyr[3], cov_bias_acc[3];
  double mean_acc_norm;
  double acc_s_last[3], angvel_last[3];
  double last_lidar_end_time;
  ImuSample last_imu;
};
static_assert(sizeof(ImuCarry) == (12 + 1 + 6 + 1 + 7) * 8, "ImuCarry layout");

// UndistortPcl forward half, IMU_Processing.hpp:258-358.  `imu` are this scan's samples (meas.imu);
// carry.last_imu is pushed to the front exactly as :259.
static int imu_forward(const ImuSample* imu, int n_imu, double pcl_beg_time, double pcl_end_time, State& x, double* P,
                       ImuCarry& c, Pose6D* poses, int cap) {
  std::vector<ImuSample> v;
  v.push_back(c.last_imu);
  for (int i = 0; i < n_imu; ++i) v.push_back(imu[i]);
  const double imu_end_time = v.back().t;
  int np = 0;
  auto push_pose = [&](double t, const double a[3], const double g[3], const State& s) {
    if (np >= cap) return;
    Pose6D& p = poses[np++];
    p.offset_time = t;
    for (int i = 0; i < 3; ++i) {
      p.acc[i] = a[i];
      p.gyr[i] = g[i];
      p.vel[i] = s.vel[i];
      p.pos[i] = s.pos[i];
    }
    quat_to_mat(s.rot, p.rot);
  };
  push_pose(0.0, c.acc_s_last, c.angvel_last, x);
  double Q[144] = {0};
  double in_acc[3] = {0, 0, 0}, in_gyr[3] = {0, 0, 0};
  // process_noise_cov() defaults (use-ikfom.hpp:40-48) are overwritten on the diagonal each step (:319-322).
  double dt = 0;
  for (size_t k = 0; k + 1 < v.size(); ++k) {
    const ImuSample& head = v[k];
    const ImuSample& tail = v[k + 1];
    if (tail.t < c.last_lidar_end_time) continue;
    double angvel_avr[3], acc_avr[3];
    for (int i = 0; i < 3; ++i) {
      angvel_avr[i] = 0.5 * (head.gyr[i] + tail.gyr[i]);
      acc_avr[i] = 0.5 * (head.acc[i] + tail.acc[i]);
      acc_avr[i] = acc_avr[i] * 9.81 / c.mean_acc_norm;
    }
    if (head.t < c.last_lidar_end_time)
      dt = tail.t - c.last_lidar_end_time;
    else
      dt = tail.t - head.t;
    for (int i = 0; i < 3; ++i) {
      in_acc[i] = acc_avr[i];
      in_gyr[i] = angvel_avr[i];
      Q[(0 + i) * 12 + 0 + i] = c.cov_gyr[i];
      Q[(3 + i) * 12 + 3 + i] = c.cov_acc[i];
      Q[(6 + i) * 12 + 6 + i] = c.cov_bias_gyr[i];
      Q[(9 + i) * 12 + 9 + i] = c.cov_bias_acc[i];
    }
    predict(x, P, dt, Q, in_acc, in_gyr);
    double a_s[3], tmp[3];
    for (int i = 0; i < 3; ++i) {
      c.angvel_last[i] = tail.gyr[i] - x.bg[i];
      a_s[i] = tail.acc[i] * 9.81 / c.mean_acc_norm;
      tmp[i] = a_s[i] - x.ba[i];
    }
    quat_rotate(x.rot, tmp, a_s);
    for (int i = 0; i < 3; ++i) c.acc_s_last[i] = a_s[i] + x.grav[i];
    push_pose(tail.t - pcl_beg_time, c.acc_s_last, c.angvel_last, x);
  }
  dt = std::fabs(pcl_end_time - imu_end_time);
  for (int i = 0; i < 3; ++i) {  // Q diagonal as left by the loop (or the defaults when the loop never ran)
    if (v.size() < 2) {
      Q[(0 + i) * 12 + 0 + i] = 0.0001;
      Q[(3 + i) * 12 + 3 + i] = 0.0001;
      Q[(6 + i) * 12 + 6 + i] = 0.00001;
      Q[(9 + i) * 12 + 9 + i] = 0.00001;
    }
  }
  predict(x, P, dt, Q, in_acc, in_gyr);
  if (n_imu > 0) c.last_imu = imu[n_imu - 1];
  c.last_lidar_end_time = pcl_end_time;
  return np;
}

// UndistortPcl backward half, IMU_Processing.hpp:361-401, literal (including the `break` quirk at :399).
// pts: N x 4 (x,y,z,curvature[ms]) ALREADY sorted by curvature ascending (the caller sorts, :269).
static void undistort_backward(float* pts, int64_t n, const Pose6D* poses, int n_poses, const State& end) {
  if (n == 0 || n_poses < 2) return;
  double Rli[9], Rend[9];
  quat_to_mat(end.rli, Rli);
  quat_to_mat(end.rot, Rend);
  int64_t it = n - 1;
  for (int kp = n_poses - 1; kp != 0; --kp) {
    const Pose6D& head = poses[kp - 1];
    const Pose6D& tail = poses[kp];
    for (; (double)pts[4 * it + 3] / double(1000) > head.offset_time; --it) {
      const double dt = (double)pts[4 * it + 3] / double(1000) - head.offset_time;
      const double wdt[3] = {tail.gyr[0] * dt, tail.gyr[1] * dt, tail.gyr[2] * dt};
      double E[9], Ri[9];
      quat_to_mat(so3_exp(wdt), E);
      mat3_mul(head.rot, E, Ri);
      const double Pi[3] = {pts[4 * it], pts[4 * it + 1], pts[4 * it + 2]};
      double T_ei[3];
      for (int i = 0; i < 3; ++i)
        T_ei[i] = ((head.pos[i] + head.vel[i] * dt) + ((0.5 * tail.acc[i]) * dt) * dt) - end.pos[i];
      double a[3], b[3], cvec[3], d[3], e[3];
      mat3_vec(Rli, Pi, a);
      for (int i = 0; i < 3; ++i) a[i] += end.tli[i];
      mat3_vec(Ri, a, b);
      for (int i = 0; i < 3; ++i) b[i] += T_ei[i];
      mat3T_vec(Rend, b, cvec);
      for (int i = 0; i < 3; ++i) d[i] = cvec[i] - end.tli[i];
      mat3T_vec(Rli, d, e);
      pts[4 * it] = (float)e[0];
      pts[4 * it + 1] = (float)e[1];
      pts[4 * it + 2] = (float)e[2];
      if (it == 0) break;
    }
  }
}

// ------------------------------------------------------------------------------------
// pcl::VoxelGrid<PointXYZINormal>::applyFilter restated ([ext]).
// in : N x 5  (x,y,z,intensity,curvature);  out: M x 5 centroids ascending by (kz,ky,kx).
// Sums are FP32, sequential in ascending input index within a voxel (PCL's std::sort is unstable, so
// PCL itself does not define the order; this is the oracle's definition).
// Returns M, or -1 on PCL's "leaf size too small" overflow (output == input in PCL).
// ------------------------------------------------------------------------------------
static int64_t voxel_grid(const float* in, int64_t n, float leaf, float* out, int32_t* keys_out, int32_t* point_keys) {
  if (n == 0) return 0;
  const float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], in[5 * i + a]);
      mx[a] = std::max(mx[a], in[5 * i + a]);
    }
  int64_t d[3];
  for (int a = 0; a < 3; ++a) d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
  if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) return -1;
  int minb[3], maxb[3], divb[3];
  for (int a = 0; a < 3; ++a) {
    minb[a] = (int)std::floor(mn[a] * inv);
    maxb[a] = (int)std::floor(mx[a] * inv);
    divb[a] = maxb[a] - minb[a] + 1;
  }
  const int mul[3] = {1, divb[0], divb[0] * divb[1]};
  std::vector<std::pair<int, int64_t>> iv((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    int ijk[3];
    for (int a = 0; a < 3; ++a) ijk[a] = (int)(std::floor(in[5 * i + a] * inv) - (float)minb[a]);
    iv[i] = {ijk[0] * mul[0] + ijk[1] * mul[1] + ijk[2] * mul[2], i};
    if (point_keys)
      for (int a = 0; a < 3; ++a) point_keys[3 * i + a] = ijk[a] + minb[a];
  }
  std::stable_sort(iv.begin(), iv.end(),
                   [](const std::pair<int, int64_t>& a, const std::pair<int, int64_t>& b) { return a.first < b.first; });
  int64_t m = 0;
  size_t s = 0;
  while (s < iv.size()) {
    size_t e = s + 1;
    while (e < iv.size() && iv[e].first == iv[s].first) ++e;
    float acc[5] = {0, 0, 0, 0, 0};
    for (size_t k = s; k < e; ++k)
      for (int f = 0; f < 5; ++f) acc[f] = acc[f] + in[5 * iv[k].second + f];
    const float cnt = (float)(e - s);
    for (int f = 0; f < 5; ++f) out[5 * m + f] = acc[f] / cnt;
    if (keys_out) {
      int idx = iv[s].first;
      keys_out[3 * m + 2] = idx / mul[2] + minb[2];
      idx %= mul[2];
      keys_out[3 * m + 1] = idx / mul[1] + minb[1];
      keys_out[3 * m + 0] = idx % mul[1] + minb[0];
    }
    ++m;
    s = e;
  }
  return m;
}

// ------------------------------------------------------------------------------------
// Map: the OBSERVABLE semantics of the reference ikd-Tree (Build / Nearest_Search / Add_Points /
// Delete_Point_Boxes), on a hashed uniform grid.  Results are canonical: neighbours ascending by
// (d2 FP32 bits, point id).  d2 is the un-fused FP32 expression of ikd_Tree.cpp:1539-1544.
// Point ids: Build -> 0..n-1 in input order; Add_Points -> next_id + input index (ids are unique and
// increasing in insertion order; gaps are allowed).  The reference tree has no ids;
// tests compare coordinates against it.
// ------------------------------------------------------------------------------------
struct MapPoint {
  float x, y, z;
  int32_t id;
};
struct CellKey {
  int32_t x, y, z;
  bool operator==(const CellKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct CellHash {
  size_t operator()(const CellKey& k) const {
    uint64_t h = (uint64_t)(uint32_t)k.x * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)k.y + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= ((uint64_t)(uint32_t)k.z + 0x165667B1ull) * 0xD6E8FEB86659FD93ull;
    return (size_t)(h ^ (h >> 29));
  }
};
static inline float dist2_f(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return (dx * dx + dy * dy) + dz * dz;
}
struct Map {
  float cell;  // grid cell edge (search acceleration only; does not change results)
  float inv_cell;
  int32_t next_id = 0;
  int64_t n_live = 0;
  std::unordered_map<CellKey, std::vector<MapPoint>, CellHash> cells;
  int32_t lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};  // cells ever used
  explicit Map(float c) : cell(c), inv_cell(1.0f / c) {}
  CellKey key_of(float x, float y, float z) const {
    return CellKey{(int32_t)std::floor(x * inv_cell), (int32_t)std::floor(y * inv_cell),
                   (int32_t)std::floor(z * inv_cell)};
  }
  void clear() {
    cells.clear();
    next_id = 0;
    n_live = 0;
    for (int a = 0; a < 3; ++a) {
      lo[a] = INT32_MAX;
      hi[a] = INT32_MIN;
    }
  }
  void insert(float x, float y, float z, int32_t id) {
    const CellKey k = key_of(x, y, z);
    const int32_t kk[3] = {k.x, k.y, k.z};
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], kk[a]);
      hi[a] = std::max(hi[a], kk[a]);
    }
    cells[k].push_back(MapPoint{x, y, z, id});
    ++n_live;
  }
  void build(const float* xyz, int64_t n) {
    clear();
    for (int64_t i = 0; i < n; ++i) insert(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], (int32_t)i);
    next_id = (int32_t)n;
  }
  // Exact k-NN among points with d2 <= max_d2 (the reference's `dist <= max_dist_sqr`, ikd_Tree.cpp:980),
  // canonical order.  max_d2 = INFINITY is the reference's default (ikd_Tree.h:285; what esekfom.hpp:140 gets): the
  // shells then grow until the k-th best beats them or every cell ever used has been visited.
  // Returns count found (<= k).
  int knn(const float q[3], int k, float max_d2, MapPoint* out, float* out_d2) const {
    struct Cand {
      float d2;
      int32_t id;
      MapPoint p;
    };
    std::vector<Cand> best;  // kept sorted, size <= k
    const CellKey c = key_of(q[0], q[1], q[2]);
    int rmax = 0, r0 = 0;
    const int32_t cc[3] = {c.x, c.y, c.z};
    for (int a = 0; a < 3; ++a)
      if (lo[a] <= hi[a]) {
        rmax = std::max(rmax, std::max(std::abs(cc[a] - lo[a]), std::abs(cc[a] - hi[a])));  // farthest used cell
        r0 = std::max(r0, std::max(lo[a] - cc[a], cc[a] - hi[a]));                          // nearest used cell
      }
    if (!std::isinf(max_d2)) rmax = std::min(rmax, (int)std::ceil(std::sqrt((double)max_d2) * (double)inv_cell) + 1);
    auto visit = [&](int dx, int dy, int dz) {
      const CellKey ck{c.x + dx, c.y + dy, c.z + dz};
      if (ck.x < lo[0] || ck.x > hi[0] || ck.y < lo[1] || ck.y > hi[1] || ck.z < lo[2] || ck.z > hi[2]) return;
      auto it = cells.find(ck);
      if (it == cells.end()) return;
      for (const MapPoint& p : it->second) {
        const float d2 = dist2_f(q[0], q[1], q[2], p.x, p.y, p.z);
        if (!(d2 <= max_d2)) continue;
        Cand cd{d2, p.id, p};
        auto less = [](const Cand& a, const Cand& b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.id < b.id); };
        if ((int)best.size() == k && !less(cd, best.back())) continue;
        best.insert(std::upper_bound(best.begin(), best.end(), cd, less), cd);
        if ((int)best.size() > k) best.pop_back();
      }
    };
    if (std::isinf(max_d2) && r0 > 64) {
      // query far outside everything the map holds: walking empty shells would cost more than looking at every cell
      for (const auto& kv : cells) visit(kv.first.x - c.x, kv.first.y - c.y, kv.first.z - c.z);
      rmax = -1;
    }
    for (int r = r0; r <= rmax; ++r) {
      // shell r can only hold points at distance >= (r-1)*cell from q; stop once the k-th best beats it
      if (r >= 1 && (int)best.size() == k) {
        const double cover = (double)(r - 1) * (double)cell;
        if ((double)best.back().d2 < cover * cover) break;
      }
      // the cells at Chebyshev distance exactly r: two full faces, and the perimeter of every layer between them
      for (int dz = -r; dz <= r; ++dz) {
        if (dz == -r || dz == r) {
          for (int dy = -r; dy <= r; ++dy)
            for (int dx = -r; dx <= r; ++dx) visit(dx, dy, dz);
        } else {
          for (int dx = -r; dx <= r; ++dx) {
            visit(dx, -r, dz);
            visit(dx, r, dz);
          }
          for (int dy = -r + 1; dy <= r - 1; ++dy) {
            visit(-r, dy, dz);
            visit(r, dy, dz);
          }
        }
      }
    }
    for (size_t i = 0; i < best.size(); ++i) {
      out[i] = best[i].p;
      out_d2[i] = best[i].d2;
    }
    return (int)best.size();
  }
  // Add_Points (ikd_Tree.cpp:419-512).  With downsample: per voxel [floor(x/ds)*ds, +ds) the reference's
  // sequential rule (winner = nearest to the voxel centre among {new, existing}, a new point wins ties; the box
  // is collapsed to the winner whenever it held >1 point or the winner is new) is order-independent up to
  // exact ties, and is restated here per batch: w = argmin over existing U new of (dist to centre), new beats
  // existing on ties, later new beats earlier new, lower id beats higher id among existing.  If the voxel holds
  // <= 1 existing point and w is that point nothing changes; otherwise the voxel becomes {w} (an existing
  // winner keeps its id, a new winner gets id next_id + its input index).
  // Returns the number of new points that ended up in the map.
  int add_points(const float* xyz, int64_t n, bool downsample_on, float ds) {
    int added = 0;
    if (!downsample_on) {
      for (int64_t i = 0; i < n; ++i) insert(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], next_id + (int32_t)i);
      next_id += (int32_t)n;
      return (int)n;
    }
    struct VoxBest {
      float bmin[3];
      float dist;
      int64_t idx;
    };
    struct VKey {
      uint32_t bx, by, bz;  // bit patterns of the FP32 box minimum: the voxel's identity
      bool operator==(const VKey& o) const { return bx == o.bx && by == o.by && bz == o.bz; }
    };
    struct VHash {
      size_t operator()(const VKey& k) const {
        return (size_t)(((uint64_t)k.bx * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)k.by * 0xC2B2AE3D27D4EB4Full) ^
                        ((uint64_t)k.bz * 0xD6E8FEB86659FD93ull));
      }
    };
    std::unordered_map<VKey, VoxBest, VHash> vox;
    std::vector<VKey> order;
    for (int64_t i = 0; i < n; ++i) {
      const float* p = xyz + 3 * i;
      float bmin[3], mid[3];
      for (int a = 0; a < 3; ++a) {
        bmin[a] = std::floor(p[a] / ds) * ds + 0.0f;  // +0 folds -0 into +0 (same voxel)
        const float bmax = bmin[a] + ds;
        mid[a] = (float)((double)bmin[a] + (double)(bmax - bmin[a]) / 2.0);
      }
      const float d = dist2_f(p[0], p[1], p[2], mid[0], mid[1], mid[2]);
      VKey k;
      std::memcpy(&k.bx, &bmin[0], 4);
      std::memcpy(&k.by, &bmin[1], 4);
      std::memcpy(&k.bz, &bmin[2], 4);
      auto it = vox.find(k);
      if (it == vox.end()) {
        vox.emplace(k, VoxBest{{bmin[0], bmin[1], bmin[2]}, d, i});
        order.push_back(k);
      } else if (d <= it->second.dist) {  // later new point wins ties
        it->second.dist = d;
        it->second.idx = i;
      }
    }
    for (const VKey& k : order) {
      const VoxBest& vb = vox[k];
      float bmax[3], mid[3];
      for (int a = 0; a < 3; ++a) {
        bmax[a] = vb.bmin[a] + ds;
        mid[a] = (float)((double)vb.bmin[a] + (double)(bmax[a] - vb.bmin[a]) / 2.0);
      }
      // existing points inside the half-open box
      struct Ref {
        std::vector<MapPoint>* cellv;
        size_t i;
      };
      std::vector<Ref> inside;
      const CellKey c0 = key_of(vb.bmin[0], vb.bmin[1], vb.bmin[2]);
      const CellKey c1 = key_of(bmax[0], bmax[1], bmax[2]);
      for (int cz = c0.z; cz <= c1.z; ++cz)
        for (int cy = c0.y; cy <= c1.y; ++cy)
          for (int cx = c0.x; cx <= c1.x; ++cx) {
            auto it = cells.find(CellKey{cx, cy, cz});
            if (it == cells.end()) continue;
            for (size_t i = 0; i < it->second.size(); ++i) {
              const MapPoint& p = it->second[i];
              if (vb.bmin[0] <= p.x && bmax[0] > p.x && vb.bmin[1] <= p.y && bmax[1] > p.y && vb.bmin[2] <= p.z &&
                  bmax[2] > p.z)
                inside.push_back(Ref{&it->second, i});
            }
          }
      // best existing: strictly closer than the new winner; lower id wins among equals
      int best_e = -1;
      float best_d = vb.dist;
      int32_t best_id = 0;
      for (size_t e = 0; e < inside.size(); ++e) {
        const MapPoint& p = (*inside[e].cellv)[inside[e].i];
        const float d = dist2_f(p.x, p.y, p.z, mid[0], mid[1], mid[2]);
        if (d < best_d || (best_e >= 0 && d == best_d && p.id < best_id)) {
          best_d = d;
          best_e = (int)e;
          best_id = p.id;
        }
      }
      if (best_e >= 0 && inside.size() <= 1) continue;  // single existing point survives, nothing changes
      // collapse the voxel to the winner
      MapPoint keep;
      if (best_e >= 0) {
        keep = (*inside[best_e].cellv)[inside[best_e].i];
      } else {
        keep = MapPoint{xyz[3 * vb.idx], xyz[3 * vb.idx + 1], xyz[3 * vb.idx + 2], next_id + (int32_t)vb.idx};
        ++added;
      }
      // erase existing (mark then compact)
      for (const Ref& r : inside) (*r.cellv)[r.i].id = -1;
      n_live -= (int64_t)inside.size();
      for (int cz = c0.z; cz <= c1.z; ++cz)
        for (int cy = c0.y; cy <= c1.y; ++cy)
          for (int cx = c0.x; cx <= c1.x; ++cx) {
            auto it = cells.find(CellKey{cx, cy, cz});
            if (it == cells.end()) continue;
            auto& v = it->second;
            v.erase(std::remove_if(v.begin(), v.end(), [](const MapPoint& p) { return p.id < 0; }), v.end());
          }
      insert(keep.x, keep.y, keep.z, keep.id);
    }
    next_id += (int32_t)n;
    return added;
  }
  // Delete_Point_Boxes (ikd_Tree.cpp:559-579; half-open boxes, :685-699,713-718).  Returns #deleted.
  int delete_boxes(const float* boxes6, int nb) {
    int deleted = 0;
    for (auto& kv : cells) {
      auto& v = kv.second;
      size_t before = v.size();
      v.erase(std::remove_if(v.begin(), v.end(),
                             [&](const MapPoint& p) {
                               for (int b = 0; b < nb; ++b) {
                                 const float* mn = boxes6 + 6 * b;
                                 const float* mx = mn + 3;
                                 if (mn[0] <= p.x && mx[0] > p.x && mn[1] <= p.y && mx[1] > p.y && mn[2] <= p.z &&
                                     mx[2] > p.z)
                                   return true;
                               }
                               return false;
                             }),
              v.end());
      deleted += (int)(before - v.size());
    }
    n_live -= deleted;
    return deleted;
  }
  int64_t dump(float* xyz, int32_t* ids, int64_t cap) const {
    std::vector<MapPoint> all;
    for (auto& kv : cells)
      for (auto& p : kv.second) all.push_back(p);
    std::sort(all.begin(), all.end(), [](const MapPoint& a, const MapPoint& b) { return a.id < b.id; });
    for (int64_t i = 0; i < (int64_t)all.size() && i < cap; ++i) {
      if (xyz) {
        xyz[3 * i] = all[i].x;
        xyz[3 * i + 1] = all[i].y;
        xyz[3 * i + 2] = all[i].z;
      }
      if (ids) ids[i] = all[i].id;
    }
    return (int64_t)all.size();
  }
};

// kNN back-end used by h_share_model: either the Map above or the reference ikd-Tree
// (function pointer taken from oracle/_ref/libikd_ref.so by the caller).
// out_pts: k x 4 floats (x,y,z, id-as-int-bits); returns count found.
typedef int (*knn_fn)(void* ctx, const float* q, int k, float max_d2, float* out_pts, float* out_d2);

static int map_knn_cb(void* ctx, const float* q, int k, float max_d2, float* out_pts, float* out_d2) {
  MapPoint tmp[16];
  const int n = static_cast<Map*>(ctx)->knn(q, k, max_d2, tmp, out_d2);
  for (int i = 0; i < n; ++i) {
    out_pts[4 * i] = tmp[i].x;
    out_pts[4 * i + 1] = tmp[i].y;
    out_pts[4 * i + 2] = tmp[i].z;
    std::memcpy(&out_pts[4 * i + 3], &tmp[i].id, 4);
  }
  return n;
}

// ------------------------------------------------------------------------------------
// h_share_model (esekfom.hpp:106-227).  Per-scan persistent arrays are owned by the caller (Scan).
// ------------------------------------------------------------------------------------
struct Scan {
  int64_t m = 0;
  const float* body = nullptr;      // M x 3
  std::vector<float> world;         // M x 3   (FP32 p_world of the last pass)
  std::vector<float> near_pts;      // M x 5 x 4  cached Nearest_Points (xyz + id bits)
  std::vector<float> near_d2;       // M x 5
  std::vector<int32_t> near_cnt;    // M
  std::vector<uint8_t> selected;    // M  point_selected_surf
  std::vector<float> normvec;       // M x 4  (a,b,c,pd2)
  std::vector<double> h_x;          // V x 12
  std::vector<double> h;            // V
  std::vector<int32_t> valid_index; // V  (ascending i)
};

struct Params {
  float max_d2 = INFINITY;    // Nearest_Search's max_dist: esekfom.hpp:140-141 passes none => ikd_Tree.h:285 default
  float plane_thr = 0.1f;     // esekfom.hpp:157
  int k = 5;                  // NUM_MATCH_POINTS
  int threads = 1;            // MP_PROC_NUM
};

static bool h_share_model(const State& x, Scan& s, bool converge, bool extrinsic_est, knn_fn knn, void* knn_ctx,
                          const Params& prm) {
  const int64_t M = s.m;
  const int K = prm.k;
#pragma omp parallel for num_threads(prm.threads) schedule(static)
  for (int64_t i = 0; i < M; ++i) {
    const double pb[3] = {s.body[3 * i], s.body[3 * i + 1], s.body[3 * i + 2]};
    double pi[3], pg[3];
    quat_rotate(x.rli, pb, pi);
    for (int a = 0; a < 3; ++a) pi[a] += x.tli[a];
    quat_rotate(x.rot, pi, pg);
    float pw[3];
    for (int a = 0; a < 3; ++a) {
      pg[a] += x.pos[a];
      pw[a] = (float)pg[a];
      s.world[3 * i + a] = pw[a];
    }
    if (converge) {
      float* np = &s.near_pts[(size_t)i * K * 4];
      float* nd = &s.near_d2[(size_t)i * K];
      for (int j = 0; j < K; ++j) {
        np[4 * j] = np[4 * j + 1] = np[4 * j + 2] = 0.f;
        int32_t neg = -1;
        std::memcpy(&np[4 * j + 3], &neg, 4);
        nd[j] = INFINITY;
      }
      // esekfom.hpp:140-141: no max_dist argument, so Nearest_Points[i] holds min(5, #live map points) neighbours
      // however far away they are; map_incremental reads them later (laserMapping.cpp:391-423) whatever gate 1 says.
      const int cnt = knn(knn_ctx, pw, K, prm.max_d2, np, nd);
      s.near_cnt[i] = cnt;
      s.selected[i] = (cnt < K) ? 0 : (nd[K - 1] > 5 ? 0 : 1);
    }
    if (!s.selected[i]) continue;
    s.selected[i] = 0;
    float pts[15];
    for (int j = 0; j < 5; ++j)
      for (int a = 0; a < 3; ++a) pts[3 * j + a] = s.near_pts[((size_t)i * K + j) * 4 + a];
    float pabcd[4];
    if (esti_plane(pabcd, pts, prm.plane_thr)) {
      const float pd2 = ((pabcd[0] * pw[0] + pabcd[1] * pw[1]) + pabcd[2] * pw[2]) + pabcd[3];
      // esekfom.hpp:163  float s = 1 - 0.9 * fabs(pd2) / sqrt(p_body.norm());   (double arithmetic, narrowed)
      const float sc = (float)(1.0 - 0.9 * std::fabs((double)pd2) / std::sqrt(norm3(pb)));
      if ((double)sc > 0.9) {
        s.selected[i] = 1;
        s.normvec[4 * i] = pabcd[0];
        s.normvec[4 * i + 1] = pabcd[1];
        s.normvec[4 * i + 2] = pabcd[2];
        s.normvec[4 * i + 3] = pd2;
      }
    }
  }
  s.valid_index.clear();
  for (int64_t i = 0; i < M; ++i)
    if (s.selected[i]) s.valid_index.push_back((int32_t)i);
  const int64_t V = (int64_t)s.valid_index.size();
  if (V < 1) return false;
  s.h_x.assign((size_t)V * 12, 0.0);
  s.h.assign((size_t)V, 0.0);
  double Rt[9], Rli[9];
  quat_to_mat(x.rot, Rt);
  quat_to_mat(x.rli, Rli);
  for (int64_t r = 0; r < V; ++r) {
    const int64_t i = s.valid_index[r];
    const double p[3] = {s.body[3 * i], s.body[3 * i + 1], s.body[3 * i + 2]};
    double pI[3];
    quat_rotate(x.rli, p, pI);
    for (int a = 0; a < 3; ++a) pI[a] += x.tli[a];
    const double nv[3] = {s.normvec[4 * i], s.normvec[4 * i + 1], s.normvec[4 * i + 2]};
    double C[3], A[3];
    mat3T_vec(Rt, nv, C);
    const double pIx[9] = {0.0, -pI[2], pI[1], pI[2], 0.0, -pI[0], -pI[1], pI[0], 0.0};
    mat3_vec(pIx, C, A);
    double* row = &s.h_x[(size_t)r * 12];
    row[0] = nv[0];
    row[1] = nv[1];
    row[2] = nv[2];
    row[3] = A[0];
    row[4] = A[1];
    row[5] = A[2];
    if (extrinsic_est) {
      const double px[9] = {0.0, -p[2], p[1], p[2], 0.0, -p[0], -p[1], p[0], 0.0};
      double M1[9], B[3];
      // point_crossmat * R_LI^T * C  evaluated left to right: (px * Rli^T) * C
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) M1[3 * a + b] = (px[3 * a] * Rli[3 * b] + px[3 * a + 1] * Rli[3 * b + 1]) + px[3 * a + 2] * Rli[3 * b + 2];
      mat3_vec(M1, C, B);
      row[6] = B[0];
      row[7] = B[1];
      row[8] = B[2];
      row[9] = C[0];
      row[10] = C[1];
      row[11] = C[2];
    }
    s.h[r] = -(double)s.normvec[4 * i + 3];
  }
  return true;
}

// One pass record for stage-wise parity (tests compare these against the CUDA path).
struct PassTrace {
  int32_t searched, valid, n_valid, converged;
  double blob[90];  // HtH upper triangle (row-major, 78) + Hth (12)
  double dx[24];
};

// update_iterated_dyn_share_modified (esekfom.hpp:270-346), literal dense algebra.
static int update_iterated(State& x, double* P, double R, Scan& s, knn_fn knn, void* knn_ctx, int maximum_iter,
                           bool extrinsic_est, const Params& prm, PassTrace* trace, int trace_cap, int* n_valid_last) {
  std::fill(s.normvec.begin(), s.normvec.end(), 0.f);
  bool converge = true;
  int t = 0;
  const State x_prop = x;
  int passes = 0;
  if (n_valid_last) *n_valid_last = 0;
  for (int it = -1; it < maximum_iter; ++it) {
    PassTrace tr;
    std::memset(&tr, 0, sizeof(tr));
    tr.searched = converge ? 1 : 0;
    const bool valid = h_share_model(x, s, converge, extrinsic_est, knn, knn_ctx, prm);
    tr.valid = valid ? 1 : 0;
    tr.n_valid = valid ? (int32_t)s.valid_index.size() : 0;
    if (n_valid_last) *n_valid_last = tr.n_valid;
    if (!valid) {
      if (trace && passes < trace_cap) trace[passes] = tr;
      ++passes;
      continue;
    }
    const int64_t V = (int64_t)s.valid_index.size();
    double dx_new[24];
    boxminus(x, x_prop, dx_new);
    const std::vector<double>& H = s.h_x;
    double HTH[576] = {0};
    for (int a = 0; a < 12; ++a)
      for (int b = 0; b < 12; ++b) {
        double acc = 0;
        for (int64_t r = 0; r < V; ++r) acc += H[r * 12 + a] * H[r * 12 + b];
        HTH[a * 24 + b] = acc;
      }
    {
      int e = 0;
      for (int a = 0; a < 12; ++a)
        for (int b = a; b < 12; ++b) tr.blob[e++] = HTH[a * 24 + b];
      for (int a = 0; a < 12; ++a) {
        double acc = 0;
        for (int64_t r = 0; r < V; ++r) acc += H[r * 12 + a] * s.h[r];
        tr.blob[78 + a] = acc;
      }
    }
    double Pinv[576], A[576], Kf[576];
    inverse_lu(P, Pinv, 24);
    for (int i = 0; i < 576; ++i) A[i] = HTH[i] / R + Pinv[i];
    inverse_lu(A, Kf, 24);
    std::vector<double> K((size_t)24 * V);
    for (int r = 0; r < 24; ++r)
      for (int64_t i = 0; i < V; ++i) {
        double acc = 0;
        for (int c = 0; c < 12; ++c) acc += Kf[r * 24 + c] * H[i * 12 + c];
        K[(size_t)r * V + i] = acc / R;
      }
    double KH[576] = {0};
    for (int r = 0; r < 24; ++r)
      for (int c = 0; c < 12; ++c) {
        double acc = 0;
        for (int64_t i = 0; i < V; ++i) acc += K[(size_t)r * V + i] * H[i * 12 + c];
        KH[r * 24 + c] = acc;
      }
    double dx[24];
    for (int r = 0; r < 24; ++r) {
      double kh = 0;
      for (int64_t i = 0; i < V; ++i) kh += K[(size_t)r * V + i] * s.h[i];
      double acc = 0;
      for (int c = 0; c < 24; ++c) acc += (KH[r * 24 + c] - (r == c ? 1.0 : 0.0)) * dx_new[c];
      dx[r] = kh + acc;
    }
    x = boxplus(x, dx);
    converge = true;
    for (int j = 0; j < 24; ++j)
      if (std::fabs(dx[j]) > 0.001) {
        converge = false;
        break;
      }
    tr.converged = converge ? 1 : 0;
    std::memcpy(tr.dx, dx, sizeof(dx));
    if (trace && passes < trace_cap) trace[passes] = tr;
    ++passes;
    if (converge) t++;
    if (!t && it == maximum_iter - 2) converge = true;
    if (t > 1 || it == maximum_iter - 1) {
      double Pn[576];
      for (int r = 0; r < 24; ++r)
        for (int c = 0; c < 24; ++c) {
          double acc = 0;
          for (int k = 0; k < 24; ++k) acc += ((r == k ? 1.0 : 0.0) - KH[r * 24 + k]) * P[k * 24 + c];
          Pn[r * 24 + c] = acc;
        }
      std::memcpy(P, Pn, sizeof(Pn));
      return passes;
    }
  }
  return passes;
}

// pointBodyToWorld (laserMapping.cpp:277-288): matrix form, FP64 -> FP32.
static void body_to_world(const State& x, const float* body, int64_t m, float* world) {
  double R[9], Rli[9];
  quat_to_mat(x.rot, R);
  quat_to_mat(x.rli, Rli);
  for (int64_t i = 0; i < m; ++i) {
    const double p[3] = {body[3 * i], body[3 * i + 1], body[3 * i + 2]};
    double a[3], b[3];
    mat3_vec(Rli, p, a);
    for (int k = 0; k < 3; ++k) a[k] += x.tli[k];
    mat3_vec(R, a, b);
    for (int k = 0; k < 3; ++k) world[3 * i + k] = (float)(b[k] + x.pos[k]);
  }
}

// map_incremental policy (laserMapping.cpp:382-433).  class_out[i]: 0 = skip, 1 = PointToAdd (downsample),
// 2 = PointNoNeedDownsample.
static void map_incremental_classify(const float* world, int64_t m, const float* near_pts /*M x 5 x 4*/,
                                     const int32_t* near_cnt, bool ekf_inited, float fsm /*filter_size_map_min*/,
                                     uint8_t* class_out) {
  const double fs = (double)fsm;
  for (int64_t i = 0; i < m; ++i) {
    const float* pw = world + 3 * i;
    if (near_cnt[i] > 0 && ekf_inited) {
      const float* np = near_pts + (size_t)i * 20;
      float mid[3];
      for (int a = 0; a < 3; ++a) mid[a] = (float)(std::floor((double)pw[a] / fs) * fs + 0.5 * fs);
      const float dist = dist2_f(pw[0], pw[1], pw[2], mid[0], mid[1], mid[2]);
      if (std::fabs((double)(np[0] - mid[0])) > 0.5 * fs && std::fabs((double)(np[1] - mid[1])) > 0.5 * fs &&
          std::fabs((double)(np[2] - mid[2])) > 0.5 * fs) {
        class_out[i] = 2;
        continue;
      }
      bool need_add = true;
      for (int j = 0; j < 5; ++j) {
        if (near_cnt[i] < 5) break;
        if (dist2_f(np[4 * j], np[4 * j + 1], np[4 * j + 2], mid[0], mid[1], mid[2]) < dist) {
          need_add = false;
          break;
        }
      }
      class_out[i] = need_add ? 1 : 0;
    } else {
      class_out[i] = 1;
    }
  }
}

}  // namespace orc

// =====================================================================================
// extern "C" surface for ctypes (tests / bench cpu_baseline only)
// =====================================================================================
using namespace orc;

extern "C" {

int orc_abi_version() { return 1; }

void orc_so3_exp(const double w[3], double q[4]) {
  Quat r = so3_exp(w);
  q[0] = r.w;
  q[1] = r.x;
  q[2] = r.y;
  q[3] = r.z;
}
void orc_so3_log(const double q[4], double w[3]) { so3_log(Quat{q[0], q[1], q[2], q[3]}, w); }
void orc_quat_to_mat(const double q[4], double m[9]) { quat_to_mat(Quat{q[0], q[1], q[2], q[3]}, m); }
void orc_mat_to_quat(const double m[9], double q[4]) {
  Quat r = mat_to_quat(m);
  q[0] = r.w;
  q[1] = r.x;
  q[2] = r.y;
  q[3] = r.z;
}
void orc_quat_rotate(const double q[4], const double v[3], double o[3]) { quat_rotate(Quat{q[0], q[1], q[2], q[3]}, v, o); }
void orc_boxplus(const double* x26, const double* f24, double* out26) {
  State r = boxplus(*reinterpret_cast<const State*>(x26), f24);
  std::memcpy(out26, &r, sizeof(State));
}
void orc_boxminus(const double* x1, const double* x2, double* out24) {
  boxminus(*reinterpret_cast<const State*>(x1), *reinterpret_cast<const State*>(x2), out24);
}
int orc_inverse(const double* A, double* Ainv, int n) { return inverse_lu(A, Ainv, n) ? 0 : -1; }
int orc_qr_solve_5x3(const float* pts15, float* x3) { return qr_solve_5x3(pts15, x3) ? 1 : 0; }
int orc_esti_plane(const float* pts15, float thr, float* pabcd) { return esti_plane(pabcd, pts15, thr) ? 1 : 0; }
// batch: pts m x 15 -> pabcd m x 4, ok m
void orc_esti_plane_batch(const float* pts, int64_t m, float thr, float* pabcd, uint8_t* ok) {
  for (int64_t i = 0; i < m; ++i) ok[i] = esti_plane(pabcd + 4 * i, pts + 15 * i, thr) ? 1 : 0;
}

void orc_predict(double* x26, double* P576, double dt, const double* Q144, const double* acc, const double* gyro) {
  predict(*reinterpret_cast<State*>(x26), P576, dt, Q144, acc, gyro);
}
int orc_imu_forward(const double* imu7, int n_imu, double pcl_beg, double pcl_end, double* x26, double* P576,
                    double* carry27, double* poses22, int cap) {
  return imu_forward(reinterpret_cast<const ImuSample*>(imu7), n_imu, pcl_beg, pcl_end, *reinterpret_cast<State*>(x26),
                     P576, *reinterpret_cast<ImuCarry*>(carry27), reinterpret_cast<Pose6D*>(poses22), cap);
}

// IMU_init (IMU_Processing.hpp:180-244).  acc/gyr running stats live in stats[13] =
// {mean_acc[3], mean_gyr[3], cov_acc[3], cov_gyr[3], N}; first_frame resets them as :187-198.
void orc_imu_init(const double* imu7, int n_imu, int first_frame, double* stats13, double* x26, double* P576,
                  const double* tli3, const double* rli_mat9) {
  const ImuSample* s = reinterpret_cast<const ImuSample*>(imu7);
  double* mean_acc = stats13;
  double* mean_gyr = stats13 + 3;
  double* cov_acc = stats13 + 6;
  double* cov_gyr = stats13 + 9;
  double& N = stats13[12];
  if (first_frame) {
    N = 1;
    for (int i = 0; i < 3; ++i) {
      mean_acc[i] = s[0].acc[i];
      mean_gyr[i] = s[0].gyr[i];
    }
  }
  for (int k = 0; k < n_imu; ++k) {
    for (int i = 0; i < 3; ++i) {
      const double ca = s[k].acc[i], cg = s[k].gyr[i];
      mean_acc[i] += (ca - mean_acc[i]) / N;
      mean_gyr[i] += (cg - mean_gyr[i]) / N;
      cov_acc[i] = cov_acc[i] * (N - 1.0) / N + (ca - mean_acc[i]) * (ca - mean_acc[i]) / N;
      cov_gyr[i] = cov_gyr[i] * (N - 1.0) / N + (cg - mean_gyr[i]) * (cg - mean_gyr[i]) / N / N * (N - 1);
    }
    N += 1;
  }
  State& x = *reinterpret_cast<State*>(x26);
  const double n = norm3(mean_acc);
  for (int i = 0; i < 3; ++i) {
    x.grav[i] = -mean_acc[i] / n * 9.81;
    x.bg[i] = mean_gyr[i];
    x.tli[i] = tli3[i];
  }
  x.rli = mat_to_quat(rli_mat9);
  for (int i = 0; i < 576; ++i) P576[i] = 0.0;
  for (int i = 0; i < 24; ++i) P576[i * 24 + i] = 1.0;
  for (int i = 6; i < 12; ++i) P576[i * 24 + i] = 0.00001;
  for (int i = 15; i < 18; ++i) P576[i * 24 + i] = 0.0001;
  for (int i = 18; i < 21; ++i) P576[i * 24 + i] = 0.001;
  for (int i = 21; i < 24; ++i) P576[i * 24 + i] = 0.00001;
}

// std::sort by curvature (IMU_Processing.hpp:269) done here as a stable sort; order_out[j] = input index
// of sorted position j.  pts4 (N x 4: x,y,z,curvature_ms) is rewritten in sorted order, then compensated.
void orc_undistort(float* pts4, int64_t n, const double* poses22, int n_poses, const double* end26, int64_t* order_out) {
  std::vector<int64_t> ord((size_t)n);
  for (int64_t i = 0; i < n; ++i) ord[i] = i;
  std::stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return pts4[4 * a + 3] < pts4[4 * b + 3]; });
  std::vector<float> tmp(pts4, pts4 + 4 * n);
  for (int64_t j = 0; j < n; ++j) std::memcpy(pts4 + 4 * j, &tmp[4 * ord[j]], 16);
  if (order_out) std::memcpy(order_out, ord.data(), sizeof(int64_t) * n);
  undistort_backward(pts4, n, reinterpret_cast<const Pose6D*>(poses22), n_poses, *reinterpret_cast<const State*>(end26));
}

int64_t orc_voxel_grid(const float* in5, int64_t n, float leaf, float* out5, int32_t* keys_out, int32_t* point_keys) {
  return voxel_grid(in5, n, leaf, out5, keys_out, point_keys);
}

// ---- map ----
void* orc_map_create(float cell) { return new Map(cell); }
void orc_map_destroy(void* m) { delete static_cast<Map*>(m); }
void orc_map_build(void* m, const float* xyz, int64_t n) { static_cast<Map*>(m)->build(xyz, n); }
int orc_map_add(void* m, const float* xyz, int64_t n, int downsample_on, float ds) {
  return static_cast<Map*>(m)->add_points(xyz, n, downsample_on != 0, ds);
}
int orc_map_delete_boxes(void* m, const float* boxes6, int nb) { return static_cast<Map*>(m)->delete_boxes(boxes6, nb); }
int64_t orc_map_size(void* m) { return static_cast<Map*>(m)->n_live; }
int64_t orc_map_dump(void* m, float* xyz, int32_t* ids, int64_t cap) { return static_cast<Map*>(m)->dump(xyz, ids, cap); }
void orc_map_knn(void* m, const float* q, int64_t nq, int k, float max_d2, int32_t* idx, float* d2, float* nbr_xyz,
                 int threads) {
  Map* mp = static_cast<Map*>(m);
  if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t i = 0; i < nq; ++i) {
    MapPoint tmp[16];
    float dd[16];
    const int c = mp->knn(q + 3 * i, k, max_d2, tmp, dd);
    for (int j = 0; j < k; ++j) {
      const bool ok = j < c;
      idx[i * k + j] = ok ? tmp[j].id : -1;
      d2[i * k + j] = ok ? dd[j] : INFINITY;
      if (nbr_xyz) {
        nbr_xyz[(i * k + j) * 3] = ok ? tmp[j].x : 0.f;
        nbr_xyz[(i * k + j) * 3 + 1] = ok ? tmp[j].y : 0.f;
        nbr_xyz[(i * k + j) * 3 + 2] = ok ? tmp[j].z : 0.f;
      }
    }
  }
}
void* orc_map_knn_callback() { return (void*)&map_knn_cb; }

// ---- scan context for h_share_model / update ----
void* orc_scan_create(const float* body_xyz, int64_t m) {
  Scan* s = new Scan();
  s->m = m;
  float* b = new float[(size_t)3 * m + 1];
  std::memcpy(b, body_xyz, sizeof(float) * 3 * m);
  s->body = b;
  s->world.assign((size_t)3 * m, 0.f);
  s->near_pts.assign((size_t)20 * m, 0.f);
  s->near_d2.assign((size_t)5 * m, INFINITY);
  s->near_cnt.assign((size_t)m, 0);
  s->selected.assign((size_t)m, 0);
  if (m > 0) s->selected[0] = 1;  // `bool point_selected_surf[100000] = {1}` (esekfom.hpp:29)
  s->normvec.assign((size_t)4 * m, 0.f);
  return s;
}
void orc_scan_destroy(void* sp) {
  Scan* s = static_cast<Scan*>(sp);
  delete[] s->body;
  delete s;
}
// One h_share_model call.  Returns V (0 => valid=false).
int64_t orc_h_share_model(void* sp, const double* x26, int converge, int extrinsic_est, void* knn_cb, void* knn_ctx,
                          int threads) {
  Scan* s = static_cast<Scan*>(sp);
  Params prm;
  prm.threads = threads < 1 ? 1 : threads;
  const bool ok = h_share_model(*reinterpret_cast<const State*>(x26), *s, converge != 0, extrinsic_est != 0,
                                (knn_fn)knn_cb, knn_ctx, prm);
  return ok ? (int64_t)s->valid_index.size() : 0;
}
// Accessors for per-point / per-row results of the last h_share_model call.
void orc_scan_get(void* sp, float* world3, float* near_pts20, float* near_d2_5, int32_t* near_cnt, uint8_t* selected,
                  float* normvec4) {
  Scan* s = static_cast<Scan*>(sp);
  const size_t m = (size_t)s->m;
  if (world3) std::memcpy(world3, s->world.data(), 12 * m);
  if (near_pts20) std::memcpy(near_pts20, s->near_pts.data(), 80 * m);
  if (near_d2_5) std::memcpy(near_d2_5, s->near_d2.data(), 20 * m);
  if (near_cnt) std::memcpy(near_cnt, s->near_cnt.data(), 4 * m);
  if (selected) std::memcpy(selected, s->selected.data(), m);
  if (normvec4) std::memcpy(normvec4, s->normvec.data(), 16 * m);
}
void orc_scan_get_rows(void* sp, double* h_x, double* h, int32_t* valid_index) {
  Scan* s = static_cast<Scan*>(sp);
  const size_t v = s->valid_index.size();
  if (h_x) std::memcpy(h_x, s->h_x.data(), 96 * v);
  if (h) std::memcpy(h, s->h.data(), 8 * v);
  if (valid_index) std::memcpy(valid_index, s->valid_index.data(), 4 * v);
}
// Full update.  trace: array of PassTrace (4 int32 + 90 + 24 doubles = 928 bytes each).  Returns #passes.
int orc_update(void* sp, double* x26, double* P576, double R, int max_iter, int extrinsic_est, void* knn_cb,
               void* knn_ctx, int threads, void* trace, int trace_cap, int* n_valid_last) {
  Scan* s = static_cast<Scan*>(sp);
  Params prm;
  prm.threads = threads < 1 ? 1 : threads;
  return update_iterated(*reinterpret_cast<State*>(x26), P576, R, *s, (knn_fn)knn_cb, knn_ctx, max_iter,
                         extrinsic_est != 0, prm, static_cast<PassTrace*>(trace), trace_cap, n_valid_last);
}
int orc_sizeof_pass_trace() { return (int)sizeof(PassTrace); }

void orc_body_to_world(const double* x26, const float* body, int64_t m, float* world) {
  body_to_world(*reinterpret_cast<const State*>(x26), body, m, world);
}
void orc_map_incremental_classify(const float* world, int64_t m, const float* near_pts20, const int32_t* near_cnt,
                                  int ekf_inited, float filter_size_map, uint8_t* cls) {
  map_incremental_classify(world, m, near_pts20, near_cnt, ekf_inited != 0, filter_size_map, cls);
}

}  // extern "C"
