// fp64 building blocks of the PnP stage, written once for device and host (the host build exists only so the
// CPU test-suite can check this math against the oracle without a GPU: tests/harness/pnp_host.cu).
//   * EPnP (Lepetit, Moreno-Noguer, Fua 2009) for pose hypotheses from >= 4 correspondences
//   * cyclic-Jacobi symmetric eigen-solver (3x3 control-point PCA, 12x12 M^T M, 4x4 Horn quaternion matrix)
//   * Levenberg-Marquardt refit on (so(3) x R^3), lane-parallel over points with an all-reduce policy
// Replaces cv2.solvePnPRansac / cv2.solvePnP at 3_6Dpose_estimator/utils/utils.py:25-36.
#pragma once
#include <math.h>
#include <stdint.h>

// loop unrolling matters on the device only (arrays with compile-time indices live in registers); the host pass of nvcc and
// plain g++ do not know the pragma
#if defined(__CUDA_ARCH__)
#define BP_UNROLL _Pragma("unroll")
#define BP_UNROLL_SMALL(N) _Pragma("unroll")
#else
#define BP_UNROLL
#define BP_UNROLL_SMALL(N)
#endif
#if defined(__CUDACC__)
#define BP_HD __host__ __device__ __forceinline__
// NOTE: these were __noinline__ at first; cicc 12.9 -O3 then merges the stack slots of the caller's local arrays that
// are passed by pointer (jacobi_eig's A and V came back aliased: eigenvalues all 1.0 on the device, correct with
// -G / -Xcicc -O1 and on the host).  Forced inlining avoids the miscompile; each function has one call site.
#define BP_HD_NOINLINE __host__ __device__ __forceinline__
#else
#define BP_HD inline
#define BP_HD_NOINLINE
#endif

namespace bp {
namespace pnp {

// ------------------------------------------------------------------ sampling (spec shared with oracle/pnp.py)
BP_HD uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
BP_HD uint32_t xorshift32(uint32_t s) {
  s ^= s << 13;
  s ^= s >> 17;
  s ^= s << 5;
  return s;
}
// first m entries of a partial Fisher-Yates shuffle of `pool` (caller fills pool[0..n) with candidate ids)
BP_HD void sample_subset(int* pool, int n, int h, uint32_t seed, int m) {
  uint32_t s = hash32(seed * 0x9E3779B9u + (uint32_t)h * 0x85EBCA6Bu + 1u);
  if (s == 0) s = 1;
  for (int i = 0; i < m; ++i) {
    s = xorshift32(s);
    const int j = i + (int)(s % (uint32_t)(n - i));
    const int tmp = pool[i];
    pool[i] = pool[j];
    pool[j] = tmp;
  }
}

// ------------------------------------------------------------------ small dense helpers
// cyclic Jacobi on a symmetric N x N matrix (row-major, destroyed); V gets eigenvectors in columns
template <int N>
BP_HD_NOINLINE void jacobi_eig(double* A, double* V, double* w) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i * N + j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int p = 0; p < N; ++p) {
      diag += A[p * N + p] * A[p * N + p];
      for (int q = p + 1; q < N; ++q) off += A[p * N + q] * A[p * N + q];
    }
    if (off < 1e-300) break;
    // quadratic convergence: once the off-diagonal mass is below 1e-22 of the diagonal one more sweep reaches the
    // rounding floor
    const bool last = off <= 1e-22 * diag;
    // N <= 4 (control-point PCA, Horn's quaternion matrix): fully unrolled, matrices in registers
BP_UNROLL_SMALL(N)
    for (int p = 0; p < N - 1; ++p) {
BP_UNROLL_SMALL(N)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
      }
    }
    if (last) break;
  }
  for (int i = 0; i < N; ++i) w[i] = A[i * N + i];
}

// Symmetric eigen-decomposition by Householder tridiagonalisation + implicit-shift QL (the EISPACK tred2 / tql2 scheme):
// a (row-major, symmetric) is overwritten by the eigenvectors (column i <-> d[i], unsorted).  ~5x fewer flops than
// cyclic Jacobi at N = 12 (one pass instead of ~10 sweeps), which is what lets the hypothesis kernel give every EPnP
// hypothesis ONE thread and no shared memory.  Deflation uses an absolute tolerance eps * |A| (as LAPACK's eigh, which
// the oracle calls): M^T M of a minimal sample is rank deficient and its zero eigenvalues never pass a relative test.
template <int N>
BP_HD_NOINLINE void sym_eig_ql(double* a, double* d) {
  double e[N];
  for (int i = N - 1; i > 0; --i) {
    const int l = i - 1;
    double h = 0.0, scale = 0.0;
    if (l > 0) {
      for (int k = 0; k <= l; ++k) scale += fabs(a[i * N + k]);
      if (scale == 0.0) {
        e[i] = a[i * N + l];
      } else {
        const double inv = 1.0 / scale;
        for (int k = 0; k <= l; ++k) {
          a[i * N + k] *= inv;
          h += a[i * N + k] * a[i * N + k];
        }
        double f = a[i * N + l];
        double g = f >= 0.0 ? -sqrt(h) : sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        a[i * N + l] = f - g;
        f = 0.0;
        const double ih = 1.0 / h;
        for (int j = 0; j <= l; ++j) {
          a[j * N + i] = a[i * N + j] * ih;
          g = 0.0;
          for (int k = 0; k <= j; ++k) g += a[j * N + k] * a[i * N + k];
          for (int k = j + 1; k <= l; ++k) g += a[k * N + j] * a[i * N + k];
          e[j] = g * ih;
          f += e[j] * a[i * N + j];
        }
        const double hh = f / (h + h);
        for (int j = 0; j <= l; ++j) {
          f = a[i * N + j];
          e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; ++k) a[j * N + k] -= f * e[k] + g * a[i * N + k];
        }
      }
    } else {
      e[i] = a[i * N + l];
    }
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (int i = 0; i < N; ++i) {
    const int l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; ++j) {
        double g = 0.0;
        for (int k = 0; k <= l; ++k) g += a[i * N + k] * a[k * N + j];
        for (int k = 0; k <= l; ++k) a[k * N + j] -= g * a[k * N + i];
      }
    }
    d[i] = a[i * N + i];
    a[i * N + i] = 1.0;
    for (int j = 0; j <= l; ++j) a[j * N + i] = a[i * N + j] = 0.0;
  }
  for (int i = 1; i < N; ++i) e[i - 1] = e[i];
  e[N - 1] = 0.0;
  double anorm = 0.0;
  for (int i = 0; i < N; ++i) anorm = fmax(anorm, fabs(d[i]) + fabs(e[i]));
  const double tol = 2.220446049250313e-16 * anorm;
  for (int l = 0; l < N; ++l) {
    for (int iter = 0; iter < 60; ++iter) {
      int m = l;
      for (; m < N - 1; ++m)
        if (fabs(e[m]) <= tol) break;
      if (m == l) break;
      double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
      double r = sqrt(g * g + 1.0);
      g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? r : -r));
      double s = 1.0, c = 1.0, p = 0.0;
      int i = m - 1;
      for (; i >= l; --i) {
        double f = s * e[i];
        const double b = c * e[i];
        r = sqrt(f * f + g * g);
        e[i + 1] = r;
        if (r == 0.0) {
          d[i + 1] -= p;
          e[m] = 0.0;
          break;
        }
        s = f / r;
        c = g / r;
        g = d[i + 1] - p;
        r = (d[i] - g) * s + 2.0 * c * b;
        p = s * r;
        d[i + 1] = g + p;
        g = c * r - b;
        for (int k = 0; k < N; ++k) {
          f = a[k * N + i + 1];
          a[k * N + i + 1] = s * a[k * N + i] + c * f;
          a[k * N + i] = c * a[k * N + i] - s * f;
        }
      }
      if (r == 0.0 && i >= l) continue;
      d[l] -= p;
      e[l] = g;
      e[m] = 0.0;
    }
  }
}

// solve the N x N system A x = b in place by Gaussian elimination with partial pivoting; false if singular
template <int N>
BP_HD bool solve_linear(double* A, double* b) {
  for (int c = 0; c < N; ++c) {
    int piv = c;
    double best = fabs(A[c * N + c]);
    for (int r = c + 1; r < N; ++r)
      if (fabs(A[r * N + c]) > best) {
        best = fabs(A[r * N + c]);
        piv = r;
      }
    if (!(best > 1e-300)) return false;
    if (piv != c) {
      for (int k = 0; k < N; ++k) {
        const double t = A[c * N + k];
        A[c * N + k] = A[piv * N + k];
        A[piv * N + k] = t;
      }
      const double t = b[c];
      b[c] = b[piv];
      b[piv] = t;
    }
    const double inv = 1.0 / A[c * N + c];
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r * N + c] * inv;
      if (f == 0.0) continue;
      for (int k = c; k < N; ++k) A[r * N + k] -= f * A[c * N + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = N - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < N; ++k) s -= A[r * N + k] * b[k];
    b[r] = s / A[r * N + r];
  }
  return true;
}

// solve the symmetric positive definite N x N system A x = b by an LDL^T factorisation, every loop unrolled at compile time
// so that A lives in registers (solve_linear's partial pivoting indexes rows dynamically, which puts its matrix in local
// memory: on the device that was most of a Levenberg-Marquardt iteration).  A: full symmetric matrix (only the lower
// triangle is read), destroyed.  false if a pivot is not positive (A not positive definite).
template <int N>
BP_HD bool solve_spd(double* A, double* b) {
  double d[N];
  bool ok = true;  // (no early exit: it would keep the compiler from unrolling the column loop)
BP_UNROLL
  for (int j = 0; j < N; ++j) {
    double dj = A[j * N + j];
BP_UNROLL
    for (int k = 0; k < j; ++k) dj -= A[j * N + k] * A[j * N + k] * d[k];
    ok = ok && (dj > 0.0);
    d[j] = dj;
    const double inv = 1.0 / dj;
BP_UNROLL
    for (int i = j + 1; i < N; ++i) {
      double lij = A[i * N + j];
BP_UNROLL
      for (int k = 0; k < j; ++k) lij -= A[i * N + k] * A[j * N + k] * d[k];
      A[i * N + j] = lij * inv;
    }
  }
BP_UNROLL
  for (int i = 0; i < N; ++i) {  // L z = b
BP_UNROLL
    for (int k = 0; k < i; ++k) b[i] -= A[i * N + k] * b[k];
  }
BP_UNROLL
  for (int i = 0; i < N; ++i) b[i] /= d[i];  // D y = z
BP_UNROLL
  for (int i = N - 1; i >= 0; --i) {  // L^T x = y
BP_UNROLL
    for (int k = i + 1; k < N; ++k) b[i] -= A[k * N + i] * b[k];
  }
  return ok;
}

// least squares  min |A x - b|  for a 6 x NC system via (ridge-stabilised) normal equations
template <int NC>
BP_HD bool lstsq6(const double* A /*6 x NC*/, const double* b /*6*/, double* x /*NC*/) {
  double AtA[NC * NC], Atb[NC];
  double tr = 0.0;
BP_UNROLL
  for (int i = 0; i < NC; ++i) {
BP_UNROLL
    for (int j = 0; j < NC; ++j) {
      double s = 0.0;
BP_UNROLL
      for (int r = 0; r < 6; ++r) s += A[r * NC + i] * A[r * NC + j];
      AtA[i * NC + j] = s;
    }
    double s = 0.0;
BP_UNROLL
    for (int r = 0; r < 6; ++r) s += A[r * NC + i] * b[r];
    Atb[i] = s;
    tr += AtA[i * NC + i];
  }
BP_UNROLL
  for (int i = 0; i < NC; ++i) AtA[i * NC + i] += 1e-14 * tr + 1e-300;
  if (!solve_spd<NC>(AtA, Atb)) return false;  // A^T A + ridge: symmetric positive definite; unrolled, in registers
BP_UNROLL
  for (int i = 0; i < NC; ++i) x[i] = Atb[i];
  return true;
}

BP_HD void project(const double* R, const double* t, const double* X, double fx, double fy, double cx, double cy,
                   double* u, double* v, double* z) {
  const double xc = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  const double yc = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  const double zc = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
  const double iz = 1.0 / zc;
  *u = fx * xc * iz + cx;
  *v = fy * yc * iz + cy;
  *z = zc;
}

// Horn's closed-form absolute orientation: proper rotation R (row-major) with  dst ~ R src, from the
// cross-covariance S[a][b] = sum src_a * dst_b
BP_HD_NOINLINE void horn_rotation(const double S[9], double* R) {
  const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5], Szx = S[6], Szy = S[7], Szz = S[8];
  double N4[16] = {Sxx + Syy + Szz, Syz - Szy,       Szx - Sxz,        Sxy - Syx,
                   Syz - Szy,       Sxx - Syy - Szz, Sxy + Syx,        Szx + Sxz,
                   Szx - Sxz,       Sxy + Syx,       -Sxx + Syy - Szz, Syz + Szy,
                   Sxy - Syx,       Szx + Sxz,       Syz + Szy,        -Sxx - Syy + Szz};
  double V[16], w[4];
  jacobi_eig<4>(N4, V, w);
  int m = 0;
  for (int i = 1; i < 4; ++i)
    if (w[i] > w[m]) m = i;
  double q0 = V[0 * 4 + m], q1 = V[1 * 4 + m], q2 = V[2 * 4 + m], q3 = V[3 * 4 + m];
  const double nn = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  q0 /= nn; q1 /= nn; q2 /= nn; q3 /= nn;
  R[0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
  R[1] = 2 * (q1 * q2 - q0 * q3);
  R[2] = 2 * (q1 * q3 + q0 * q2);
  R[3] = 2 * (q2 * q1 + q0 * q3);
  R[4] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3;
  R[5] = 2 * (q2 * q3 - q0 * q1);
  R[6] = 2 * (q3 * q1 - q0 * q2);
  R[7] = 2 * (q3 * q2 + q0 * q1);
  R[8] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
}

// ------------------------------------------------------------------ EPnP
// Eigen-solver policy for the 12 x 12 M^T M: `solve` destroys A, fills w with the eigenvalues and returns the
// eigenvectors (columns, row-major 12 x 12).  SerialEig12 is the single-thread cyclic Jacobi (host build, tests);
// the device kernel passes a policy that runs the rotations of one Jacobi round on 16 cooperating lanes (pnp.cu).
struct SerialEig12 {
  double V[144];
  BP_HD const double* solve(double* A, double* w) {
    jacobi_eig<12>(A, V, w);
    return V;
  }
};

// single-thread policy on the QL solver above: the device throughput path (one thread per hypothesis) and the host build
struct QlEig12 {
  BP_HD const double* solve(double* A, double* w) {
    sym_eig_ql<12>(A, w);
    return A;
  }
};

// pw[n][3] world points, uv[n][2] pixels, ids[0..n) selects the correspondences.  Returns false on a
// degenerate configuration.  Scratch lives on the caller's stack/local memory (~3.5 KB).
template <class Eig>
BP_HD_NOINLINE bool epnp(Eig& eig, const double* pw_all, const double* uv_all, const int* ids, int n, double fx, double fy,
                         double cx, double cy, double* R_out, double* t_out
#ifdef BP_PNP_DEBUG
                         , double* dbg
#endif
) {
  // 1. control points: centroid + principal axes
  double c0[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) c0[k] += pw_all[ids[i] * 3 + k];
  for (int k = 0; k < 3; ++k) c0[k] /= n;
  double C3[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    double d[3];
    for (int k = 0; k < 3; ++k) d[k] = pw_all[ids[i] * 3 + k] - c0[k];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) C3[a * 3 + b] += d[a] * d[b];
  }
  double V3[9], w3[3];
  jacobi_eig<3>(C3, V3, w3);
  double cws[4][3];
  for (int k = 0; k < 3; ++k) cws[0][k] = c0[k];
  for (int i = 0; i < 3; ++i) {
    if (!(w3[i] > 0.0)) return false;
    const double s = sqrt(w3[i] / n);
    for (int k = 0; k < 3; ++k) cws[i + 1][k] = c0[k] + s * V3[k * 3 + i];
  }
  // 2. barycentric coordinates: alpha_{1..3} = CC^-1 (pw - c0)
  double CC[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) CC[i * 3 + j] = cws[j + 1][i] - cws[0][i];
  const double det = CC[0] * (CC[4] * CC[8] - CC[5] * CC[7]) - CC[1] * (CC[3] * CC[8] - CC[5] * CC[6]) +
                     CC[2] * (CC[3] * CC[7] - CC[4] * CC[6]);
  if (!(fabs(det) > 1e-300)) return false;
  const double id = 1.0 / det;
  double Ci[9];
  Ci[0] = (CC[4] * CC[8] - CC[5] * CC[7]) * id;
  Ci[1] = (CC[2] * CC[7] - CC[1] * CC[8]) * id;
  Ci[2] = (CC[1] * CC[5] - CC[2] * CC[4]) * id;
  Ci[3] = (CC[5] * CC[6] - CC[3] * CC[8]) * id;
  Ci[4] = (CC[0] * CC[8] - CC[2] * CC[6]) * id;
  Ci[5] = (CC[2] * CC[3] - CC[0] * CC[5]) * id;
  Ci[6] = (CC[3] * CC[7] - CC[4] * CC[6]) * id;
  Ci[7] = (CC[1] * CC[6] - CC[0] * CC[7]) * id;
  Ci[8] = (CC[0] * CC[4] - CC[1] * CC[3]) * id;
  auto alphas_of = [&](int pid, double* a) {
    double d[3];
    for (int k = 0; k < 3; ++k) d[k] = pw_all[pid * 3 + k] - c0[k];
    for (int j = 0; j < 3; ++j) a[j + 1] = Ci[j * 3] * d[0] + Ci[j * 3 + 1] * d[1] + Ci[j * 3 + 2] * d[2];
    a[0] = 1.0 - a[1] - a[2] - a[3];
  };
  // 3. M^T M (12 x 12)
  double MtM[144];
  for (int i = 0; i < 144; ++i) MtM[i] = 0.0;
  for (int i = 0; i < n; ++i) {
    double a[4];
    alphas_of(ids[i], a);
    const double u = uv_all[ids[i] * 2], v = uv_all[ids[i] * 2 + 1];
    double m1[12], m2[12];
    for (int j = 0; j < 4; ++j) {
      m1[3 * j] = a[j] * fx; m1[3 * j + 1] = 0.0;       m1[3 * j + 2] = a[j] * (cx - u);
      m2[3 * j] = 0.0;       m2[3 * j + 1] = a[j] * fy; m2[3 * j + 2] = a[j] * (cy - v);
    }
    for (int r = 0; r < 12; ++r)
      for (int c = r; c < 12; ++c) MtM[r * 12 + c] += m1[r] * m1[c] + m2[r] * m2[c];
  }
  for (int r = 0; r < 12; ++r)
    for (int c = 0; c < r; ++c) MtM[r * 12 + c] = MtM[c * 12 + r];
  double w12[12];
  const double* V12 = eig.solve(MtM, w12);
  // the four eigenvectors of smallest eigenvalue, v[0] = smallest
  int order[4];
  {
    bool used[12];
    for (int i = 0; i < 12; ++i) used[i] = false;
    for (int k = 0; k < 4; ++k) {
      int m = -1;
      for (int i = 0; i < 12; ++i)
        if (!used[i] && (m < 0 || w12[i] < w12[m])) m = i;
      used[m] = true;
      order[k] = m;
    }
  }
#ifdef BP_PNP_DEBUG
  for (int i = 0; i < 12; ++i) dbg[i] = w12[i];
  for (int i = 0; i < 4; ++i) dbg[12 + i] = order[i];
  for (int i = 0; i < 3; ++i) dbg[16 + i] = w3[i];
#endif
  double v[4][12];
  for (int k = 0; k < 4; ++k)
    for (int i = 0; i < 12; ++i) v[k][i] = V12[i * 12 + order[k]];
  // 4. L (6 x 10) and rho (6)
  const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
  double L[60], rho[6];
  for (int p = 0; p < 6; ++p) {
    double dv[4][3];
    for (int k = 0; k < 4; ++k)
      for (int c = 0; c < 3; ++c) dv[k][c] = v[k][3 * pa[p] + c] - v[k][3 * pb[p] + c];
    auto dot = [&](int a, int b) { return dv[a][0] * dv[b][0] + dv[a][1] * dv[b][1] + dv[a][2] * dv[b][2]; };
    double* Lr = L + p * 10;
    Lr[0] = dot(0, 0); Lr[1] = 2 * dot(0, 1); Lr[2] = dot(1, 1); Lr[3] = 2 * dot(0, 2); Lr[4] = 2 * dot(1, 2);
    Lr[5] = dot(2, 2); Lr[6] = 2 * dot(0, 3); Lr[7] = 2 * dot(1, 3); Lr[8] = 2 * dot(2, 3); Lr[9] = dot(3, 3);
    double s = 0.0;
    for (int c = 0; c < 3; ++c) s += (cws[pa[p]][c] - cws[pb[p]][c]) * (cws[pa[p]][c] - cws[pb[p]][c]);
    rho[p] = s;
  }
  // 5. three beta initialisations, Gauss-Newton, pick the smallest reprojection error
  double best_err = INFINITY;
  bool have = false;
  for (int cand = 0; cand < 3; ++cand) {
    double be[4] = {0, 0, 0, 0};
    if (cand == 0) {
      double A[24], x[4];
      for (int r = 0; r < 6; ++r) { A[r * 4] = L[r * 10]; A[r * 4 + 1] = L[r * 10 + 1]; A[r * 4 + 2] = L[r * 10 + 3]; A[r * 4 + 3] = L[r * 10 + 6]; }
      if (!lstsq6<4>(A, rho, x)) continue;
      if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = -x[1] / be[0]; be[2] = -x[2] / be[0]; be[3] = -x[3] / be[0]; }
      else { be[0] = sqrt(x[0]); be[1] = x[1] / be[0]; be[2] = x[2] / be[0]; be[3] = x[3] / be[0]; }
    } else if (cand == 1) {
      double A[18], x[3];
      for (int r = 0; r < 6; ++r) { A[r * 3] = L[r * 10]; A[r * 3 + 1] = L[r * 10 + 1]; A[r * 3 + 2] = L[r * 10 + 2]; }
      if (!lstsq6<3>(A, rho, x)) continue;
      if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = x[2] < 0 ? sqrt(-x[2]) : 0.0; }
      else { be[0] = sqrt(x[0]); be[1] = x[2] > 0 ? sqrt(x[2]) : 0.0; }
      if (x[1] < 0) be[0] = -be[0];
    } else {
      double A[30], x[5];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 5; ++c) A[r * 5 + c] = L[r * 10 + c];
      if (!lstsq6<5>(A, rho, x)) continue;
      if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = x[2] < 0 ? sqrt(-x[2]) : 0.0; }
      else { be[0] = sqrt(x[0]); be[1] = x[2] > 0 ? sqrt(x[2]) : 0.0; }
      if (x[1] < 0) be[0] = -be[0];
      be[2] = be[0] != 0.0 ? x[3] / be[0] : 0.0;
    }
#ifdef BP_PNP_DEBUG
    for (int i = 0; i < 4; ++i) dbg[40 + cand * 4 + i] = be[i];
#endif
    // Gauss-Newton on the 4 betas (5 iterations)
    for (int it = 0; it < 5; ++it) {
      double A[24], r6[6], x[4];
      for (int r = 0; r < 6; ++r) {
        const double* Lr = L + r * 10;
        A[r * 4 + 0] = 2 * Lr[0] * be[0] + Lr[1] * be[1] + Lr[3] * be[2] + Lr[6] * be[3];
        A[r * 4 + 1] = Lr[1] * be[0] + 2 * Lr[2] * be[1] + Lr[4] * be[2] + Lr[7] * be[3];
        A[r * 4 + 2] = Lr[3] * be[0] + Lr[4] * be[1] + 2 * Lr[5] * be[2] + Lr[8] * be[3];
        A[r * 4 + 3] = Lr[6] * be[0] + Lr[7] * be[1] + Lr[8] * be[2] + 2 * Lr[9] * be[3];
        r6[r] = rho[r] - (Lr[0] * be[0] * be[0] + Lr[1] * be[0] * be[1] + Lr[2] * be[1] * be[1] + Lr[3] * be[0] * be[2] +
                          Lr[4] * be[1] * be[2] + Lr[5] * be[2] * be[2] + Lr[6] * be[0] * be[3] + Lr[7] * be[1] * be[3] +
                          Lr[8] * be[2] * be[3] + Lr[9] * be[3] * be[3]);
      }
      if (!lstsq6<4>(A, r6, x)) break;
      for (int k = 0; k < 4; ++k) be[k] += x[k];
    }
    // camera-frame control points, per-point camera coordinates, sign, absolute orientation
    double ccs[4][3];
    for (int j = 0; j < 4; ++j)
      for (int c = 0; c < 3; ++c) ccs[j][c] = be[0] * v[0][3 * j + c] + be[1] * v[1][3 * j + c] + be[2] * v[2][3 * j + c] + be[3] * v[3][3 * j + c];
    double a0[4];
    alphas_of(ids[0], a0);
    const double z0 = a0[0] * ccs[0][2] + a0[1] * ccs[1][2] + a0[2] * ccs[2][2] + a0[3] * ccs[3][2];
    if (z0 < 0)
      for (int j = 0; j < 4; ++j)
        for (int c = 0; c < 3; ++c) ccs[j][c] = -ccs[j][c];
    double pc0[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
      double a[4];
      alphas_of(ids[i], a);
      for (int c = 0; c < 3; ++c) pc0[c] += a[0] * ccs[0][c] + a[1] * ccs[1][c] + a[2] * ccs[2][c] + a[3] * ccs[3][c];
    }
    for (int c = 0; c < 3; ++c) pc0[c] /= n;
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // S[a][b] = sum (pw - pw0)_a (pc - pc0)_b
    for (int i = 0; i < n; ++i) {
      double a[4], pc[3];
      alphas_of(ids[i], a);
      for (int c = 0; c < 3; ++c) pc[c] = a[0] * ccs[0][c] + a[1] * ccs[1][c] + a[2] * ccs[2][c] + a[3] * ccs[3][c] - pc0[c];
      for (int aa = 0; aa < 3; ++aa)
        for (int bb = 0; bb < 3; ++bb) S[aa * 3 + bb] += (pw_all[ids[i] * 3 + aa] - c0[aa]) * pc[bb];
    }
    double Rc[9], tc[3];
    horn_rotation(S, Rc);
    for (int r = 0; r < 3; ++r) tc[r] = pc0[r] - (Rc[r * 3] * c0[0] + Rc[r * 3 + 1] * c0[1] + Rc[r * 3 + 2] * c0[2]);
    double err = 0.0;
    for (int i = 0; i < n; ++i) {
      double u, vv, z;
      project(Rc, tc, pw_all + ids[i] * 3, fx, fy, cx, cy, &u, &vv, &z);
      const double du = u - uv_all[ids[i] * 2], dv2 = vv - uv_all[ids[i] * 2 + 1];
      err += sqrt(du * du + dv2 * dv2);
    }
#ifdef BP_PNP_DEBUG
    dbg[20 + cand * 5 + 4] = err;
    for (int i = 0; i < 4; ++i) dbg[20 + cand * 5 + i] = be[i];
#endif
    if (err == err && err < best_err) {  // finite and better
      best_err = err;
      have = true;
      for (int k = 0; k < 9; ++k) R_out[k] = Rc[k];
      for (int k = 0; k < 3; ++k) t_out[k] = tc[k];
    }
  }
  return have;
}

#ifndef BP_PNP_DEBUG
BP_HD bool epnp(const double* pw_all, const double* uv_all, const int* ids, int n, double fx, double fy, double cx, double cy,
                double* R_out, double* t_out) {
  QlEig12 eig;
  return epnp(eig, pw_all, uv_all, ids, n, fx, fy, cx, cy, R_out, t_out);
}
#endif

// ------------------------------------------------------------------ Levenberg-Marquardt refit
// Accumulate, over this lane's share of the masked points, J^T J (21 upper entries), J^T r (6) and the cost.
// Left-multiplicative update R <- exp(w) R: d(pc)/dw = -[R X]x, d(pc)/dt = I.
template <class Lanes>
BP_HD void lm_accumulate(const Lanes& ln, const double* R, const double* t, const double* pw, const double* uv,
                         const uint8_t* mask, int n, double fx, double fy, double cx, double cy, double* acc /*28*/) {
BP_UNROLL
  for (int i = 0; i < 28; ++i) acc[i] = 0.0;
  for (int i = ln.lane(); i < n; i += ln.count()) {
    if (!mask[i]) continue;
    const double* X = pw + 3 * i;
    const double qx = R[0] * X[0] + R[1] * X[1] + R[2] * X[2];
    const double qy = R[3] * X[0] + R[4] * X[1] + R[5] * X[2];
    const double qz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2];
    const double xc = qx + t[0], yc = qy + t[1], zc = qz + t[2];
    const double iz = 1.0 / zc;
    const double ru = fx * xc * iz + cx - uv[2 * i], rv = fy * yc * iz + cy - uv[2 * i + 1];
    const double du[3] = {fx * iz, 0.0, -fx * xc * iz * iz};
    const double dv[3] = {0.0, fy * iz, -fy * yc * iz * iz};
    double Ju[6], Jv[6];
    Ju[0] = du[2] * qy - du[1] * qz; Ju[1] = du[0] * qz - du[2] * qx; Ju[2] = du[1] * qx - du[0] * qy;
    Ju[3] = du[0]; Ju[4] = du[1]; Ju[5] = du[2];
    Jv[0] = dv[2] * qy - dv[1] * qz; Jv[1] = dv[0] * qz - dv[2] * qx; Jv[2] = dv[1] * qx - dv[0] * qy;
    Jv[3] = dv[0]; Jv[4] = dv[1]; Jv[5] = dv[2];
    // (a, b) -> packed upper-triangle index 6a - a(a-1)/2 + (b - a): compile-time after unrolling, so acc stays in registers
BP_UNROLL
    for (int a = 0; a < 6; ++a)
BP_UNROLL
      for (int b = a; b < 6; ++b) acc[6 * a - a * (a - 1) / 2 + (b - a)] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
BP_UNROLL
    for (int a = 0; a < 6; ++a) acc[21 + a] += Ju[a] * ru + Jv[a] * rv;
    acc[27] += ru * ru + rv * rv;
  }
  ln.template allreduce_n<28>(acc);
}

template <class Lanes>
BP_HD double lm_cost(const Lanes& ln, const double* R, const double* t, const double* pw, const double* uv,
                     const uint8_t* mask, int n, double fx, double fy, double cx, double cy) {
  double c = 0.0;
  for (int i = ln.lane(); i < n; i += ln.count()) {
    if (!mask[i]) continue;
    double u, v, z;
    project(R, t, pw + 3 * i, fx, fy, cx, cy, &u, &v, &z);
    const double ru = u - uv[2 * i], rv = v - uv[2 * i + 1];
    c += ru * ru + rv * rv;
  }
  ln.template allreduce_n<1>(&c);
  return c;
}

BP_HD void so3_exp_mul(const double* w, const double* R, double* Rn) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double a, b;  // exp(w) = I + a K + b K^2
  if (th < 1e-12) {
    a = 1.0;
    b = 0.0;
  } else {
    a = sin(th) / th;
    b = (1.0 - cos(th)) / th2;
  }
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double E[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double k2 = 0.0;
      for (int k = 0; k < 3; ++k) k2 += K[i * 3 + k] * K[k * 3 + j];
      E[i * 3 + j] = (i == j ? 1.0 : 0.0) + a * K[i * 3 + j] + b * k2;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rn[i * 3 + j] = E[i * 3] * R[j] + E[i * 3 + 1] * R[3 + j] + E[i * 3 + 2] * R[6 + j];
}

// Every lane ends with identical R, t (all-reduced sums are bitwise identical across lanes).
template <class Lanes>
BP_HD_NOINLINE void lm_refine(const Lanes& ln, double* R, double* t, const double* pw, const double* uv, const uint8_t* mask,
                              int n, double fx, double fy, double cx, double cy, int iters) {
  double lam = 1e-3;
  double acc[28];
  ln.accumulate(R, t, pw, uv, mask, n, fx, fy, cx, cy, acc);
  double cost = acc[27];
  for (int it = 0; it < iters; ++it) {
    bool improved = false;
    double step = 0.0, dc = 0.0;
    for (int tr = 0; tr < 10; ++tr) {
      double A[36], g[6];
BP_UNROLL
      for (int a = 0; a < 6; ++a)
BP_UNROLL
        for (int b = a; b < 6; ++b) {
          A[a * 6 + b] = acc[6 * a - a * (a - 1) / 2 + (b - a)];
          A[b * 6 + a] = A[a * 6 + b];
        }
BP_UNROLL
      for (int a = 0; a < 6; ++a) {
        A[a * 6 + a] += lam * A[a * 6 + a];
        g[a] = -acc[21 + a];
      }
      if (!solve_spd<6>(A, g)) {  // J^T J + lam diag(J^T J): symmetric positive definite unless the points are degenerate
        lam *= 10;
        continue;
      }
      double Rn[9], tn[3];
      so3_exp_mul(g, R, Rn);
      for (int a = 0; a < 3; ++a) tn[a] = t[a] + g[3 + a];
      const double cn = lm_cost(ln, Rn, tn, pw, uv, mask, n, fx, fy, cx, cy);
      if (cn == cn && cn <= cost) {
        improved = true;
        for (int a = 0; a < 6; ++a) step = fmax(step, fabs(g[a]));
        for (int a = 0; a < 9; ++a) R[a] = Rn[a];
        for (int a = 0; a < 3; ++a) t[a] = tn[a];
        dc = cost - cn;
        cost = cn;
        lam = fmax(lam * 0.1, 1e-12);
        break;
      }
      lam *= 10;
    }
    // (tight on purpose: kernel, host build and oracle must run well-posed refits to the same minimum, to ~1e-9)
    if (!improved || step < 1e-13 || dc <= 1e-16 * fmax(cost, 1e-300)) break;
    ln.accumulate(R, t, pw, uv, mask, n, fx, fy, cx, cy, acc);
  }
  // one Gram-Schmidt pass against accumulated rounding drift
  double* r0 = R; double* r1 = R + 3; double* r2 = R + 6;
  double n0 = sqrt(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]);
  for (int k = 0; k < 3; ++k) r0[k] /= n0;
  double d01 = r0[0] * r1[0] + r0[1] * r1[1] + r0[2] * r1[2];
  for (int k = 0; k < 3; ++k) r1[k] -= d01 * r0[k];
  double n1 = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
  for (int k = 0; k < 3; ++k) r1[k] /= n1;
  r2[0] = r0[1] * r1[2] - r0[2] * r1[1];
  r2[1] = r0[2] * r1[0] - r0[0] * r1[2];
  r2[2] = r0[0] * r1[1] - r0[1] * r1[0];
}

struct SingleLane {  // host / single-thread policy
  BP_HD int lane() const { return 0; }
  BP_HD int count() const { return 1; }
  BP_HD void allreduce(double*, int) const {}
  template <int N>
  BP_HD void allreduce_n(double*) const {}
  BP_HD void accumulate(const double* R, const double* t, const double* pw, const double* uv, const uint8_t* mask, int n,
                        double fx, double fy, double cx, double cy, double* acc) const {
    lm_accumulate(*this, R, t, pw, uv, mask, n, fx, fy, cx, cy, acc);
  }
};

// is point i within the reprojection threshold (and in front of the camera) under (R, t)?
BP_HD bool within_threshold(const double* R, const double* t, const double* pw, const double* uv, int i, double fx, double fy,
                            double cx, double cy, double thr2) {
  double u, v, z;
  project(R, t, pw + 3 * i, fx, fy, cx, cy, &u, &v, &z);
  const double e2 = (u - uv[2 * i]) * (u - uv[2 * i]) + (v - uv[2 * i + 1]) * (v - uv[2 * i + 1]);
  return z > 0 && e2 <= thr2;
}

// number of consensus re-estimation rounds after the winning hypothesis (classify -> LM refit -> classify ...)
#define BP_PNP_LO_ROUNDS 4
// Levenberg-Marquardt iteration cap per refit, as in OpenCV's iterative solvePnP (which solvePnPRansac refits with).
// Well-posed refits stop on the criteria above in < 10 iterations; the cap only bounds ill-posed ones (junk
// key-points), where no two implementations agree on the local solution anyway.  Same value in oracle/pnp.py.
#define BP_PNP_LM_ITERS 20

// score one hypothesis against all candidate points: consensus count and summed squared error of inliers
BP_HD void score_hypothesis(const double* R, const double* t, const double* pw, const double* uv, const uint8_t* sel,
                            int n, double fx, double fy, double cx, double cy, double thr2, int* count, double* total) {
  int c = 0;
  double tot = 0.0;
  for (int i = 0; i < n; ++i) {
    if (!sel[i]) continue;
    double u, v, z;
    project(R, t, pw + 3 * i, fx, fy, cx, cy, &u, &v, &z);
    const double e2 = (u - uv[2 * i]) * (u - uv[2 * i]) + (v - uv[2 * i + 1]) * (v - uv[2 * i + 1]);
    if (z > 0 && e2 <= thr2) {
      ++c;
      tot += e2;
    }
  }
  *count = c;
  *total = tot;
}

}  // namespace pnp
}  // namespace bp
