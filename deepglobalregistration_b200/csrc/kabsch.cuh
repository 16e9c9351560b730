// Shared by registration.cu (weighted Procrustes, ICP) and ransac.cu (4-point hypotheses).
#pragma once

// ---------------------------------------------------------------------------------------
// 3x3 SVD (one-sided Jacobi, fp64) -> proper rotation U diag(1,1,det(U)det(V)) V^T
// ---------------------------------------------------------------------------------------
__device__ inline double det3(const double m[3][3]) {
  return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
         m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
         m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

__device__ inline void kabsch_rotation(const double S[3][3], double R[3][3]) {
  double A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      A[i][j] = S[i][j];
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) {
          alpha += A[k][p] * A[k][p];
          beta += A[k][q] * A[k][q];
          gamma += A[k][p] * A[k][q];
        }
        if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          double ap = A[k][p], aq = A[k][q];
          A[k][p] = c * ap - s * aq;
          A[k][q] = s * ap + c * aq;
          double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq;
          V[k][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sig[3];
  for (int j = 0; j < 3; ++j) sig[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  int ord[3] = {0, 1, 2};   // descending singular values
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (sig[ord[b]] > sig[ord[a]]) { int tmp = ord[a]; ord[a] = ord[b]; ord[b] = tmp; }
  if (sig[ord[0]] <= 1e-300) {   // zero covariance (all points coincide): LAPACK returns U = V = I
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
    return;
  }
  double U[3][3], W[3][3];
  const double tiny = 1e-300 + 1e-14 * sig[ord[0]];
  for (int j = 0; j < 3; ++j) {
    int o = ord[j];
    for (int k = 0; k < 3; ++k) {
      W[k][j] = V[k][o];
      U[k][j] = sig[o] > tiny ? A[k][o] / sig[o] : 0.0;
    }
  }
  if (sig[ord[1]] <= tiny) {   // rank <= 1: complete with any unit vector orthogonal to u0
    double ax = fabs(U[0][0]), ay = fabs(U[1][0]), az = fabs(U[2][0]);
    double e[3] = {0, 0, 0};
    e[(ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2)] = 1.0;
    double d = e[0] * U[0][0] + e[1] * U[1][0] + e[2] * U[2][0];
    double v[3] = {e[0] - d * U[0][0], e[1] - d * U[1][0], e[2] - d * U[2][0]};
    double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (nv < 1e-300) { v[0] = 1; v[1] = 0; v[2] = 0; nv = 1; }
    for (int k = 0; k < 3; ++k) U[k][1] = v[k] / nv;
  }
  if (sig[ord[2]] <= tiny) {   // rank <= 2: u2 = u0 x u1 (sign is absorbed by the det fix)
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
  const double sgn = (det3(U) * det3(W) < 0) ? -1.0 : 1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R[i][j] = U[i][0] * W[j][0] + U[i][1] * W[j][1] + sgn * U[i][2] * W[j][2];
}

