// Small FP64 vector helpers (3-vectors, 3x3 row-major matrices, 6-d spatial vectors [linear; angular]).
// Compiles for device (nvcc) and host (g++, used by the CPU emulation of the kernels in tests/).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PLM_HD __host__ __device__ __forceinline__
#else
#define PLM_HD inline
#endif

namespace plm {

PLM_HD void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
// o += a x b
PLM_HD void cross3_acc(const double* a, const double* b, double* o) {
  o[0] += a[1] * b[2] - a[2] * b[1];
  o[1] += a[2] * b[0] - a[0] * b[2];
  o[2] += a[0] * b[1] - a[1] * b[0];
}
PLM_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
PLM_HD double dot6(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
// o = R v (R row-major)
PLM_HD void matvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
// o = R^T v
PLM_HD void matTvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
PLM_HD void matmul3(const double* A, const double* B, double* C) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  for (int i = 0; i < 9; ++i) C[i] = t[i];
}
// motion x motion: o = a x b
PLM_HD void mxm(const double* a, const double* b, double* o) {
  double t[6];
  cross3(a + 3, b, t);
  cross3_acc(a, b + 3, t);
  cross3(a + 3, b + 3, t + 3);
  for (int i = 0; i < 6; ++i) o[i] = t[i];
}
// motion x* force: o = a x* f
PLM_HD void mxf(const double* a, const double* f, double* o) {
  double t[6];
  cross3(a + 3, f, t);
  cross3(a + 3, f + 3, t + 3);
  cross3_acc(a, f, t + 3);
  for (int i = 0; i < 6; ++i) o[i] = t[i];
}
// World-frame rigid-body inertia about the world origin: mass m, first moment mc, rotational inertia Ib
// (symmetric xx xy xz yy yz zz).  o = I * mot.
PLM_HD void inertia_mul(double m, const double* mc, const double* Ib, const double* mot, double* o) {
  const double* v = mot;
  const double* w = mot + 3;
  double t[6];
  cross3(w, mc, t);
  t[0] += m * v[0]; t[1] += m * v[1]; t[2] += m * v[2];
  cross3(mc, v, t + 3);
  t[3] += Ib[0] * w[0] + Ib[1] * w[1] + Ib[2] * w[2];
  t[4] += Ib[1] * w[0] + Ib[3] * w[1] + Ib[4] * w[2];
  t[5] += Ib[2] * w[0] + Ib[4] * w[1] + Ib[5] * w[2];
  for (int i = 0; i < 6; ++i) o[i] = t[i];
}

// Coefficients of the SO(3)/SE(3) exponential: A = sin t / t, B = (1 - cos t)/t^2, C = (t - sin t)/t^3,
// power series below t = 0.5 (the first SQP iterate has DX = 0 exactly).
PLM_HD void exp_coeffs(double t2, double* A, double* B, double* C) {
  if (t2 < 0.25) {
    double a = 0, b = 0, c = 0;
    // Horner over 11 terms: sum (-1)^i t2^i / (2i+k)!
    for (int i = 11; i >= 0; --i) {
      double k1 = (2.0 * i + 2.0) * (2.0 * i + 3.0);   // ratio for A terms (k=1)
      double k2 = (2.0 * i + 3.0) * (2.0 * i + 4.0);   // k=2
      double k3 = (2.0 * i + 4.0) * (2.0 * i + 5.0);   // k=3
      a = 1.0 - t2 * a / k1;
      b = 1.0 - t2 * b / k2;
      c = 1.0 - t2 * c / k3;
    }
    *A = a; *B = b * 0.5; *C = c / 6.0;
  } else {
    double t = sqrt(t2);
    double s = sin(t), co = cos(t);
    *A = s / t; *B = (1.0 - co) / t2; *C = (t - s) / (t2 * t);
  }
}

// R = Exp(w) (row-major)
PLM_HD void exp3(const double* w, double* R) {
  double A, B, C;
  double t2 = dot3(w, w);
  exp_coeffs(t2, &A, &B, &C);
  double xx = w[0] * w[0], yy = w[1] * w[1], zz = w[2] * w[2];
  double xy = w[0] * w[1], xz = w[0] * w[2], yz = w[1] * w[2];
  R[0] = 1.0 - B * (yy + zz); R[1] = -A * w[2] + B * xy;  R[2] = A * w[1] + B * xz;
  R[3] = A * w[2] + B * xy;   R[4] = 1.0 - B * (xx + zz); R[5] = -A * w[0] + B * yz;
  R[6] = -A * w[1] + B * xz;  R[7] = A * w[0] + B * yz;   R[8] = 1.0 - B * (xx + yy);
}

// Right Jacobian of SO(3): Jr = I - B [w]x + C [w]x^2  (d Exp(w + e) = Exp(w) Exp(Jr e))
PLM_HD void jr3(const double* w, double* Jr) {
  double A, B, C;
  double t2 = dot3(w, w);
  exp_coeffs(t2, &A, &B, &C);
  double xx = w[0] * w[0], yy = w[1] * w[1], zz = w[2] * w[2];
  double xy = w[0] * w[1], xz = w[0] * w[2], yz = w[1] * w[2];
  Jr[0] = 1.0 - C * (yy + zz); Jr[1] = B * w[2] + C * xy;   Jr[2] = -B * w[1] + C * xz;
  Jr[3] = -B * w[2] + C * xy;  Jr[4] = 1.0 - C * (xx + zz); Jr[5] = B * w[0] + C * yz;
  Jr[6] = B * w[1] + C * xz;   Jr[7] = -B * w[0] + C * yz;  Jr[8] = 1.0 - C * (xx + yy);
}

// Unit quaternion [x,y,z,w] -> row-major rotation
PLM_HD void quat_to_R(const double* q, double* R) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

}  // namespace plm
