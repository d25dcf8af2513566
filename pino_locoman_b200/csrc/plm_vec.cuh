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
    // (the ratios are compile-time reciprocals once the loop is unrolled: no FP64 divisions on the device)
#pragma unroll
    for (int i = 11; i >= 0; --i) {
      const double r1 = 1.0 / ((2.0 * i + 2.0) * (2.0 * i + 3.0));   // ratio for A terms (k=1)
      const double r2 = 1.0 / ((2.0 * i + 3.0) * (2.0 * i + 4.0));   // k=2
      const double r3 = 1.0 / ((2.0 * i + 4.0) * (2.0 * i + 5.0));   // k=3
      a = 1.0 - t2 * a * r1;
      b = 1.0 - t2 * b * r2;
      c = 1.0 - t2 * c * r3;
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

// SE(3) logarithm of M0^-1 M1 from positions and unit quaternions [x,y,z,w]  (pin.difference, base part)
PLM_HD void se3_difference(const double* p0, const double* q0, const double* p1, const double* q1, double* nu) {
  // relative quaternion qr = conj(q0) * q1
  double ax = -q0[0], ay = -q0[1], az = -q0[2], aw = q0[3];
  double bx = q1[0], by = q1[1], bz = q1[2], bw = q1[3];
  double qr[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                  aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz};
  if (qr[3] < 0) { qr[0] = -qr[0]; qr[1] = -qr[1]; qr[2] = -qr[2]; qr[3] = -qr[3]; }
  double n = sqrt(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2]);
  double w[3];
  if (n < 1e-12) {
    double s = 2.0 / qr[3];
    w[0] = s * qr[0]; w[1] = s * qr[1]; w[2] = s * qr[2];
  } else {
    double s = 2.0 * atan2(n, qr[3]) / n;
    w[0] = s * qr[0]; w[1] = s * qr[1]; w[2] = s * qr[2];
  }
  double R0[9], d[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, pl[3];
  quat_to_R(q0, R0);
  matTvec3(R0, d, pl);
  double t2 = dot3(w, w), beta;
  if (t2 < 1e-2) beta = 1.0 / 12 + t2 / 720 + t2 * t2 / 30240 + t2 * t2 * t2 / 1209600;
  else {
    double t = sqrt(t2);
    beta = 1.0 / t2 - (1.0 + cos(t)) / (2.0 * t * sin(t));
  }
  // Vinv p = p - 0.5 w x p + beta w x (w x p)
  double wp[3], wwp[3];
  cross3(w, pl, wp);
  cross3(w, wp, wwp);
  for (int i = 0; i < 3; ++i) nu[i] = pl[i] - 0.5 * wp[i] + beta * wwp[i];
  nu[3] = w[0]; nu[4] = w[1]; nu[5] = w[2];
}


// pin.integrate on SE(3): (p, quat) * exp6([rho; w]); quaternion [x,y,z,w], renormalised
PLM_HD void se3_integrate(const double* p0, const double* q0, const double* nu, double* p1, double* q1) {
  const double* rho = nu;
  const double* w = nu + 3;
  double A, B, C;
  const double t2 = dot3(w, w);
  exp_coeffs(t2, &A, &B, &C);
  double wr[3], wwr[3], pl[3], R0[9], pw[3];
  cross3(w, rho, wr);
  cross3(w, wr, wwr);
  for (int i = 0; i < 3; ++i) pl[i] = rho[i] + B * wr[i] + C * wwr[i];     // V(w) rho
  quat_to_R(q0, R0);
  matvec3(R0, pl, pw);
  for (int i = 0; i < 3; ++i) p1[i] = p0[i] + pw[i];
  // quaternion of Exp(w): [sin(t/2)/t w, cos(t/2)] with sin(t/2)/t = 0.5 * sinc(t/2)
  double Ah, Bh, Ch;
  exp_coeffs(0.25 * t2, &Ah, &Bh, &Ch);
  const double s = 0.5 * Ah, c = 1.0 - 0.25 * t2 * Bh;
  const double e[4] = {s * w[0], s * w[1], s * w[2], c};
  double r[4] = {q0[3] * e[0] + q0[0] * e[3] + q0[1] * e[2] - q0[2] * e[1],
                 q0[3] * e[1] - q0[0] * e[2] + q0[1] * e[3] + q0[2] * e[0],
                 q0[3] * e[2] + q0[0] * e[1] - q0[1] * e[0] + q0[2] * e[3],
                 q0[3] * e[3] - q0[0] * e[0] - q0[1] * e[1] - q0[2] * e[2]};
  const double nrm = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
  for (int i = 0; i < 4; ++i) q1[i] = r[i] / nrm;
}

}  // namespace plm
