// ba_math.cuh — per-element math of the BA hot path, usable from device and (for CPU unit tests of
// the formulas only) host code.
//
// SE3 element = [tx ty tz qx qy qz qw], tangent = [tau(3) phi(3)]. The group math follows the
// reference's Eigen templates (main/backend/lietorch/include/so3.h, se3.h): quaternions are
// re-normalised on every load (so3.h:31-37) and after every product (so3.h:51-53).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define BA_HD __host__ __device__ __forceinline__
#else
#define BA_HD inline
#endif

namespace ba {

constexpr float kEps = 1e-6f;          // common.h:7
constexpr float kMinDepth = 0.2f;      // projective_ops.py:7

struct Quat { float x, y, z, w; };
struct Vec3 { float x, y, z; };
struct Pose { Vec3 t; Quat q; };

BA_HD Quat qnormalize(Quat q) {        // Eigen::Quaternion::normalize()
  float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}
BA_HD Quat qmul(Quat a, Quat b) {      // Eigen quaternion product
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
          a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
BA_HD Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
BA_HD Vec3 qrotate(Quat q, Vec3 p) {   // so3.h:55-60
  Vec3 v{q.x, q.y, q.z};
  Vec3 uv = cross(v, p);
  uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  Vec3 c = cross(v, uv);
  return {p.x + q.w * uv.x + c.x, p.y + q.w * uv.y + c.y, p.z + q.w * uv.z + c.z};
}
// Eigen::Quaternion::toRotationMatrix(), row-major R[9]
BA_HD void qmatrix(Quat q, float *R) {
  float tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

BA_HD Pose pose_load(const float *p) { // SE3(const Scalar*) -> SO3(const Scalar*) normalises
  return {{p[0], p[1], p[2]}, qnormalize({p[3], p[4], p[5], p[6]})};
}
BA_HD void pose_store(const Pose &P, float *p) {
  p[0] = P.t.x; p[1] = P.t.y; p[2] = P.t.z; p[3] = P.q.x; p[4] = P.q.y; p[5] = P.q.z; p[6] = P.q.w;
}
BA_HD Pose pose_inv(const Pose &X) {   // se3.h:36-38
  Quat qi = qnormalize({-X.q.x, -X.q.y, -X.q.z, X.q.w});
  Vec3 r = qrotate(qi, X.t);
  return {{-r.x, -r.y, -r.z}, qi};
}
BA_HD Pose pose_mul(const Pose &X, const Pose &Y) {   // se3.h:45-47
  Vec3 r = qrotate(X.q, Y.t);
  return {{X.t.x + r.x, X.t.y + r.y, X.t.z + r.z}, qnormalize(qmul(X.q, Y.q))};
}

// SO3::Exp (so3.h:153-170) and SE3::Exp (se3.h:134-142) with the Taylor branches
BA_HD Pose pose_exp(const float *a) {
  float px = a[3], py = a[4], pz = a[5];
  float th2 = px * px + py * py + pz * pz, th = sqrtf(th2);
  float imag, real, c1, c2;
  if (th < kEps) {
    float th4 = th2 * th2;
    imag = 0.5f - (1.0f / 48.0f) * th2 + (1.0f / 3840.0f) * th4;
    real = 1.0f - (1.0f / 8.0f) * th2 + (1.0f / 384.0f) * th4;
    c1 = 0.5f - (1.0f / 24.0f) * th2;                 // so3.h:181-187
    c2 = (1.0f / 6.0f) - (1.0f / 120.0f) * th2;
  } else {
    imag = sinf(0.5f * th) / th;
    real = cosf(0.5f * th);
    c1 = (1.0f - cosf(th)) / th2;
    c2 = (th - sinf(th)) / (th2 * th);
  }
  Quat q = qnormalize({imag * px, imag * py, imag * pz, real});
  // V = I + c1 Phi + c2 Phi^2 ; t = V tau
  Vec3 phi{px, py, pz}, tau{a[0], a[1], a[2]};
  Vec3 c = cross(phi, tau), cc = cross(phi, c);
  return {{tau.x + c1 * c.x + c2 * cc.x, tau.y + c1 * c.y + c2 * cc.y, tau.z + c1 * c.z + c2 * cc.z}, q};
}

// SO3::Log (so3.h:115-151) + SE3::Log (se3.h:124-132)
BA_HD void pose_log(const Pose &X, float *a) {
  float sq = X.q.x * X.q.x + X.q.y * X.q.y + X.q.z * X.q.z, n = sqrtf(sq), w = X.q.w, f;
  if (sq < kEps * kEps) {
    f = 2.0f / w - (2.0f / 3.0f) * sq / (w * w * w);
  } else if (fabsf(w) < kEps) {
    f = (w > 0 ? 3.14159265358979323846f : -3.14159265358979323846f) / n;
  } else {
    f = 2.0f * atanf(n / w) / n;
  }
  Vec3 phi{f * X.q.x, f * X.q.y, f * X.q.z};
  float th2 = phi.x * phi.x + phi.y * phi.y + phi.z * phi.z, th = sqrtf(th2), half = 0.5f * th;
  float c2 = (th < kEps) ? (1.0f / 12.0f) : (1.0f - th * cosf(half) / (2.0f * sinf(half))) / (th * th);
  Vec3 c = cross(phi, X.t), cc = cross(phi, c);        // Vinv = I - Phi/2 + c2 Phi^2
  a[0] = X.t.x - 0.5f * c.x + c2 * cc.x; a[1] = X.t.y - 0.5f * c.y + c2 * cc.y; a[2] = X.t.z - 0.5f * c.z + c2 * cc.z;
  a[3] = phi.x; a[4] = phi.y; a[5] = phi.z;
}

// Ad(X)^T a  with Ad = [[R, t^R],[0, R]] (se3.h:58-67,84-86):  [R^T a_tau ; R^T (a_phi - t x a_tau)]
BA_HD void adjT_apply(const float *R, Vec3 t, const float *a, float *b) {
  Vec3 at{a[0], a[1], a[2]};
  Vec3 c = cross(t, at);
  float u0 = a[3] - c.x, u1 = a[4] - c.y, u2 = a[5] - c.z;
  b[0] = R[0] * at.x + R[3] * at.y + R[6] * at.z;
  b[1] = R[1] * at.x + R[4] * at.y + R[7] * at.z;
  b[2] = R[2] * at.x + R[5] * at.y + R[8] * at.z;
  b[3] = R[0] * u0 + R[3] * u1 + R[6] * u2;
  b[4] = R[1] * u0 + R[4] * u1 + R[7] * u2;
  b[5] = R[2] * u0 + R[5] * u1 + R[8] * u2;
}
// Ad(X) a = [R a_tau + t x (R a_phi) ; R a_phi]
BA_HD void adj_apply(const float *R, Vec3 t, const float *a, float *b) {
  Vec3 rt{R[0] * a[0] + R[1] * a[1] + R[2] * a[2], R[3] * a[0] + R[4] * a[1] + R[5] * a[2], R[6] * a[0] + R[7] * a[1] + R[8] * a[2]};
  Vec3 rp{R[0] * a[3] + R[1] * a[4] + R[2] * a[5], R[3] * a[3] + R[4] * a[4] + R[5] * a[5], R[6] * a[3] + R[7] * a[4] + R[8] * a[5]};
  Vec3 c = cross(t, rp);
  b[0] = rt.x + c.x; b[1] = rt.y + c.y; b[2] = rt.z + c.z; b[3] = rp.x; b[4] = rp.y; b[5] = rp.z;
}

// ---- one (source frame i, target frame j) pair: everything that does not depend on the patch ----
struct PairConst {
  float R[9];          // rotation of Gij = Tj * Ti^-1
  Vec3 t;              // translation of Gij
  float fxi, fyi, cxi, cyi, fxj, fyj, cxj, cyj;
};

BA_HD PairConst pair_const(const float *pose_i, const float *pose_j, const float *intr_i, const float *intr_j) {
  Pose Gij = pose_mul(pose_load(pose_j), pose_inv(pose_load(pose_i)));   // projective_ops.py:61
  // the reference re-loads Gij (normalising again) inside act4 / adjT (lietorch_gpu.cu:225,161)
  Gij.q = qnormalize(Gij.q);
  PairConst c;
  qmatrix(Gij.q, c.R);
  c.t = Gij.t;
  c.fxi = intr_i[0]; c.fyi = intr_i[1]; c.cxi = intr_i[2]; c.cyi = intr_i[3];
  c.fxj = intr_j[0]; c.fyj = intr_j[1]; c.cxj = intr_j[2]; c.cyj = intr_j[3];
  return c;
}

// What one edge contributes, before any summation.
struct EdgeTerms {
  float u, v;          // reprojected pixel (projective_ops.py:43-45)
  float Jj0[6], Jj1[6];// d(pixel)/d(pose j), rows x and y (projective_ops.py:83-95)
  float Jz0, Jz1;      // d(pixel)/d(inverse depth) (projective_ops.py:98)
  float r0, r1;        // masked residual (ba.py:226,250)
  float w0, w1;        // masked, robustified weights (ba.py:247-251)
  float valid;
};

// fast reciprocal / reciprocal square root on the device (a few ulp; the per-edge terms are fp32 anyway and
// parity is judged at 1e-4 against fp64), exact forms on the host
#ifdef __CUDA_ARCH__
BA_HD float ba_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }      // 1 ulp
BA_HD float ba_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }  // 2 ulp
#else
BA_HD float ba_rcp(float x) { return 1.0f / x; }
BA_HD float ba_rsqrt(float x) { return 1.0f / sqrtf(x); }
#endif

BA_HD float robust_weight(float r, int loss) {        // ba.py:81-100
  float s = r * r;
  if (loss == 1) return s > 1.0f ? ba_rsqrt(s) : 1.0f;   // huber: 1/|r|
  if (loss == 2) return ba_rcp(1.0f + s);
  return 1.0f;
}

// ifxi / ifyi = 1/fx_i, 1/fy_i (per-position constants; (x - cx) * (1/fx) differs from the reference's
// (x - cx) / fx by at most one fp32 ulp)
BA_HD void edge_terms(const PairConst &c, float ifxi, float ifyi, float px, float py, float pd, float tx, float ty,
                      float wx, float wy, const float *bounds, int loss, EdgeTerms &o) {
  // iproj (projective_ops.py:19-29)
  float x0 = (px - c.cxi) * ifxi, y0 = (py - c.cyi) * ifyi;
  // act4 (se3.h:53-56): R * X0[:3] + t * X0[3]
  float X = c.R[0] * x0 + c.R[1] * y0 + c.R[2] + c.t.x * pd;
  float Y = c.R[3] * x0 + c.R[4] * y0 + c.R[5] + c.t.y * pd;
  float Z = c.R[6] * x0 + c.R[7] * y0 + c.R[8] + c.t.z * pd;
  float H = pd;
  // proj (projective_ops.py:43-45)
  float dc = ba_rcp(fmaxf(Z, 1e-2f));
  o.u = c.fxj * (dc * X) + c.cxj;
  o.v = c.fyj * (dc * Y) + c.cyj;
  // Jacobians (projective_ops.py:80-98)
  float dj = fabsf(Z) > kMinDepth ? ba_rcp(Z) : 0.0f;
  float a = c.fxj * dj, b = -c.fxj * X * dj * dj;
  float cc = c.fyj * dj, e = -c.fyj * Y * dj * dj;
  o.Jj0[0] = a * H; o.Jj0[1] = 0.0f;   o.Jj0[2] = b * H; o.Jj0[3] = b * Y;           o.Jj0[4] = a * Z - b * X; o.Jj0[5] = -a * Y;
  o.Jj1[0] = 0.0f;  o.Jj1[1] = cc * H; o.Jj1[2] = e * H; o.Jj1[3] = -cc * Z + e * Y; o.Jj1[4] = -e * X;        o.Jj1[5] = cc * X;
  o.Jz0 = a * c.t.x + b * c.t.z;
  o.Jz1 = cc * c.t.y + e * c.t.z;
  // residual + validity (ba.py:226-242, projective_ops.py:100)
  float r0 = tx - o.u, r1 = ty - o.v;
  bool ok = Z > kMinDepth;
  ok = ok && (r0 * r0 + r1 * r1 < 250.0f * 250.0f);      // ||r|| < 250
  ok = ok && (o.u > bounds[0]) && (o.v > bounds[1]) && (o.u < bounds[2]) && (o.v < bounds[3]);
  // the mask MULTIPLIES like the reference's (ba.py:241-242, :250): a NaN / inf target or weight of a rejected edge still
  // poisons the sums exactly as it does there (0 * NaN = NaN), which is what reaches the NaN retry of ba.py:324-325
  o.valid = ok ? 1.0f : 0.0f;
  o.w0 = o.valid * (wx * robust_weight(r0, loss));
  o.w1 = o.valid * (wy * robust_weight(r1, loss));
  o.r0 = o.valid * r0;
  o.r1 = o.valid * r1;
}

// index of (a,b), a >= b, in a packed lower-triangular 6x6 (21 entries)
BA_HD constexpr int tri(int a, int b) { return a * (a + 1) / 2 + b; }

}  // namespace ba
