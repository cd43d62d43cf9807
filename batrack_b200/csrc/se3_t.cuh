// se3_t.cuh — the SE3 group math of the standalone lietorch_backends entry points, templated on the scalar type
// (the reference dispatches float and double: main/backend/lietorch/include/dispatch.h:37-45). Same formulas and Taylor
// branches as ba_math.cuh (so3.h:31-60,115-190; se3.h:36-142); the fused BA kernels keep the float-only header.
#pragma once
#include <math.h>

namespace ba {
namespace se3t {

#define SE3T_HD __host__ __device__ __forceinline__

template <typename T> struct Q { T x, y, z, w; };
template <typename T> struct V { T x, y, z; };
template <typename T> struct P { V<T> t; Q<T> q; };

template <typename T> SE3T_HD Q<T> qnormalize(Q<T> q) {
  const T n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}
template <typename T> SE3T_HD Q<T> qmul(Q<T> a, Q<T> b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
template <typename T> SE3T_HD V<T> cross(V<T> a, V<T> b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
template <typename T> SE3T_HD V<T> qrotate(Q<T> q, V<T> p) {          // so3.h:55-60
  const V<T> v{q.x, q.y, q.z};
  V<T> uv = cross(v, p);
  uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  const V<T> c = cross(v, uv);
  return {p.x + q.w * uv.x + c.x, p.y + q.w * uv.y + c.y, p.z + q.w * uv.z + c.z};
}
template <typename T> SE3T_HD void qmatrix(Q<T> q, T *R) {            // Eigen::Quaternion::toRotationMatrix()
  const T tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const T tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
template <typename T> SE3T_HD P<T> load(const T *p) { return {{p[0], p[1], p[2]}, qnormalize(Q<T>{p[3], p[4], p[5], p[6]})}; }
template <typename T> SE3T_HD void store(const P<T> &X, T *p) {
  p[0] = X.t.x; p[1] = X.t.y; p[2] = X.t.z; p[3] = X.q.x; p[4] = X.q.y; p[5] = X.q.z; p[6] = X.q.w;
}
template <typename T> SE3T_HD P<T> inv(const P<T> &X) {               // se3.h:36-38
  const Q<T> qi = qnormalize(Q<T>{-X.q.x, -X.q.y, -X.q.z, X.q.w});
  const V<T> r = qrotate(qi, X.t);
  return {{-r.x, -r.y, -r.z}, qi};
}
template <typename T> SE3T_HD P<T> mul(const P<T> &X, const P<T> &Y) { // se3.h:45-47
  const V<T> r = qrotate(X.q, Y.t);
  return {{X.t.x + r.x, X.t.y + r.y, X.t.z + r.z}, qnormalize(qmul(X.q, Y.q))};
}
template <typename T> SE3T_HD P<T> exp(const T *a) {                  // so3.h:153-170,172-190; se3.h:134-142
  const T eps = T(1e-6);
  const T px = a[3], py = a[4], pz = a[5];
  const T th2 = px * px + py * py + pz * pz, th = sqrt(th2);
  T imag, real, c1, c2;
  if (th < eps) {
    const T th4 = th2 * th2;
    imag = T(0.5) - (T(1) / T(48)) * th2 + (T(1) / T(3840)) * th4;
    real = T(1) - (T(1) / T(8)) * th2 + (T(1) / T(384)) * th4;
    c1 = T(0.5) - (T(1) / T(24)) * th2;
    c2 = (T(1) / T(6)) - (T(1) / T(120)) * th2;
  } else {
    imag = sin(T(0.5) * th) / th;
    real = cos(T(0.5) * th);
    c1 = (T(1) - cos(th)) / th2;
    c2 = (th - sin(th)) / (th2 * th);
  }
  const Q<T> q = qnormalize(Q<T>{imag * px, imag * py, imag * pz, real});
  const V<T> phi{px, py, pz}, tau{a[0], a[1], a[2]};
  const V<T> c = cross(phi, tau), cc = cross(phi, c);
  return {{tau.x + c1 * c.x + c2 * cc.x, tau.y + c1 * c.y + c2 * cc.y, tau.z + c1 * c.z + c2 * cc.z}, q};
}
template <typename T> SE3T_HD void log(const P<T> &X, T *a) {         // so3.h:115-151; se3.h:124-132
  const T eps = T(1e-6), pi = T(3.14159265358979323846);
  const T sq = X.q.x * X.q.x + X.q.y * X.q.y + X.q.z * X.q.z, n = sqrt(sq), w = X.q.w;
  T f;
  if (sq < eps * eps) f = T(2) / w - (T(2) / T(3)) * sq / (w * w * w);
  else if (fabs(w) < eps) f = (w > 0 ? pi : -pi) / n;
  else f = T(2) * atan(n / w) / n;
  const V<T> phi{f * X.q.x, f * X.q.y, f * X.q.z};
  const T th2 = phi.x * phi.x + phi.y * phi.y + phi.z * phi.z, th = sqrt(th2), half = T(0.5) * th;
  const T c2 = (th < eps) ? (T(1) / T(12)) : (T(1) - th * cos(half) / (T(2) * sin(half))) / (th * th);
  const V<T> c = cross(phi, X.t), cc = cross(phi, c);
  a[0] = X.t.x - T(0.5) * c.x + c2 * cc.x; a[1] = X.t.y - T(0.5) * c.y + c2 * cc.y; a[2] = X.t.z - T(0.5) * c.z + c2 * cc.z;
  a[3] = phi.x; a[4] = phi.y; a[5] = phi.z;
}
template <typename T> SE3T_HD void adjT(const T *R, V<T> t, const T *a, T *b) {   // se3.h:84-86
  const V<T> at{a[0], a[1], a[2]};
  const V<T> c = cross(t, at);
  const T u0 = a[3] - c.x, u1 = a[4] - c.y, u2 = a[5] - c.z;
  b[0] = R[0] * at.x + R[3] * at.y + R[6] * at.z; b[1] = R[1] * at.x + R[4] * at.y + R[7] * at.z; b[2] = R[2] * at.x + R[5] * at.y + R[8] * at.z;
  b[3] = R[0] * u0 + R[3] * u1 + R[6] * u2; b[4] = R[1] * u0 + R[4] * u1 + R[7] * u2; b[5] = R[2] * u0 + R[5] * u1 + R[8] * u2;
}
template <typename T> SE3T_HD void adj(const T *R, V<T> t, const T *a, T *b) {    // se3.h:58-67
  const V<T> rt{R[0] * a[0] + R[1] * a[1] + R[2] * a[2], R[3] * a[0] + R[4] * a[1] + R[5] * a[2], R[6] * a[0] + R[7] * a[1] + R[8] * a[2]};
  const V<T> rp{R[0] * a[3] + R[1] * a[4] + R[2] * a[5], R[3] * a[3] + R[4] * a[4] + R[5] * a[5], R[6] * a[3] + R[7] * a[4] + R[8] * a[5]};
  const V<T> c = cross(t, rp);
  b[0] = rt.x + c.x; b[1] = rt.y + c.y; b[2] = rt.z + c.z; b[3] = rp.x; b[4] = rp.y; b[5] = rp.z;
}

}  // namespace se3t
}  // namespace ba
