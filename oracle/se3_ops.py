"""ORACLE (test infrastructure, not product code): CPU restatement of the SE3 group math.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this. The product path (batrack_b200/) never does.

Restates, in batched torch (fp32 or fp64, CPU), the per-element Eigen math of the reference's
`lietorch_backends` extension, which cannot be compiled here (Eigen 3.4.0 is not vendored:
reference setup.py:20, README.md:46-49):

  quaternion normalise-on-load ........ main/backend/lietorch/include/so3.h:31-37
  quaternion conjugate / product ...... so3.h:43-45, 51-53 (product is re-normalised by the ctor)
  rotate a point ...................... so3.h:55-60
  SO3 Exp with Taylor branch .......... so3.h:153-170   (EPS = 1e-6, common.h:7)
  SO3 Log (atan form) ................. so3.h:115-151
  SO3 left Jacobian (+ inverse) ....... so3.h:172-190, 192-208
  SE3 inv / mul / act4 ................ se3.h:36-38, 45-47, 53-56
  SE3 Adj, AdjT ....................... se3.h:58-67, 84-86
  SE3 Exp / Log ....................... se3.h:134-142, 124-132

Data layout: SE3 element = [tx, ty, tz, qx, qy, qz, qw]; tangent = [tau(3), phi(3)].
All functions take 2-D contiguous [B, dim] tensors, like the extension's entry points
(main/backend/lietorch/src/lietorch.cpp:18-283).
"""
import torch

EPS = 1e-6


def _split(X):
    return X[:, 0:3], X[:, 3:7]


def quat_normalize(q):
    # Eigen::Quaternion::normalize(): coeffs /= sqrt(squaredNorm)
    return q / torch.sqrt((q * q).sum(dim=1, keepdim=True))


def quat_mul(a, b):
    """Hamilton product of [x,y,z,w] quaternions (Eigen operator*)."""
    ax, ay, az, aw = a.unbind(dim=1)
    bx, by, bz, bw = b.unbind(dim=1)
    return torch.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz,
    ], dim=1)


def quat_conj(q):
    return torch.cat([-q[:, :3], q[:, 3:]], dim=1)


def quat_rotate(q, p):
    """so3.h:55-60: uv = 2 (q.vec x p); p + w*uv + q.vec x uv."""
    qv, w = q[:, :3], q[:, 3:4]
    uv = torch.linalg.cross(qv, p, dim=1)
    uv = uv + uv
    return p + w * uv + torch.linalg.cross(qv, uv, dim=1)


def quat_to_matrix(q):
    """Eigen::Quaternion::toRotationMatrix()."""
    x, y, z, w = q.unbind(dim=1)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = torch.stack([
        1 - (tyy + tzz), txy - twz, txz + twy,
        txy + twz, 1 - (txx + tzz), tyz - twx,
        txz - twy, tyz + twx, 1 - (txx + tyy),
    ], dim=1)
    return R.view(-1, 3, 3)


def hat(phi):
    x, y, z = phi.unbind(dim=1)
    o = torch.zeros_like(x)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=1).view(-1, 3, 3)


def so3_exp(phi):
    theta2 = (phi * phi).sum(dim=1)
    theta = torch.sqrt(theta2)
    small = theta < EPS
    theta4 = theta2 * theta2
    imag_s = 0.5 - (1.0 / 48.0) * theta2 + (1.0 / 3840.0) * theta4
    real_s = 1.0 - (1.0 / 8.0) * theta2 + (1.0 / 384.0) * theta4
    safe = torch.where(small, torch.ones_like(theta), theta)
    imag_l = torch.sin(0.5 * safe) / safe
    real_l = torch.cos(0.5 * safe)
    imag = torch.where(small, imag_s, imag_l)
    real = torch.where(small, real_s, real_l)
    q = torch.cat([imag[:, None] * phi, real[:, None]], dim=1)
    return quat_normalize(q)      # SO3(Quaternion) ctor normalises


def so3_left_jacobian(phi):
    I = torch.eye(3, dtype=phi.dtype).expand(phi.shape[0], 3, 3)
    Phi = hat(phi)
    Phi2 = Phi @ Phi
    theta2 = (phi * phi).sum(dim=1)
    theta = torch.sqrt(theta2)
    small = theta < EPS
    safe2 = torch.where(small, torch.ones_like(theta2), theta2)
    safe = torch.where(small, torch.ones_like(theta), theta)
    c1 = torch.where(small, 0.5 - (1.0 / 24.0) * theta2, (1.0 - torch.cos(safe)) / safe2)
    c2 = torch.where(small, 1.0 / 6.0 - (1.0 / 120.0) * theta2, (safe - torch.sin(safe)) / (safe2 * safe))
    return I + c1[:, None, None] * Phi + c2[:, None, None] * Phi2


def so3_left_jacobian_inverse(phi):
    I = torch.eye(3, dtype=phi.dtype).expand(phi.shape[0], 3, 3)
    Phi = hat(phi)
    Phi2 = Phi @ Phi
    theta2 = (phi * phi).sum(dim=1)
    theta = torch.sqrt(theta2)
    half = 0.5 * theta
    small = theta < EPS
    safe = torch.where(small, torch.ones_like(theta), theta)
    shalf = torch.where(small, torch.ones_like(theta), half)
    c2 = torch.where(small, torch.full_like(theta, 1.0 / 12.0),
                     (1.0 - safe * torch.cos(shalf) / (2.0 * torch.sin(shalf))) / (safe * safe))
    return I - 0.5 * Phi + c2[:, None, None] * Phi2


def so3_log(q):
    qv, w = q[:, :3], q[:, 3]
    sq_n = (qv * qv).sum(dim=1)
    n = torch.sqrt(sq_n)
    small = sq_n < EPS * EPS
    safe_n = torch.where(small, torch.ones_like(n), n)
    f_small = 2.0 / w - (2.0 / 3.0) * sq_n / (w * w * w)
    pi = 3.14159265358979323846
    f_w0 = torch.where(w > 0, pi / safe_n, -pi / safe_n)
    safe_w = torch.where(w.abs() < EPS, torch.ones_like(w), w)
    f_reg = 2.0 * torch.atan(safe_n / safe_w) / safe_n
    f = torch.where(small, f_small, torch.where(w.abs() < EPS, f_w0, f_reg))
    return f[:, None] * qv


# ---- SE3 entry points (the slice of lietorch.cpp:286-316 the BA path reaches) -----------------

def se3_load(X):
    t, q = _split(X)
    return t, quat_normalize(q)


def se3_inv(X):
    t, q = se3_load(X)
    qi = quat_normalize(quat_conj(q))          # so3.inv() -> SO3(Quaternion) ctor normalises
    ti = -quat_rotate(qi, t)
    return torch.cat([ti, qi], dim=1)


def se3_mul(X, Y):
    tx, qx = se3_load(X)
    ty, qy = se3_load(Y)
    q = quat_normalize(quat_mul(qx, qy))
    t = tx + quat_rotate(qx, ty)
    return torch.cat([t, q], dim=1)


def se3_act4(X, p):
    t, q = se3_load(X)
    p3 = quat_rotate(q, p[:, :3]) + t * p[:, 3:4]
    return torch.cat([p3, p[:, 3:4]], dim=1)


def se3_act3(X, p):
    t, q = se3_load(X)
    return quat_rotate(q, p) + t


def se3_adj_matrix(X):
    t, q = se3_load(X)
    R = quat_to_matrix(q)
    tR = hat(t) @ R
    Z = torch.zeros_like(R)
    top = torch.cat([R, tR], dim=2)
    bot = torch.cat([Z, R], dim=2)
    return torch.cat([top, bot], dim=1)


def se3_adj(X, a):
    return (se3_adj_matrix(X) @ a[:, :, None])[:, :, 0]


def se3_adjT(X, a):
    return (se3_adj_matrix(X).transpose(1, 2) @ a[:, :, None])[:, :, 0]


def se3_exp(a):
    tau, phi = a[:, :3], a[:, 3:]
    q = so3_exp(phi)
    t = (so3_left_jacobian(phi) @ tau[:, :, None])[:, :, 0]
    return torch.cat([t, q], dim=1)


def se3_log(X):
    t, q = se3_load(X)
    phi = so3_log(q)
    Vinv = so3_left_jacobian_inverse(phi)
    tau = (Vinv @ t[:, :, None])[:, :, 0]
    return torch.cat([tau, phi], dim=1)


def se3_matrix(X):
    t, q = se3_load(X)
    T = torch.zeros(X.shape[0], 4, 4, dtype=X.dtype)
    T[:, :3, :3] = quat_to_matrix(q)
    T[:, :3, 3] = t
    T[:, 3, 3] = 1
    return T
