"""`SE3`: the Lie-group wrapper that crosses the BA operator boundary in both directions.

Mirrors the slice of the reference's class the BA caller uses (main/backend/lietorch/groups.py:51-285:
`.data[...,7]`, indexing, `inv`, `mul`/`*`, `act`, `adjT`, `adj`, `exp`, `log`, `retr`, `matrix`,
`translation`, `vec`, `Identity`), forward only. Element layout [tx ty tz qx qy qz qw]; tangent
[tau, phi]. All math runs in the CUDA kernels behind batrack_b200.lietorch_backends.
"""
import torch

from .. import lietorch_backends as _be


def _flat_pair(x, y):
    """Broadcast the batch dims of x [..., a] and y [..., b] and flatten to contiguous 2-D."""
    bs = torch.broadcast_shapes(x.shape[:-1], y.shape[:-1])
    xf = x.expand(bs + x.shape[-1:]).reshape(-1, x.shape[-1]).contiguous()
    yf = y.expand(bs + y.shape[-1:]).reshape(-1, y.shape[-1]).contiguous()
    return xf, yf, tuple(bs)


class SE3:
    group_name = "SE3"
    group_id = 3
    manifold_dim = 6
    embedded_dim = 7

    def __init__(self, data):
        self.data = data

    def __repr__(self):
        return f"SE3: size={tuple(self.shape)}, device={self.device}, dtype={self.dtype}"

    # ---- shape plumbing ----
    @property
    def shape(self):
        return self.data.shape[:-1]

    @property
    def device(self):
        return self.data.device

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def tangent_shape(self):
        return self.data.shape[:-1] + (6,)

    def __getitem__(self, index):
        return SE3(self.data[index])

    def __setitem__(self, index, item):
        self.data[index] = item.data

    def view(self, dims):
        return SE3(self.data.view(tuple(dims) + (7,)))

    def detach(self):
        return SE3(self.data.detach())

    def to(self, *a, **k):
        return SE3(self.data.to(*a, **k))

    def cuda(self):
        return SE3(self.data.cuda())

    def cpu(self):
        return SE3(self.data.cpu())

    def unbind(self, dim=0):
        return [SE3(x) for x in self.data.unbind(dim=dim)]

    def vec(self):
        return self.data

    @classmethod
    def Identity(cls, *batch_shape, device="cuda", dtype=torch.float32):
        if len(batch_shape) == 1 and isinstance(batch_shape[0], (tuple, list)):
            batch_shape = tuple(batch_shape[0])
        data = torch.zeros(tuple(batch_shape) + (7,), device=device, dtype=dtype)
        data[..., 6] = 1.0
        return cls(data)

    @classmethod
    def IdentityLike(cls, G):
        return cls.Identity(tuple(G.shape), device=G.device, dtype=G.dtype)

    @classmethod
    def InitFromVec(cls, data):
        return cls(data)

    # ---- group operations ----
    def _unary(self, fn, cols):
        flat = self.data.reshape(-1, 7).contiguous()
        return fn(3, flat).view(tuple(self.shape) + cols)

    @classmethod
    def exp(cls, a):
        return cls(_be.expm(3, a.reshape(-1, 6).contiguous()).view(a.shape[:-1] + (7,)))

    def log(self):
        return self._unary(_be.logm, (6,))

    def inv(self):
        return SE3(self._unary(_be.inv, (7,)))

    def mul(self, other):
        x, y, bs = _flat_pair(self.data, other.data)
        return SE3(_be.mul(3, x, y).view(bs + (7,)))

    def retr(self, a):
        """Exp(a) * X  (groups.py:153-156)"""
        return SE3.exp(a).mul(self)

    def adj(self, a):
        x, y, bs = _flat_pair(self.data, a)
        return _be.adj(3, x, y).view(bs + (6,))

    def adjT(self, a):
        x, y, bs = _flat_pair(self.data, a)
        return _be.adjT(3, x, y).view(bs + (6,))

    def act(self, p):
        x, y, bs = _flat_pair(self.data, p)
        if p.shape[-1] == 3:
            return _be.act(3, x, y).view(bs + (3,))
        if p.shape[-1] == 4:
            return _be.act4(3, x, y).view(bs + (4,))
        raise ValueError("act: points must have 3 or 4 components")

    def matrix(self):
        return self._unary(_be.as_matrix, (4, 4))

    def translation(self):
        """groups.py:186-190: act on the homogeneous origin."""
        p = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=self.dtype, device=self.device)
        return self.act(p.view((1,) * (self.data.dim() - 1) + (4,)))

    def scale(self, s):
        t, q = self.data.split([3, 4], -1)
        return SE3(torch.cat([t * s.unsqueeze(-1), q], dim=-1))

    def __mul__(self, other):
        if isinstance(other, SE3):
            return self.mul(other)
        if isinstance(other, torch.Tensor):
            return self.act(other)
        return NotImplemented
