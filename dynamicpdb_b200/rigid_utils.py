"""Quaternion / rotation / rigid-frame algebra with the call surface of the reference's
``openfold/utils/rigid_utils.py``.

This is a from-scratch implementation: the reference builds rotations by contracting lookup tables
(``_QTR_MAT`` rigid_utils.py:173-205, ``_QUAT_MULTIPLY`` :230-275); here every product is written as the
closed-form component expression, which is what the CUDA kernels in ``csrc/rigid.cu`` keep in registers.
On CUDA tensors the hot entry points (``quat_to_rot``, ``Rigid.apply`` / ``invert_apply`` on quaternion
frames, ``Rigid.compose_q_update_vec``) run the hand-written kernels through ``kernels``; the remaining
methods are thin tensor plumbing around them.

Public names and behaviours mirrored (reference file:line):
  rot_matmul :22, rot_vec_mul :82, identity_rot_mats :109, identity_trans :124, identity_quats :139,
  quat_to_rot :185, rot_to_quat :208, quat_multiply :254, quat_multiply_by_vec :266,
  invert_rot_mat :278, invert_quat :282, Rotation :289-850, Rigid :853-1448
  (incl. the fork's ``update_mask`` keyword, :590 and :1041).
"""
from typing import Any, Callable, Optional, Sequence, Tuple

import torch

from . import kernels as _k


# --------------------------------------------------------------------------------------------------
# free functions
# --------------------------------------------------------------------------------------------------
def rot_matmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """[*,3,3] @ [*,3,3] written element-wise (no autocast down-casting). ref :22-79."""
    if a.is_cuda and b.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32:
        return _k.rot_compose(a, None, b, None)[0]
    a0, a1, a2 = a.unbind(-2)  # rows of a: each [*,3]
    rows = []
    for ar in (a0, a1, a2):
        # row_i = sum_k a[i,k] * b[k,:]
        rows.append(ar[..., 0:1] * b[..., 0, :] + ar[..., 1:2] * b[..., 1, :] + ar[..., 2:3] * b[..., 2, :])
    return torch.stack(rows, dim=-2)


def rot_vec_mul(r: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """[*,3,3] applied to [*,3]. ref :82-106."""
    if r.is_cuda and t.is_cuda and r.dtype == torch.float32 and t.dtype == torch.float32:
        return _k.rot_compose(r, None, None, t)[1]
    x, y, z = t[..., 0:1], t[..., 1:2], t[..., 2:3]
    return r[..., :, 0] * x + r[..., :, 1] * y + r[..., :, 2] * z


def identity_rot_mats(batch_dims, dtype=None, device=None, requires_grad: bool = True) -> torch.Tensor:
    eye = torch.eye(3, dtype=dtype, device=device, requires_grad=requires_grad)
    return eye.view(*((1,) * len(batch_dims)), 3, 3).expand(*batch_dims, -1, -1)


def identity_trans(batch_dims, dtype=None, device=None, requires_grad: bool = True) -> torch.Tensor:
    return torch.zeros((*batch_dims, 3), dtype=dtype, device=device, requires_grad=requires_grad)


def identity_quats(batch_dims, dtype=None, device=None, requires_grad: bool = True) -> torch.Tensor:
    q = torch.zeros((*batch_dims, 4), dtype=dtype, device=device, requires_grad=requires_grad)
    with torch.no_grad():
        q[..., 0] = 1
    return q


def _quat_to_rot_torch(quat: torch.Tensor) -> torch.Tensor:
    a, b, c, d = quat.unbind(-1)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    ab, ac, ad = a * b, a * c, a * d
    bc, bd, cd = b * c, b * d, c * d
    rows = [
        torch.stack([aa + bb - cc - dd, 2 * (bc - ad), 2 * (bd + ac)], dim=-1),
        torch.stack([2 * (bc + ad), aa - bb + cc - dd, 2 * (cd - ab)], dim=-1),
        torch.stack([2 * (bd - ac), 2 * (cd + ab), aa - bb - cc + dd], dim=-1),
    ]
    return torch.stack(rows, dim=-2)


def quat_to_rot(quat: torch.Tensor) -> torch.Tensor:
    """(w,x,y,z) -> [*,3,3]; NOT normalised (a non-unit quaternion gives |q|^2 R). ref :185-205."""
    if quat.is_cuda and quat.dtype == torch.float32:
        return _k.quat_to_rot(quat)
    return _quat_to_rot_torch(quat)


def rot_to_quat(rot: torch.Tensor) -> torch.Tensor:
    """Rotation matrix -> quaternion as the top eigenvector of the symmetric 4x4 K matrix. ref :208-227.

    The reference moves K to the host for ``eigh`` (:226); here the decomposition runs on the tensor's
    own device (no host sync).  The eigenvector sign is arbitrary in both.
    """
    if rot.shape[-2:] != (3, 3):
        raise ValueError("Input rotation is incorrectly shaped")
    xx, xy, xz = rot[..., 0, 0], rot[..., 0, 1], rot[..., 0, 2]
    yx, yy, yz = rot[..., 1, 0], rot[..., 1, 1], rot[..., 1, 2]
    zx, zy, zz = rot[..., 2, 0], rot[..., 2, 1], rot[..., 2, 2]
    k = torch.stack(
        [
            torch.stack([xx + yy + zz, zy - yz, xz - zx, yx - xy], dim=-1),
            torch.stack([zy - yz, xx - yy - zz, xy + yx, xz + zx], dim=-1),
            torch.stack([xz - zx, xy + yx, yy - xx - zz, yz + zy], dim=-1),
            torch.stack([yx - xy, xz + zx, yz + zy, zz - xx - yy], dim=-1),
        ],
        dim=-2,
    ) * (1.0 / 3.0)
    _, vecs = torch.linalg.eigh(k)
    return vecs[..., -1]


def quat_multiply(quat1: torch.Tensor, quat2: torch.Tensor) -> torch.Tensor:
    """Hamilton product. ref :254-263."""
    if quat1.is_cuda and quat2.is_cuda and quat1.dtype == torch.float32 and quat2.dtype == torch.float32:
        return _k.quat_mul(quat1, quat2)
    a1, b1, c1, d1 = quat1.unbind(-1)
    a2, b2, c2, d2 = quat2.unbind(-1)
    return torch.stack(
        [
            a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2,
            a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
            a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2,
            a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2,
        ],
        dim=-1,
    )


def quat_multiply_by_vec(quat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """quat * (0, vec). ref :266-275."""
    if quat.is_cuda and vec.is_cuda and quat.dtype == torch.float32 and vec.dtype == torch.float32:
        return _k.quat_mul(quat, vec, b_is_vec=True)
    a, b, c, d = quat.unbind(-1)
    x, y, z = vec.unbind(-1)
    return torch.stack(
        [
            -b * x - c * y - d * z,
            a * x + c * z - d * y,
            a * y - b * z + d * x,
            a * z + b * y - c * x,
        ],
        dim=-1,
    )


def invert_rot_mat(rot_mat: torch.Tensor) -> torch.Tensor:
    return rot_mat.transpose(-1, -2)


def invert_quat(quat: torch.Tensor) -> torch.Tensor:
    """Conjugate / |q|^2. ref :282-286."""
    conj = torch.cat([quat[..., :1], -quat[..., 1:]], dim=-1)      # no host-side constant: stays graph-capturable
    return conj / torch.sum(quat * quat, dim=-1, keepdim=True)


# --------------------------------------------------------------------------------------------------
# Rotation
# --------------------------------------------------------------------------------------------------
class Rotation:
    """A tensor-like batch of 3D rotations held either as matrices or as quaternions. ref :289-850."""

    def __init__(self, rot_mats: Optional[torch.Tensor] = None, quats: Optional[torch.Tensor] = None,
                 normalize_quats: bool = True):
        if (rot_mats is None) == (quats is None):
            raise ValueError("Exactly one input argument must be specified")
        if (rot_mats is not None and rot_mats.shape[-2:] != (3, 3)) or \
                (quats is not None and quats.shape[-1] != 4):
            raise ValueError("Incorrectly shaped rotation matrix or quaternion")
        # full precision always (ref :326-329)
        if quats is not None:
            quats = quats.type(torch.float32)
            if normalize_quats:
                quats = quats / torch.linalg.norm(quats, dim=-1, keepdim=True)
        else:
            rot_mats = rot_mats.type(torch.float32)
        self._rot_mats = rot_mats
        self._quats = quats

    # -- construction ------------------------------------------------------------------------------
    @staticmethod
    def identity(shape, dtype=None, device=None, requires_grad: bool = True, fmt: str = "quat"):
        if fmt == "rot_mat":
            return Rotation(rot_mats=identity_rot_mats(shape, dtype, device, requires_grad))
        if fmt == "quat":
            return Rotation(quats=identity_quats(shape, dtype, device, requires_grad), normalize_quats=False)
        raise ValueError(f"Invalid format: f{fmt}")

    def _like(self, rot_mats=None, quats=None):
        if rot_mats is not None:
            return Rotation(rot_mats=rot_mats)
        return Rotation(quats=quats, normalize_quats=False)

    def _map(self, fn_mat: Callable, fn_quat: Callable):
        if self._rot_mats is not None:
            return Rotation(rot_mats=fn_mat(self._rot_mats))
        if self._quats is not None:
            return Rotation(quats=fn_quat(self._quats), normalize_quats=False)
        raise ValueError("Both rotations are None")

    # -- tensor protocol ---------------------------------------------------------------------------
    def __getitem__(self, index: Any):
        if type(index) != tuple:
            index = (index,)
        return self._map(lambda m: m[index + (slice(None), slice(None))],
                         lambda q: q[index + (slice(None),)])

    def __mul__(self, right: torch.Tensor):
        if not isinstance(right, torch.Tensor):
            raise TypeError("The other multiplicand must be a Tensor")
        return self._map(lambda m: m * right[..., None, None], lambda q: q * right[..., None])

    def __rmul__(self, left: torch.Tensor):
        return self.__mul__(left)

    @property
    def shape(self) -> torch.Size:
        return self._quats.shape[:-1] if self._quats is not None else self._rot_mats.shape[:-2]

    @property
    def dtype(self) -> torch.dtype:
        return self.get_cur_rot().dtype

    @property
    def device(self) -> torch.device:
        return self.get_cur_rot().device

    @property
    def requires_grad(self) -> bool:
        return self.get_cur_rot().requires_grad

    def get_rot_mats(self) -> torch.Tensor:
        if self._rot_mats is not None:
            return self._rot_mats
        if self._quats is None:
            raise ValueError("Both rotations are None")
        return quat_to_rot(self._quats)

    def get_quats(self) -> torch.Tensor:
        if self._quats is not None:
            return self._quats
        if self._rot_mats is None:
            raise ValueError("Both rotations are None")
        return rot_to_quat(self._rot_mats)

    def get_cur_rot(self) -> torch.Tensor:
        if self._rot_mats is not None:
            return self._rot_mats
        if self._quats is not None:
            return self._quats
        raise ValueError("Both rotations are None")

    def get_rotvec(self, eps=1e-6) -> torch.Tensor:
        """Axis-angle vector, scipy convention (w forced >= 0). ref :556-583."""
        quat = self.get_quats()
        quat = torch.where(quat[..., :1] < 0, -quat, quat)
        angle = 2 * torch.atan2(torch.linalg.norm(quat[..., 1:], dim=-1), quat[..., 0])
        a2 = angle * angle
        small = 2 + a2 / 12 + 7 * a2 * a2 / 2880
        large = angle / torch.sin(angle / 2 + eps)
        is_small = (angle <= 1e-3).float()
        scale = small * is_small + (1 - is_small) * large
        return scale[..., None] * quat[..., 1:]

    # -- algebra -----------------------------------------------------------------------------------
    def compose_q_update_vec(self, q_update_vec: torch.Tensor, normalize_quats: bool = True,
                             update_mask: torch.Tensor = None):
        """q <- normalise(q + mask * q*(0,x,y,z)). ref :587-616."""
        quats = self.get_quats()
        upd = quat_multiply_by_vec(quats, q_update_vec)
        if update_mask is not None:
            upd = upd * update_mask
        return Rotation(quats=quats + upd, normalize_quats=normalize_quats)

    def compose_r(self, r):
        return Rotation(rot_mats=rot_matmul(self.get_rot_mats(), r.get_rot_mats()))

    def compose_q(self, r, normalize_quats: bool = True):
        return Rotation(quats=quat_multiply(self.get_quats(), r.get_quats()), normalize_quats=normalize_quats)

    def apply(self, pts: torch.Tensor) -> torch.Tensor:
        return rot_vec_mul(self.get_rot_mats(), pts)

    def invert_apply(self, pts: torch.Tensor) -> torch.Tensor:
        return rot_vec_mul(invert_rot_mat(self.get_rot_mats()), pts)

    def invert(self):
        return self._map(invert_rot_mat, invert_quat)

    def unsqueeze(self, dim: int):
        if dim >= len(self.shape):
            raise ValueError("Invalid dimension")
        return self._map(lambda m: m.unsqueeze(dim if dim >= 0 else dim - 2),
                         lambda q: q.unsqueeze(dim if dim >= 0 else dim - 1))

    @staticmethod
    def cat(rs: Sequence["Rotation"], dim: int):
        """Concatenate along a batch dim; the result is always in matrix form. ref :730-754."""
        mats = torch.cat([r.get_rot_mats() for r in rs], dim=dim if dim >= 0 else dim - 2)
        return Rotation(rot_mats=mats)

    def map_tensor_fn(self, fn):
        """Map a Tensor->Tensor fn over each of the 9 (or 4) component planes. ref :756-783."""
        def on_mat(m):
            flat = m.reshape(m.shape[:-2] + (9,))
            out = torch.stack([fn(c) for c in flat.unbind(-1)], dim=-1)
            return out.view(out.shape[:-1] + (3, 3))
        return self._map(on_mat, lambda q: torch.stack([fn(c) for c in q.unbind(-1)], dim=-1))

    def cuda(self):
        return self._map(lambda m: m.cuda(), lambda q: q.cuda())

    def to(self, device: Optional[torch.device], dtype: Optional[torch.dtype]):
        return self._map(lambda m: m.to(device=device, dtype=dtype), lambda q: q.to(device=device, dtype=dtype))

    def detach(self):
        return self._map(lambda m: m.detach(), lambda q: q.detach())


# --------------------------------------------------------------------------------------------------
# Rigid
# --------------------------------------------------------------------------------------------------
class Rigid:
    """Rotation + translation with shared batch dims, tensor-like. ref :853-1448."""

    def __init__(self, rots: Optional[Rotation], trans: Optional[torch.Tensor]):
        if trans is not None:
            batch_dims, dtype, device, rg = trans.shape[:-1], trans.dtype, trans.device, trans.requires_grad
        elif rots is not None:
            batch_dims, dtype, device, rg = rots.shape, rots.dtype, rots.device, rots.requires_grad
        else:
            raise ValueError("At least one input argument must be specified")
        if rots is None:
            rots = Rotation.identity(batch_dims, dtype, device, rg)
        elif trans is None:
            trans = identity_trans(batch_dims, dtype, device, rg)
        if (rots.shape != trans.shape[:-1]) or (rots.device != trans.device):
            raise ValueError("Rots and trans incompatible")
        self._rots = rots
        self._trans = trans.type(torch.float32)

    @staticmethod
    def identity(shape, dtype=None, device=None, requires_grad: bool = True, fmt: str = "quat"):
        return Rigid(Rotation.identity(shape, dtype, device, requires_grad, fmt=fmt),
                     identity_trans(shape, dtype, device, requires_grad))

    def __getitem__(self, index: Any):
        if type(index) != tuple:
            index = (index,)
        return Rigid(self._rots[index], self._trans[index + (slice(None),)])

    def __mul__(self, right: torch.Tensor):
        if not isinstance(right, torch.Tensor):
            raise TypeError("The other multiplicand must be a Tensor")
        return Rigid(self._rots * right, self._trans * right[..., None])

    def __rmul__(self, left: torch.Tensor):
        return self.__mul__(left)

    @property
    def shape(self) -> torch.Size:
        return self._trans.shape[:-1]

    @property
    def device(self) -> torch.device:
        return self._trans.device

    def get_rots(self) -> Rotation:
        return self._rots

    def get_trans(self) -> torch.Tensor:
        return self._trans

    # -- the hot ones ------------------------------------------------------------------------------
    def compose_q_update_vec(self, q_update_vec: torch.Tensor, update_mask: torch.Tensor = None):
        """Backbone update (Alg. 23): quaternion update + rotated translation update. ref :1039-1063."""
        quats = self._rots._quats
        if (quats is not None and quats.is_cuda and q_update_vec.is_cuda
                and quats.shape == q_update_vec.shape[:-1] + (4,)
                and (update_mask is None or update_mask.shape == q_update_vec.shape[:-1] + (1,))):
            new_q, new_t = _k.compose_q_update(quats, self._trans, q_update_vec, update_mask)
            return Rigid(Rotation(quats=new_q, normalize_quats=False), new_t)
        q_vec, t_vec = q_update_vec[..., :3], q_update_vec[..., 3:]
        new_rots = self._rots.compose_q_update_vec(q_vec, update_mask=update_mask)
        trans_update = self._rots.apply(t_vec)
        if update_mask is not None:
            trans_update = trans_update * update_mask
        return Rigid(new_rots, self._trans + trans_update)

    def apply(self, pts: torch.Tensor) -> torch.Tensor:
        """R p + t. ref :1104-1116."""
        quats = self._rots._quats
        if quats is not None and quats.is_cuda and pts.is_cuda and pts.dtype == torch.float32:
            return _k.rigid_apply(quats, self._trans, pts, inverse=False)
        if quats is None and pts.is_cuda and pts.dtype == torch.float32:
            return _k.rot_compose(self._rots.get_rot_mats(), self._trans, None, pts)[1]
        return self._rots.apply(pts) + self._trans

    def invert_apply(self, pts: torch.Tensor) -> torch.Tensor:
        """R^T (p - t). ref :1118-1130."""
        quats = self._rots._quats
        if quats is not None and quats.is_cuda and pts.is_cuda and pts.dtype == torch.float32:
            return _k.rigid_apply(quats, self._trans, pts, inverse=True)
        if quats is None and pts.is_cuda and pts.dtype == torch.float32:
            return _k.rot_compose(self._rots.get_rot_mats(), self._trans, None, pts, inverse=True)[1]
        return self._rots.invert_apply(pts - self._trans)

    # -- the rest ----------------------------------------------------------------------------------
    def compose(self, r: "Rigid"):
        """self o r: (R1 R2, R1 t2 + t1), ref :1065-1079 (one kernel on CUDA)."""
        if self._trans.is_cuda and r._trans.is_cuda:
            rot, trans = _k.rot_compose(self._rots.get_rot_mats(), self._trans, r._rots.get_rot_mats(), r._trans)
            return Rigid(Rotation(rot_mats=rot), trans)
        return Rigid(self._rots.compose_r(r._rots), self._rots.apply(r._trans) + self._trans)

    def compose_r(self, rot: Rotation, order: str = "right"):
        if order == "right":
            new_rot = self._rots.compose_r(rot)
        elif order == "left":
            new_rot = rot.compose_r(self._rots)
        else:
            raise ValueError(f"Unrecognized multiplication order: {order}")
        return Rigid(new_rot, self._trans)

    def invert(self):
        rot_inv = self._rots.invert()
        return Rigid(rot_inv, -1 * rot_inv.apply(self._trans))

    def map_tensor_fn(self, fn):
        new_trans = torch.stack([fn(c) for c in self._trans.unbind(-1)], dim=-1)
        return Rigid(self._rots.map_tensor_fn(fn), new_trans)

    def to_tensor_4x4(self) -> torch.Tensor:
        out = self._trans.new_zeros((*self.shape, 4, 4))
        out[..., :3, :3] = self._rots.get_rot_mats()
        out[..., :3, 3] = self._trans
        out[..., 3, 3] = 1
        return out

    @staticmethod
    def from_tensor_4x4(t: torch.Tensor):
        if t.shape[-2:] != (4, 4):
            raise ValueError("Incorrectly shaped input tensor")
        return Rigid(Rotation(rot_mats=t[..., :3, :3]), t[..., :3, 3])

    def to_tensor_7(self) -> torch.Tensor:
        return torch.cat([self._rots.get_quats(), self._trans], dim=-1)

    @staticmethod
    def from_tensor_7(t: torch.Tensor, normalize_quats: bool = False):
        if t.shape[-1] != 7:
            raise ValueError("Incorrectly shaped input tensor")
        return Rigid(Rotation(quats=t[..., :4], normalize_quats=normalize_quats), t[..., 4:])

    @staticmethod
    def from_3_points(p_neg_x_axis: torch.Tensor, origin: torch.Tensor, p_xy_plane: torch.Tensor,
                      eps: float = 1e-8):
        """Gram-Schmidt frame from three points (Alg. 21). ref :1232-1275."""
        e0 = origin - p_neg_x_axis
        e1 = p_xy_plane - origin
        e0 = e0 / torch.sqrt(torch.sum(e0 * e0, dim=-1, keepdim=True) + eps)
        e1 = e1 - e0 * torch.sum(e0 * e1, dim=-1, keepdim=True)
        e1 = e1 / torch.sqrt(torch.sum(e1 * e1, dim=-1, keepdim=True) + eps)
        e2 = torch.cross(e0, e1, dim=-1)
        rots = torch.stack([e0, e1, e2], dim=-1)  # columns are the axes
        return Rigid(Rotation(rot_mats=rots), origin)

    def unsqueeze(self, dim: int):
        if dim >= len(self.shape):
            raise ValueError("Invalid dimension")
        return Rigid(self._rots.unsqueeze(dim), self._trans.unsqueeze(dim if dim >= 0 else dim - 1))

    @staticmethod
    def cat(ts: Sequence["Rigid"], dim: int):
        rots = Rotation.cat([t._rots for t in ts], dim)
        trans = torch.cat([t._trans for t in ts], dim=dim if dim >= 0 else dim - 1)
        return Rigid(rots, trans)

    def apply_rot_fn(self, fn):
        return Rigid(fn(self._rots), self._trans)

    def apply_trans_fn(self, fn):
        return Rigid(self._rots, fn(self._trans))

    def scale_translation(self, trans_scale_factor: float):
        return self.apply_trans_fn(lambda t: t * trans_scale_factor)

    def stop_rot_gradient(self):
        return self.apply_rot_fn(lambda r: r.detach())

    @staticmethod
    def make_transform_from_reference(n_xyz, ca_xyz, c_xyz, eps=1e-20):
        """Frame that maps the ideal backbone onto (N, CA, C). ref :1368-1439.

        Reproduces the reference's arithmetic exactly, including its quirk of writing the second
        rotation's third row into the *first* rotation (:1414-1415), so results match bit-for-bit
        in structure.
        """
        translation = -1 * ca_xyz
        n_xyz = n_xyz + translation
        c_xyz = c_xyz + translation
        c_x, c_y, c_z = c_xyz[..., 0], c_xyz[..., 1], c_xyz[..., 2]
        zeros, ones = torch.zeros_like(c_x), torch.ones_like(c_x)

        norm = torch.sqrt(eps + c_x ** 2 + c_y ** 2)
        sin_c1, cos_c1 = -c_y / norm, c_x / norm
        norm = torch.sqrt(eps + c_x ** 2 + c_y ** 2 + c_z ** 2)
        sin_c2, cos_c2 = c_z / norm, torch.sqrt(c_x ** 2 + c_y ** 2) / norm

        def mat(rows):
            return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)

        c1_rots = mat([[cos_c1, -sin_c1, zeros], [sin_c1, cos_c1, zeros], [-sin_c2, zeros, cos_c2]])
        c2_rots = mat([[cos_c2, zeros, sin_c2], [zeros, ones, zeros], [zeros, zeros, zeros]])
        c_rots = rot_matmul(c2_rots, c1_rots)
        n_xyz = rot_vec_mul(c_rots, n_xyz)
        n_y, n_z = n_xyz[..., 1], n_xyz[..., 2]
        norm = torch.sqrt(eps + n_y ** 2 + n_z ** 2)
        sin_n, cos_n = -n_z / norm, n_y / norm
        n_rots = mat([[ones, zeros, zeros], [zeros, cos_n, -sin_n], [zeros, sin_n, cos_n]])
        rots = rot_matmul(n_rots, c_rots).transpose(-1, -2)
        return Rigid(Rotation(rot_mats=rots), -1 * translation)

    def cuda(self):
        return Rigid(self._rots.cuda(), self._trans.cuda())
