"""Plain-torch (CPU-capable) statements of each operator in dynamicpdb_b200/kernels.py, with the same signatures.

TEST INFRASTRUCTURE: tests monkey-patch ``dynamicpdb_b200.kernels.<op>`` with these to exercise the product's
host-side module logic without a GPU, and the ``-m gpu`` tests compare each CUDA op against them.
"""
import math

import torch
import torch.nn.functional as F

from . import dfold_oracle as O


def linear(x, weight, bias=None, act=None, residual=None, pre_relu=False):
    if pre_relu:
        x = F.relu(x)
    y = F.linear(x.to(weight.dtype), weight, bias)
    if act == "relu":
        y = F.relu(y)
    elif act == "silu":
        y = F.silu(y)
    return y if residual is None else y + residual


def conv5x5(x, weight, bias=None, relu=True, residual=None, crop=0):
    y = F.conv2d(x.permute(2, 0, 1).unsqueeze(0), weight, bias, padding=(weight.shape[2] // 2, weight.shape[3] // 2))
    y = y.squeeze(0).permute(1, 2, 0)[crop:]
    if relu:
        y = F.relu(y)
    return y if residual is None else y + residual


def global_layernorm(x, eps=1e-4, silu=False):
    y = O.my_layernorm(x, eps)
    return F.silu(y) if silu else y


def layer_norm(x, weight, bias, eps=1e-5):
    return F.layer_norm(x, x.shape[-1:], weight, bias, eps)


def quat_to_rot(q):
    return O.quat_to_rot(q)


def rigid_apply(quat, trans, pts, inverse=False):
    rig = torch.cat([quat, trans], dim=-1)
    return O.rigid_invert_apply(rig, pts) if inverse else O.rigid_apply(rig, pts)


def ipa_points(raw, quat, trans, H):
    Fs, N, W = raw.shape
    hp = W // 3
    pts = raw.reshape(Fs, N, 3, hp).transpose(-1, -2)
    out = rigid_apply(quat[:, :, None, :], trans[:, :, None, :], pts)
    return out.reshape(out.shape[0], N, H, hp // H, 3)


def compose_q_update(quat, trans, upd6, mask=None):
    r = O.compose_q_update_vec(torch.cat([quat, trans], dim=-1), upd6, mask)
    return r[..., :4], r[..., 4:]


def keep_last_frame(x):
    return torch.cat([torch.zeros_like(x[:-1]), x[-1:]], dim=0)


def qk_logits(q, kv, b_hm, alpha, beta):
    C = q.shape[-1]
    k = kv[..., :C]
    return alpha * torch.einsum("fihc,fjhc->fhij", q, k) + beta * b_hm


def ipa_attention(logit0, kv, q_pts, kv_pts, pair, quat, trans, mask, gamma, *, Pq, Pv, dfold, inf, eps):
    Fn, N, H = q_pts.shape[:3]
    C = kv.shape[-1] // 2
    v = kv[..., C:]
    k_pts, v_pts = kv_pts[..., :Pq, :], kv_pts[..., Pq:, :]
    d = q_pts[:, :, None] - k_pts[:, None, :]                       # [F,i,j,H,Pq,3]
    pt = ((d * d).sum(-1) * gamma[:, None]).sum(-1) * (-0.5)          # [F,i,j,H]
    a = logit0 + pt.permute(0, 3, 1, 2) + (inf * (mask[:, :, None] * mask[:, None, :] - 1))[:, None]
    a = torch.softmax(a, dim=-1)
    o = torch.einsum("fhij,fjhc->fihc", a, v.expand(Fn, -1, -1, -1)).reshape(Fn, N, H * C)
    og = torch.einsum("fhij,fjhpx->fihpx", a, v_pts)
    rig = torch.cat([quat, trans], dim=-1)
    ol = O.rigid_invert_apply(rig[:, :, None, None, :], og)
    nl = torch.sqrt((ol ** 2).sum(-1) + eps).reshape(Fn, N, H * Pv)
    ng = torch.sqrt((og ** 2).sum(-1) + eps).reshape(Fn, N, H * Pv)
    ol, og = ol.reshape(Fn, N, H * Pv, 3), og.reshape(Fn, N, H * Pv, 3)
    pz = pair.expand(Fn, -1, -1, -1)
    o_pair = torch.einsum("fhij,fijc->fihc", a, pz).reshape(Fn, N, -1)
    feats = [o, ol[..., 0], ol[..., 1], ol[..., 2], nl, o_pair]
    if dfold:
        feats += [og[..., 0], og[..., 1], og[..., 2], ng]
    return torch.cat(feats, dim=-1)


def score_epilogue(q_pred, q_t, x_pred, x_t, t, grid, mask, *, max_sigma, min_sigma, min_b, max_b, r3_scale, ipa_scale, L=1000):
    """K9: masked rotation (fp64) and translation scores; x_pred is the translation before unscaling."""
    from types import SimpleNamespace
    dc = SimpleNamespace(num_sigma=grid.numel(), min_sigma=min_sigma, max_sigma=max_sigma, L=L, min_b=min_b, max_b=max_b,
                         coordinate_scaling=r3_scale)
    m = 1.0 if mask is None else mask[..., None]
    rs = O.rot_score(q_t, q_pred, t, dc) * m
    if x_pred is None:
        return rs, None
    tt = t.reshape(-1)[:1][:, None, None] if x_pred.dim() == 3 else t.reshape(-1)[:1]
    ts = O.trans_score(x_t, x_pred / ipa_scale, tt, dc) * m
    return rs, ts


def frames_to_atoms(rot, trans, alpha, aatype, tables, eager, want_frames=False, rot_is_matrix=False):
    """K10: torsion frames -> atom14 -> atom37 (homogeneous 4x4 statement of the oracle)."""
    if rot_is_matrix:
        bb = rot.new_zeros(rot.shape[:-2] + (4, 4))
        bb[..., :3, :3] = rot
        bb[..., :3, 3] = trans
        bb[..., 3, 3] = 1
        fr = O.torsion_angles_to_frames(None, alpha.to(rot.dtype), aatype, bb44=bb)
    else:
        fr = O.torsion_angles_to_frames(torch.cat([rot, trans], dim=-1), alpha.to(rot.dtype), aatype)
    a14 = O.frames_to_atom14(fr, aatype)
    a37, _ = O.atom14_to_atom37(a14, aatype)
    return (a14, a37, fr) if want_frames else (a14, a37)


def reverse_step(q_t, x_t, rot_score, trans_score, z_rot, z_trans, mask, *, g_rot, g_trans, b_t, dt, noise_scale=1.0,
                 r3_scale=1.0, center=True, diffuse_rot=True, diffuse_trans=True):
    """One reverse-diffusion step through the oracle's scipy statement (the schedule constants are re-derived from t there,
    so this seam recovers t from b_t = min_b + t (max_b - min_b) of the default schedule)."""
    dc = O.default_diffuser_conf(r3_scale)
    t = (b_t - dc.min_b) / (dc.max_b - dc.min_b)
    assert diffuse_rot and diffuse_trans
    out = O.reverse_step(torch.cat([q_t, x_t], dim=-1).cpu(), rot_score.cpu().numpy(), trans_score.cpu().numpy(), float(t), float(dt),
                         None if mask is None else mask.cpu().numpy(), z_rot.cpu().numpy(), z_trans.cpu().numpy(), dc,
                         center=center, noise_scale=noise_scale).to(q_t.device)
    return out[..., :4], out[..., 4:]


def quat_mul(a, b, b_is_vec=False):
    if b_is_vec:
        b = torch.cat([torch.zeros_like(b[..., :1]), b], dim=-1)
    return O.quat_mul(a, b)


def rot_compose(Ra, ta, Rb, tb, inverse=False):
    Ro = Ra @ Rb if Rb is not None else None
    to = None
    if tb is not None:
        if inverse:
            d = tb - ta if ta is not None else tb
            to = torch.einsum("...ba,...b->...a", Ra, d)
        else:
            to = torch.einsum("...ab,...b->...a", Ra, tb)
            if ta is not None:
                to = to + ta
    return Ro, to


ALL = ["score_epilogue", "frames_to_atoms", "reverse_step", "quat_mul", "rot_compose", "linear", "conv5x5", "global_layernorm", "layer_norm", "quat_to_rot", "rigid_apply", "ipa_points",
       "compose_q_update", "keep_last_frame", "qk_logits", "ipa_attention"]
