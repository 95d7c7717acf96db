"""Reverse-diffusion inference helper (BASELINE.json configs[1]).

In DFOLDv2 the trunk never sees the diffusion time or the noised frames: ``DFOLDIpaScore.forward`` uses ``t`` and
``rigids_t`` only in the closed-form score formulas at the very end (reference src/model/ipa_pytorch_dynamic.py:883-897;
SURVEY.md §0-4, §3.2).  Across the ``num_t`` steps of ``Experiment.inference_fn`` (train_DFOLD_dynamics.py:1469-1502) the
predicted frames, angles and atoms are therefore identical; only ``rot_score`` / ``trans_score`` change.

``MemoizedScoreNetwork`` wraps a ``FullScoreNetwork`` and returns, for every call whose trunk inputs are unchanged,
the cached trunk outputs plus freshly evaluated scores — bit-identical to calling the network again, at the cost of
the score epilogue only.  The reference's sampling loop can use it as a drop-in ``model`` object.
"""
from typing import Dict, Optional

import torch

from .rigid_utils import Rigid

_TRUNK_KEYS = ("res_mask", "fixed_mask", "seq_idx", "rigids_0", "force", "vel", "node_repr", "edge_repr",
               "torsion_angles_sin_cos", "torsion_angles_mask", "aatype")


class MemoizedScoreNetwork(torch.nn.Module):
    def __init__(self, net):
        super().__init__()
        self.net = net
        self._key = None
        self._cached: Optional[Dict[str, torch.Tensor]] = None
        self._held = None

    def _signature(self, feats):
        """Identity + version of every trunk input, and the versions of the weights.  The keyed tensors themselves are
        kept alive in ``self._held`` for as long as the entry lives: the reference deep-copies the window before every
        sampling run (train_DFOLD_dynamics.py:1442), and a freed block is handed out again at the same address with
        version 0 — (data_ptr, version, shape) alone would then collide with different contents."""
        ins = tuple((k, id(feats[k]), feats[k]._version) for k in _TRUNK_KEYS if k in feats)
        wts = tuple((id(p), p._version) for p in self.net.parameters())
        return ins, wts

    def reset(self):
        """Drop the cached trunk outputs (call after weight updates made through ``param.data``, which bump no version)."""
        self._key, self._cached, self._held = None, None, None

    @torch.no_grad()
    def forward(self, input_feats, drop_ref=False):
        sig = self._signature(input_feats)
        if self._cached is None or sig != self._key:
            self._cached = self.net(input_feats, drop_ref=drop_ref)
            self._key = sig
            self._held = [input_feats[k] for k in _TRUNK_KEYS if k in input_feats]
            return dict(self._cached)
        out = dict(self._cached)
        diffuser = self.net.diffuser
        node_mask = input_feats["res_mask"].type(torch.float32)
        init = Rigid.from_tensor_7(input_feats["rigids_t"].type(torch.float32))
        pred = Rigid.from_tensor_7(out["rigids"])             # already unscaled, as the network returns it
        t = input_feats["t"]
        # the network evaluates the rotation score on the frames before unscaling; rotations are unaffected by it
        out["rot_score"] = diffuser.calc_rot_score(init.get_rots(), pred.get_rots(), t) * node_mask[..., None]
        out["trans_score"] = diffuser.calc_trans_score(init.get_trans(), pred.get_trans(), t[:, None, None],
                                                       use_torch=True) * node_mask[..., None]
        return out


class DeviceReverseDiffusion:
    """The sampling loop of ``Experiment.inference_fn`` (reference train_DFOLD_dynamics.py:1425-1530) kept on the device:

        one trunk evaluation (memoised: the trunk does not depend on t or on the noised frames)
        + per reverse step  ONE score-epilogue kernel  +  ONE reverse-step kernel

    instead of ``num_t`` full network evaluations each followed by a numpy / scipy reverse step on the host.
    BASELINE.json configs[1]: 100 steps, N_res = 256, 32 frames."""

    def __init__(self, net):
        from .score_epilogue import SE3ScoreDiffuser
        self.net = net
        self.memo = MemoizedScoreNetwork(net)
        sd = getattr(net.score_model, "_fused_scores", None)
        if not isinstance(sd, SE3ScoreDiffuser):
            raise ValueError("DeviceReverseDiffusion needs a diffuser with the reference's logarithmic / VP-SDE schedules")
        self.diffuser = sd
        self._graphs: Dict = {}

    @torch.no_grad()
    def sample_graphed(self, data_init: Dict[str, torch.Tensor], num_t: int, min_t: float, noise_scale: float = 1.0,
                       center: bool = True, noise=None):
        """``sample`` (memoised) replayed from ONE CUDA graph holding the trunk pass and all ``num_t`` score / reverse steps:
        the per-step work is two small kernels plus a few tensor views, so launched eagerly the loop is bound by Python and
        launch latency, not by the device.  The graph is captured on the first call for a given (schedule, shapes) and reads
        its inputs from static buffers; without injected ``noise`` the normal draws come from the default CUDA generator
        (graph-aware: every replay draws fresh numbers)."""
        key = (num_t, float(min_t), float(noise_scale), bool(center), noise is not None,
               tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(data_init.items())))
        ent = self._graphs.get(key)
        if ent is None:
            static = {k: v.clone() for k, v in data_init.items()}
            static_noise = None if noise is None else (noise[0].clone(), noise[1].clone())
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):                     # warm-up: allocator pools, library load, plane caches
                self.memo.reset()
                self.sample(static, num_t, min_t, noise_scale, center, static_noise)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            self.memo.reset()                                 # the trunk pass must be part of the captured work
            from . import kernels as K
            K.invalidate_weight_cache()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                res = self.sample(static, num_t, min_t, noise_scale, center, static_noise)
            self.memo.reset()                                 # its cache now points into the graph's private pool
            K.invalidate_weight_cache()
            ent = (g, static, static_noise, res)
            self._graphs[key] = ent
        g, static, static_noise, res = ent
        for k, v in data_init.items():
            static[k].copy_(v)
        if noise is not None:
            static_noise[0].copy_(noise[0])
            static_noise[1].copy_(noise[1])
        g.replay()
        return {k: v.clone() for k, v in res.items()}

    @torch.no_grad()
    def sample(self, data_init: Dict[str, torch.Tensor], num_t: int, min_t: float, noise_scale: float = 1.0, center: bool = True,
               noise=None, generator=None, literal: bool = False):
        """-> {"prot_traj": [num_t, nf, N, 37, 3] atom37 per step (t = min_t first, as the reference flips it),
               "rigids": final noised-frame tensor [nf, N, 7], "rigid_pred": [nf, N, 7]}.
        ``noise``: optional (z_rot, z_trans) pair of [num_t, nf, N, 3] tensors; ``literal=True`` evaluates the whole network
        at every step as the reference does (for the comparison in tests / bench)."""
        import numpy as np
        feats = dict(data_init)
        dev = feats["rigids_t"].device
        reverse_steps = np.linspace(min_t, 1.0, num_t)[::-1]
        dt = 1.0 / num_t
        model = self.net if literal else self.memo
        traj = []
        out = None
        for i, t in enumerate(reverse_steps):
            feats["t"] = torch.full((1,), float(t), device=dev, dtype=feats["t"].dtype if "t" in feats else torch.float32)
            out = model(feats)
            if t > min_t:
                diffuse_mask = (1 - feats["fixed_mask"].float()) * feats["res_mask"].float()
                z = (None, None) if noise is None else (noise[0][i], noise[1][i])
                rig = self.diffuser.reverse(Rigid.from_tensor_7(feats["rigids_t"].float()), out["rot_score"], out["trans_score"],
                                            float(t), dt, diffuse_mask=diffuse_mask, center=center, noise_scale=noise_scale,
                                            z_rot=z[0], z_trans=z[1], generator=generator)
                feats["rigids_t"] = rig.to_tensor_7()
            else:
                feats["rigids_t"] = out["rigids"]
            feats["sc_ca_t"] = out["rigids"][..., 4:]
            traj.append(out["atom37"])
        return {"prot_traj": torch.stack(traj[::-1]), "rigids": feats["rigids_t"], "rigid_pred": out["rigids"]}
