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
