"""Host-side mirror of ``src/model/Dfold_network_dynamic.py``: ``FullScoreNetwork`` with the
reference's constructor, ``forward(input_feats, drop_ref=False) -> dict`` contract, output keys and ``state_dict``
layout (``embedding_layer.*``, ``score_model.*``, ``expand_node``, ``expand_edge``).

Differences that leave every returned value identical:
  * the ``DFOLDv2_Embeder`` call (ref :478-484) produces tensors nothing consumes (it materialises
    the frame-expanded edge tensor three times, ~10 GB of traffic at nf=64); its parameters are kept (checkpoint
    ABI, ``find_unused_parameters=True`` as in the reference trainer) but the dead computation is not run unless
    ``DFOLD_RUN_DEAD_EMBEDDER=1``;
  * residue tables are device-resident, so the epilogue has no host synchronisation.
"""
import functools as fn
import os
from typing import Tuple

import torch
from torch import nn

from . import feats
from . import ipa_pytorch_dynamic
from . import kernels as K
from .feats import atom14_to_atom37  # noqa: F401  (public name in the reference module)

Tensor = torch.Tensor


def get_timestep_embedding(timesteps, embedding_dim, max_positions=10000):
    """src/model/utils.py:46-58."""
    import math
    assert len(timesteps.shape) == 1
    timesteps = timesteps * max_positions
    half_dim = embedding_dim // 2
    emb = math.log(max_positions) / (half_dim - 1)
    emb = torch.exp(torch.arange(half_dim, dtype=torch.float32, device=timesteps.device) * -emb)
    emb = timesteps.float()[:, None] * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1), mode="constant")
    return emb


class DFOLDv2_Embeder(nn.Module):
    """ref :19-88.  Owns checkpoint parameters; its outputs are not consumed by the trunk."""

    def __init__(self, model_conf):
        super().__init__()
        self._model_conf = model_conf
        self._embed_conf = model_conf.embed
        node_embed_size = model_conf.node_embed_size
        edge_embed_size = model_conf.edge_embed_size
        time_embed_size = node_embed_size
        self.timestep_embed = fn.partial(get_timestep_embedding, embedding_dim=time_embed_size)
        self.node_timestep_proj = nn.Sequential(nn.Linear(time_embed_size, node_embed_size // 2), nn.SiLU(),
                                                nn.Linear(node_embed_size // 2, node_embed_size))
        self.node_ln = nn.LayerNorm(node_embed_size)
        self.edge_timestep_proj = nn.Sequential(nn.Linear(time_embed_size, edge_embed_size // 2), nn.SiLU(),
                                                nn.Linear(edge_embed_size // 2, edge_embed_size))
        self.edge_ln = nn.LayerNorm(edge_embed_size)

    def forward(self, node_repr, edge_repr, seq_idx, t):
        num_batch, num_res = seq_idx.shape
        t_embed = self.timestep_embed(t)
        node_embed = K.layer_norm(node_repr, self.node_ln.weight, self.node_ln.bias, self.node_ln.eps)
        edge_embed = K.layer_norm(edge_repr.reshape(num_batch, num_res * num_res, -1),
                                  self.edge_ln.weight, self.edge_ln.bias, self.edge_ln.eps)
        ref_edge = edge_embed[0].reshape(num_res, num_res, -1)
        return node_embed, edge_embed.reshape(num_batch, num_res, num_res, -1), node_embed[0], ref_edge, t_embed


class FullScoreNetwork(nn.Module):
    """ref :429-569."""

    def __init__(self, model_conf, diffuser):
        super().__init__()
        self._model_conf = model_conf
        self.embedding_layer = DFOLDv2_Embeder(model_conf)
        self.diffuser = diffuser
        self.score_model = ipa_pytorch_dynamic.DFOLDIpaScore(model_conf, diffuser)
        self.expand_node = nn.Linear(256, model_conf.node_embed_size)
        self.expand_edge = nn.Linear(128, model_conf.edge_embed_size)

    def _apply_mask(self, aatype_diff, aatype_0, diff_mask):
        return diff_mask * aatype_diff + (1 - diff_mask) * aatype_0

    def forward(self, input_feats, drop_ref=False):
        fixed_mask = input_feats["fixed_mask"].type(torch.float32)
        num_res = input_feats["node_repr"].shape[0]

        input_feats["expand_node_repr"] = K.linear(input_feats["node_repr"], self.expand_node.weight,
                                                   self.expand_node.bias)
        input_feats["expand_edge_repr"] = K.linear(
            input_feats["edge_repr"].reshape(num_res * num_res, -1), self.expand_edge.weight,
            self.expand_edge.bias).reshape(num_res, num_res, -1)

        if os.environ.get("DFOLD_RUN_DEAD_EMBEDDER") == "1":
            ft = self._model_conf.frame_time
            _, _, ref_node, ref_edge, t_embed = self.embedding_layer(
                node_repr=input_feats["expand_node_repr"].unsqueeze(0).expand(ft, -1, -1),
                edge_repr=input_feats["expand_edge_repr"].unsqueeze(0).expand(ft, -1, -1, -1),
                seq_idx=input_feats["seq_idx"], t=input_feats["t"])
            input_feats.update({"ref_node_repr": ref_node, "ref_edge_repr": ref_edge, "t_embed": t_embed})

        model_out = self.score_model(None, None, input_feats, drop_ref=drop_ref)

        gt_angles = input_feats["torsion_angles_sin_cos"]
        dm = 1 - fixed_mask[..., None, None]
        angles_pred = self._apply_mask(model_out["angles"], gt_angles, dm)
        unorm_angles = self._apply_mask(model_out["unorm_angles"], gt_angles, dm)
        pred_out = {
            "angles": angles_pred,
            "unorm_angles": unorm_angles,
            "rot_score": model_out["rot_score"],
            "trans_score": model_out["trans_score"],
        }
        rigids_pred = model_out["final_rigids"]
        pred_out["rigids"] = rigids_pred.to_tensor_7()
        # torsion frames -> atom14 -> atom37 (ref :532-538) in one kernel
        atom14_pos, atom37_pos = feats.frames_to_atoms(rigids_pred, angles_pred, input_feats["aatype"])
        pred_out["atom37"] = atom37_pos
        pred_out["atom14"] = atom14_pos
        pred_out["rigid_update"] = model_out["rigid_update"]
        return pred_out

    def debug_foward(self, input_feats, drop_ref=False):
        """ref :549-569: forward + the set of parameters whose module ran."""
        used_params = set()

        def hook(module, _inp, _out):
            for param in module.parameters():
                used_params.add(param)

        hooks = [m.register_forward_hook(hook) for m in self.modules()]
        output = self.forward(input_feats, drop_ref)
        for h in hooks:
            h.remove()
        return output, used_params
