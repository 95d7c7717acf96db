"""The drop-in boundary exercised by the REFERENCE'S OWN CALLER: the unmodified
``Experiment.loss_fn`` and ``Experiment.inference_fn`` (/root/reference/train_DFOLD_dynamics.py:1181, :1425) run on a
synthetic batch once over the reference's FullScoreNetwork and once over this package's FullScoreNetwork (operators routed
to the oracle so the host logic runs on CPU), with the reference's own SE3Diffuser in both cases.  Needs the reference
checkout (skipped on the GPU box, where it does not exist)."""
import copy
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from dynamicpdb_b200 import synthetic as syn
from oracle import dfold_oracle as O
from oracle import ref_shims
from tests.test_cpu_oracle import close

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="needs the reference checkout")

NF, N = 3, 12


def _models():
    ref_shims.install_trainer_stubs()
    from src.data import se3_diffuser
    from src.model import Dfold_network_dynamic as RefNet
    from dynamicpdb_b200.Dfold_network_dynamic import FullScoreNetwork
    conf = syn.model_conf(NF, **syn.PRESET_TINY)
    diffuser = se3_diffuser.SE3Diffuser(syn.diffuser_conf(1.0))
    torch.manual_seed(0)
    ref = RefNet.FullScoreNetwork(conf, diffuser)
    sd = ref.state_dict()
    syn.dezero_(sd)
    ref.load_state_dict(sd)
    ours = FullScoreNetwork(conf, diffuser)
    ours.load_state_dict(copy.deepcopy(sd), strict=True)
    return ref, ours, diffuser


def _experiment(model, diffuser):
    model_conf = SimpleNamespace(cfg_drop_in_train=True, cfg_drop_rate=0.0, embed=SimpleNamespace(embed_self_conditioning=False))
    exp_conf = O.default_exp_conf(bb_atom_loss_weight=1.0, bb_atom_loss_t_filter=0.25, dist_mat_loss_weight=1.0,
                                  dist_mat_loss_t_filter=0.25, aux_loss_weight=0.25)
    exp = ref_shims.reference_experiment(model, diffuser, model_conf, exp_conf, SimpleNamespace(diffuse_rot=True))
    exp._conf = SimpleNamespace(model=SimpleNamespace(cfg_drop_rate=0.0, cfg_gamma=1.0))
    exp._data_conf = SimpleNamespace(num_t=4, min_t=0.01)
    return exp


def _batch():
    feats = syn.add_loss_targets(syn.make_feats(NF, N, seed=21, loader_dtypes=True), seed=21)
    feats["fixed_mask"][:, 2] = 1
    return feats


def test_reference_loss_fn_runs_over_the_overlay_model(oracle_ops):
    """Same loss, aux_data and parameter gradients whether Experiment.loss_fn drives the reference network or ours."""
    ref, ours, diffuser = _models()
    res = []
    for net in (ref, ours):
        net.zero_grad(set_to_none=True)
        loss, aux = _experiment(net, diffuser).loss_fn(_batch())
        loss.backward()
        res.append((loss.detach(), aux, {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
    (l0, a0, g0), (l1, a1, g1) = res
    assert l0.dtype == l1.dtype and close(l1, l0, 1e-5)
    assert set(a0) == set(a1)
    for k in a0:
        assert close(a1[k], a0[k], 1e-5), k
    assert set(g0) == set(g1) and len(g0) > 50
    for k in g0:
        assert close(g1[k], g0[k], 2e-4), k
    # and the device-side restatement of the same loss agrees with the caller's own
    out = ours(_batch())
    l2, _ = O.loss_fn(out, _batch(), O.default_exp_conf())
    assert close(l2.detach(), l0, 1e-5)


def test_reference_inference_fn_runs_over_the_overlay_model(oracle_ops):
    """Experiment.inference_fn (reverse diffusion with the reference's numpy / scipy SE3Diffuser.reverse) over both networks
    with the same numpy random stream: identical trajectories."""
    ref, ours, diffuser = _models()
    outs = []
    for net in (ref, ours):
        np.random.seed(3)
        torch.manual_seed(3)
        feats = {k: v for k, v in syn.make_feats(NF, N, seed=22, loader_dtypes=True).items()}
        res = _experiment(net, diffuser).inference_fn(feats, num_t=4, min_t=0.01, aux_traj=True, noise_scale=0.5)
        outs.append(res)
    assert set(outs[0]) == set(outs[1])
    for k in outs[0]:
        a, b = np.asarray(outs[0][k]), np.asarray(outs[1][k])
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() <= 2e-4 * max(1.0, np.abs(a).max()), k
