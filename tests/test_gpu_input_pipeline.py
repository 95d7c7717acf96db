"""-m gpu: the featurisation kernel (csrc/featurize.cu) and the prefetching window iterator against the
fixtures generated from the unmodified loader and against the oracle."""
import numpy as np
import pytest
import torch

from dynamicpdb_b200.input_pipeline import TrajectoryStore, WindowPrefetcher, featurize_window
from oracle import dfold_oracle as O
from oracle import synth_traj
from tests.test_cpu_oracle import close, load

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check_against_gold(feats, gold):
    q = feats["rigids_0"][..., :4].double().cpu()
    assert feats["rigids_0"].dtype == torch.float32
    assert close(q.norm(dim=-1), torch.ones(q.shape[:-1], dtype=torch.float64), 1e-6)
    assert close(O.quat_to_rot(q), gold["rot_0"].double(), 2e-6)                    # same rotation (the reference's sign is arbitrary)
    assert close(feats["rigids_0"][..., 4:].cpu(), gold["rigids_0"][..., 4:], 1e-6)
    for k in ("torsion_angles_sin_cos", "alt_torsion_angles_sin_cos"):
        assert feats[k].dtype == gold[k].dtype and close(feats[k].cpu(), gold[k], 1e-6), k
    assert torch.equal(feats["torsion_angles_mask"].cpu(), gold["torsion_angles_mask"])
    assert close(feats["res_mask"].cpu().double(), gold["res_mask"].double(), 0)


def test_featurize_kernel_matches_reference_loader(tmp_path):
    traj = synth_traj.make_trajectory()
    row = synth_traj.write_files(traj, str(tmp_path))
    store = TrajectoryStore(frame_time=3, frame_sample_step=2, keep_first=8)
    ent = store.protein(*row)
    np.random.seed(5)
    sl = store.window_index(ent["atom37"].shape[0], training=True)
    feats = featurize_window(ent["atom37"][sl].to(DEV), ent["atom_mask"].to(DEV), ent["aatype"].to(DEV))
    _check_against_gold(feats, load("loader"))
    with pytest.raises(RuntimeError):
        featurize_window(ent["atom37"][sl], ent["atom_mask"], ent["aatype"])        # no CPU path


@pytest.mark.timeout(120)
def test_prefetcher_yields_reference_windows(tmp_path):
    """Three windows of the same protein through the worker thread + side stream; the first equals the golden window."""
    traj = synth_traj.make_trajectory()
    row = synth_traj.write_files(traj, str(tmp_path))
    store = TrajectoryStore(frame_time=3, frame_sample_step=2, keep_first=8)
    np.random.seed(5)
    items = list(WindowPrefetcher(store, [row, row, row], DEV, training=True))
    assert len(items) == 3
    gold = load("loader")
    _check_against_gold(items[0], gold)
    assert close(items[0]["force"].cpu(), gold["force"], 0) and items[0]["force"].dtype == torch.float64
    for it in items:
        assert it["node_repr"].shape == (17, 256) and it["edge_repr"].shape == (17, 17, 128) and it["aatype"].shape == (3, 17)
        assert it["rigids_0"].is_cuda and bool(torch.isfinite(it["rigids_0"]).all())
