"""Input pipeline on the host: the oracle restatement of the loader's featurisation and the window
arithmetic of TrajectoryStore against tests/golden/loader.pt (generated from the unmodified PdbDataset._process_csv_row)."""
import numpy as np
import torch

from dynamicpdb_b200.input_pipeline import TrajectoryStore
from oracle import dfold_oracle as O
from oracle import synth_traj
from tests.test_cpu_oracle import close, load

CONF = dict(frame_time=3, frame_sample_step=2, keep_first=8)


def _window(tmp_path):
    traj = synth_traj.make_trajectory()
    row = synth_traj.write_files(traj, str(tmp_path))
    store = TrajectoryStore(pin=False, **CONF)
    ent = store.protein(*row)
    np.random.seed(5)                                   # the golden's numpy stream
    sl = store.window_index(ent["atom37"].shape[0], training=True)
    return store, ent, sl, row


def test_window_selection_follows_the_reference_stream(tmp_path):
    gold = load("loader")
    store, ent, sl, row = _window(tmp_path)
    assert close(ent["force"][sl], gold["force"], 0) and close(ent["vel"][sl], gold["vel"], 0)
    assert ent["force"][sl].dtype == gold["force"].dtype
    assert torch.equal(ent["aatype"][None].expand(3, -1), gold["aatype"])
    assert torch.equal(ent["residue_index"][None].expand(3, -1), gold["seq_idx"])
    assert store.protein(*row) is ent                   # cached: the archive is decoded once
    assert store.window_index(8, training=False) == slice(0, 6, 2)


def test_oracle_featurisation_matches_reference_loader(tmp_path):
    gold = load("loader")
    _, ent, sl, _ = _window(tmp_path)
    R0, t0, sc, alt, mask = O.featurize_window(ent["atom37"][sl], ent["atom_mask"], ent["aatype"])
    assert close(R0, gold["rot_0"].double(), 2e-6) and close(t0, gold["rigids_0"][..., 4:].double(), 1e-6)
    assert close(sc, gold["torsion_angles_sin_cos"], 1e-6)      # the reference runs this transform on fp64 copies of fp32 data
    assert close(alt, gold["alt_torsion_angles_sin_cos"], 1e-6)
    assert torch.equal(mask, gold["torsion_angles_mask"])
    assert close(ent["atom_mask"][:, 1].double()[None].expand(3, -1), gold["res_mask"].double(), 0)
    # rigids_0: the reference's quaternion (eigen-decomposition, arbitrary sign) encodes the same rotation
    q = gold["rigids_0"][..., :4].double()
    assert close(O.quat_to_rot(q / q.norm(dim=-1, keepdim=True)), R0, 2e-6)
