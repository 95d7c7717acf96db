"""Seeded synthetic MD trajectory in the on-disk formats of the reference's loader (README "Data Preparation";
SURVEY.md Appendix B).  TEST INFRASTRUCTURE: used by oracle/make_golden.py and the input-pipeline tests."""
import os
import pickle

import numpy as np
import torch

from oracle import dfold_oracle as O


def make_trajectory(T: int = 9, N: int = 17, seed: int = 31):
    """-> dict(all_atom_positions [T,N,37,3] fp32, aatype one-hot [N,21], all_atom_mask [N,37], residue_index [N],
    force / vel [T,N,3] fp64, node_repr [N,256], edge_repr [N,N,128]).  Coordinates come from random backbone frames and
    torsions through the idealised-geometry epilogue, so every defined atom sits at a chemically sensible place; two
    residues lose atoms (missing side chain, missing O) to exercise the masks."""
    from dynamicpdb_b200 import synthetic as syn
    g = torch.Generator().manual_seed(seed)
    feats = syn.make_feats(T, N, seed=seed)
    aatype = torch.randint(0, 20, (N,), generator=g)
    ang = torch.nn.functional.normalize(torch.randn(T, N, 7, 2, generator=g), dim=-1)
    aa = aatype[None].expand(T, -1)
    frames = O.torsion_angles_to_frames(feats["rigids_0"], ang, aa)
    atom37, mask37 = O.atom14_to_atom37(O.frames_to_atom14(frames, aa), aa)
    mask = mask37[0].clone()
    mask[3, 5:] = 0                      # residue 3: side chain beyond CB missing
    mask[7, 4] = 0                       # residue 7: no O
    onehot = torch.nn.functional.one_hot(aatype, 21).float()
    return {"all_atom_positions": atom37.float().numpy(), "aatype": onehot.numpy(), "all_atom_mask": mask.numpy(),
            "residue_index": np.arange(N), "force": torch.randn(T, N, 3, generator=g, dtype=torch.float64).numpy(),
            "vel": torch.randn(T, N, 3, generator=g, dtype=torch.float64).numpy(),
            "node_repr": torch.randn(N, 256, generator=g).numpy(), "edge_repr": torch.randn(N, N, 128, generator=g).numpy()}


def write_files(traj, root, name="synth"):
    """-> the csv row fields (atlas_npz, force_path, vel_path, embed_path) of files written under `root`."""
    os.makedirs(root, exist_ok=True)
    npz = os.path.join(root, f"{name}_new_w_pp.npz")
    np.savez(npz, all_atom_positions=traj["all_atom_positions"], aatype=traj["aatype"], all_atom_mask=traj["all_atom_mask"],
             residue_index=traj["residue_index"])
    fpath, vpath = os.path.join(root, f"{name}_F.pkl"), os.path.join(root, f"{name}_V.pkl")
    with open(fpath.replace(".pkl", "_Ca.pkl"), "wb") as f:
        pickle.dump(traj["force"], f)
    with open(vpath.replace(".pkl", "_ca.pkl"), "wb") as f:
        pickle.dump(traj["vel"], f)
    emb = os.path.join(root, f"{name}.npz")
    np.savez(emb, node_repr=traj["node_repr"], edge_repr=traj["edge_repr"])
    return npz, fpath, vpath, emb
