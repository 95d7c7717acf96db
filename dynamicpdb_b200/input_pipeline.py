"""Input pipeline of the DFOLDv2 trainer on the device.

The reference builds every training sample on the host, inside ``PdbDataset.__getitem__`` (src/data/
Dfold_data_loader_dynamic.py:192-259, :312-360): it re-opens and decompresses the WHOLE trajectory archive of the protein
(``dict(np.load(...))``), unpickles the whole force / velocity trajectories, slices ``frame_time`` frames out of them, and
runs the OpenFold feature transforms (``atom37_to_frames``, ``atom37_to_torsion_angles``) plus a CPU eigen-decomposition
(``rot_to_quat``) on the slice.  At the throughput of this package (hundreds of frames per second per GPU) that loader
cannot keep one GPU fed, let alone eight.  Here:

* ``TrajectoryStore`` keeps each protein's trajectory decoded ONCE, in pinned host memory (atom37 coordinates as fp32:
  that is what the archive holds), and hands out frame windows with the reference's own index arithmetic and random
  stream (``select_random_samples`` / ``select_first_samples``, :164-190);
* ``featurize_window`` computes ``rigids_0`` and the torsion features of a window with one kernel on the device
  (csrc/featurize.cu), from the fp32 coordinates of the ``frame_time`` selected frames only;
* ``WindowPrefetcher`` overlaps the next window's slicing and host->device copy (side stream, pinned staging buffers)
  with the current training step.

The forward-marginal noising (``diffuser.forward_marginal``, :337-342) stays with the caller's diffuser, as the north star
leaves the diffusion loop in Python.  Evaluation-only fields (``atom37_pos``, ``atom14_pos``, ``rigidgroups_0``,
``residx_atom14_to_atom37``) are not produced.
"""
import pickle
import threading
from typing import Dict, Iterator, Optional, Sequence, Tuple

import numpy as np
import torch

from . import kernels as K
from .feats import table


def featurize_window(atom37: torch.Tensor, atom_mask: torch.Tensor, aatype: torch.Tensor) -> Dict[str, torch.Tensor]:
    """atom37 [nf,N,37,3] fp32 (device), atom_mask [N,37], aatype [N] int64 ->
    {rigids_0 [nf,N,7] fp32, torsion_angles_sin_cos / alt_torsion_angles_sin_cos [nf,N,7,2] fp64, torsion_angles_mask
    [nf,N,7] fp64, res_mask [nf,N]} as ``_process_csv_row`` + ``__getitem__`` (rigids_0 up to the quaternion sign, which the
    reference's eigen-decomposition leaves arbitrary)."""
    K._need_cuda(atom37, atom_mask, aatype)
    nf, N = atom37.shape[0], atom37.shape[1]
    dev = atom37.device
    pos = atom37.to(torch.float32).contiguous()
    am = atom_mask.to(torch.float32).contiguous()
    aa = aatype.to(torch.int64).contiguous()
    rig = torch.empty(nf, N, 7, dtype=torch.float32, device=dev)
    tor = torch.empty(nf, N, 7, 2, dtype=torch.float64, device=dev)
    alt = torch.empty_like(tor)
    tm = torch.empty(nf, N, 7, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        K._check(K.lib().dfold_featurize_window(K._ptr(pos), K._ptr(am), K._ptr(aa), K._ptr(table("chi_atom_indices", dev)),
                                                K._ptr(table("chi_angles_mask", dev)), K._ptr(table("chi_pi_periodic", dev)),
                                                nf, N, K._ptr(rig), K._ptr(tor), K._ptr(alt), K._ptr(tm),
                                                ctypes_stream(dev)), "dfold_featurize_window")
    return {"rigids_0": rig, "torsion_angles_sin_cos": tor, "alt_torsion_angles_sin_cos": alt, "torsion_angles_mask": tm,
            "res_mask": am[:, 1].to(torch.float64).unsqueeze(0).expand(nf, -1)}


def ctypes_stream(dev):
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class TrajectoryStore:
    """Decoded trajectories of the proteins of a split, pinned in host memory, keyed by the csv row fields the reference
    reads (``atlas_npz``, ``force_path``, ``vel_path``, ``embed_path``; README "Data Preparation")."""

    def __init__(self, frame_time: int, frame_sample_step: int = 1, keep_first: Optional[int] = None, fix_sample_start: int = 0,
                 pin: bool = True):
        self.frame_time, self.step, self.keep_first, self.fix_start = frame_time, frame_sample_step, keep_first, fix_sample_start
        self.pin = pin and torch.cuda.is_available()
        self._cache: Dict[str, Dict[str, torch.Tensor]] = {}

    def _pin(self, a: np.ndarray, dtype) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
        return t.pin_memory() if self.pin else t

    def protein(self, atlas_npz: str, force_path: str, vel_path: str, embed_path: str) -> Dict[str, torch.Tensor]:
        """Load (once) and cache one protein: the file naming follows Dfold_data_loader_dynamic.py:194-198."""
        if atlas_npz in self._cache:
            return self._cache[atlas_npz]
        with np.load(atlas_npz, allow_pickle=True) as zf:
            z = {k: zf[k] for k in ("all_atom_positions", "all_atom_mask", "aatype", "residue_index")}
        with open(force_path.replace(".pkl", "_Ca.pkl"), "rb") as f:
            force = pickle.load(f)
        with open(vel_path.replace(".pkl", "_ca.pkl"), "rb") as f:
            vel = pickle.load(f)
        with np.load(embed_path) as ef:
            emb = {k: ef[k] for k in ("node_repr", "edge_repr")}
        kf = self.keep_first
        ent = {
            "atom37": self._pin(z["all_atom_positions"][:kf], torch.float32),                 # [T,N,37,3]
            "force": self._pin(np.asarray(force)[:kf], torch.float64),                        # [T,N,3] (the loader's dtype)
            "vel": self._pin(np.asarray(vel)[:kf], torch.float64),
            "atom_mask": self._pin(z["all_atom_mask"], torch.float32),                        # [N,37]
            "aatype": torch.from_numpy(np.argmax(z["aatype"], axis=-1)).long(),               # [N]
            "residue_index": torch.from_numpy(np.asarray(z["residue_index"])).long(),
            "node_repr": self._pin(emb["node_repr"], torch.float32),
            "edge_repr": self._pin(emb["edge_repr"], torch.float32),
        }
        self._cache[atlas_npz] = ent
        return ent

    def window_index(self, n_frames: int, training: bool) -> slice:
        """The reference's window arithmetic (:164-190), drawing from numpy's global stream exactly as it does."""
        t, k = self.frame_time, self.step
        if training:
            if t > n_frames:
                raise ValueError("t cannot be greater than the number of samples")
            start = np.random.randint(0, n_frames - t * k + 1)
        else:
            start = self.fix_start
        return slice(start, start + t * k, k)


class WindowPrefetcher:
    """Iterates device-resident feature dicts, one protein window per item: while the consumer trains on window i, a worker
    thread slices window i+1 into pinned staging buffers and a side stream copies it to the device and featurises it."""

    def __init__(self, store: TrajectoryStore, rows: Sequence[Tuple[str, str, str, str]], device, training: bool = True, depth: int = 2):
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.store, self.rows, self.device, self.training, self.depth = store, list(rows), dev, training, depth
        self.stream = torch.cuda.Stream(device=self.device)
        self._staging: Dict = {}
        self._slot: Dict[str, int] = {}
        self._events: list = []

    def _stage(self, name: str, src: torch.Tensor, sl: slice) -> torch.Tensor:
        """The (possibly strided) frame window gathered into a pinned staging buffer: a ring of depth + 1 buffers per field,
        so a buffer is reused only after the consumer has received the window that was copied from it."""
        n = len(range(*sl.indices(src.shape[0])))
        shape = (n,) + tuple(src.shape[1:])
        ring = self._staging.setdefault((name, shape, src.dtype), [])
        if len(ring) <= self.depth:
            buf = torch.empty(shape, dtype=src.dtype)
            ring.append(buf.pin_memory() if self.store.pin else buf)
        slot = self._slot.get(name, 0)
        self._slot[name] = (slot + 1) % (self.depth + 1)
        buf = ring[min(slot, len(ring) - 1)]
        buf.copy_(src[sl])
        return buf

    def _make(self, row) -> Tuple[Dict[str, torch.Tensor], torch.cuda.Event]:
        ent = self.store.protein(*row)
        # the staging buffers of this window were last used depth + 1 windows ago: that copy must have left the host
        if len(self._events) > self.depth:
            self._events[-(self.depth + 1)].synchronize()
        sl = self.store.window_index(ent["atom37"].shape[0], self.training)
        nf = self.store.frame_time
        with torch.cuda.stream(self.stream):
            cp = lambda t: t.to(self.device, non_blocking=True)
            atom37 = cp(self._stage("atom37", ent["atom37"], sl))
            feats = featurize_window(atom37, cp(ent["atom_mask"]), cp(ent["aatype"]))
            feats.update({
                "aatype": cp(ent["aatype"]).unsqueeze(0).expand(nf, -1),
                "seq_idx": cp(ent["residue_index"]).unsqueeze(0).expand(nf, -1),
                "residue_index": cp(ent["residue_index"]).unsqueeze(0).expand(nf, -1),
                "force": cp(self._stage("force", ent["force"], sl)), "vel": cp(self._stage("vel", ent["vel"], sl)),
                "node_repr": cp(ent["node_repr"]), "edge_repr": cp(ent["edge_repr"]),
                "fixed_mask": torch.zeros(nf, atom37.shape[1], dtype=torch.float64, device=self.device),
                "sc_ca_t": torch.zeros(nf, atom37.shape[1], 3, dtype=torch.float32, device=self.device),
            })
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._events.append(ev)
        del self._events[:-(self.depth + 2)]
        return feats, ev

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        import queue
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)

        def work():
            try:
                torch.cuda.set_device(self.device)
                for row in self.rows:
                    q.put(self._make(row))
                q.put(None)
            except BaseException as e:       # noqa: BLE001  (hand the failure to the consumer instead of blocking it forever)
                q.put(e)
        th = threading.Thread(target=work, daemon=True)
        th.start()
        while True:
            item = q.get(timeout=600)
            if item is None:
                break
            if isinstance(item, BaseException):
                raise RuntimeError("WindowPrefetcher worker failed") from item
            feats, ev = item
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield feats
