"""One data-parallel training step of the score network, captured in a CUDA graph.

The reference trainer wraps the model in ``DistributedDataParallel(find_unused_parameters=True)`` and calls
``loss.backward(); optimizer.step()`` from Python (train_DFOLD_dynamics.py:612-616, 660-668); that keeps working on top
of this package (tests/test_cpu_ddp.py).  At B200 kernel speeds the ~2000 kernel launches of a step cost as much host
time as device time, so the B200-native step is:

    static input buffers  ->  [ forward + loss + backward + ONE flat NCCL all-reduce of the gradients + Adam ]  (one graph)

* gradients live in a single flat fp32 buffer (each ``param.grad`` is a view), so data parallelism is one
  ``all_reduce`` of 737.7 MB over NVLink per step and parameters without gradient (the dead-output embedder,
  ``linear_rbf``) simply contribute zeros — the reason the reference needs ``find_unused_parameters=True``;
* ``torch.cuda.graph`` captures the whole step (the C-ABI kernels are plain stream launches; TMA descriptors are
  by-value kernel arguments), replays have no Python / launch overhead;
* if capture is not possible the same step runs eagerly (``graph=False`` or on capture failure).
"""
from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist

from . import kernels as _kernels


class TrainStep:
    def __init__(self, net: torch.nn.Module, loss_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor],
                 example_feats: Dict[str, torch.Tensor], *, lr: float = 1e-4, world_size: int = 1, graph: bool = True,
                 warmup: int = 3):
        self.net, self.loss_fn, self.world = net, loss_fn, world_size
        self.params = [p for p in net.parameters() if p.requires_grad]
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        # Adam(amsgrad) as the reference trainer (train_DFOLD_dynamics.py:412); capturable keeps the step count on device
        self.opt = torch.optim.Adam(self.params, lr=lr, amsgrad=True, capturable=True, foreach=True)
        self.static = {k: v.clone() for k, v in example_feats.items()}
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_error: Optional[str] = None
        import os, sys, time
        verbose = os.environ.get("DFOLD_BENCH_VERBOSE", "0") == "1"
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for it in range(max(1, warmup)):
                self._eager()
                if verbose:
                    torch.cuda.synchronize()
                    print(f"[train_step t={time.perf_counter():.1f}] eager warm-up step {it} done", file=sys.stderr, flush=True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._eager()
                self.graph = g
            except Exception as e:                       # noqa: BLE001  (report and fall back to eager launches)
                self.graph_error = f"{type(e).__name__}: {str(e)[:200]}"
                self.graph = None
                torch.cuda.synchronize()

    def _eager(self):
        self.flat_grad.zero_()
        out = self.net(dict(self.static))
        loss = self.loss_fn(out)
        loss.backward()
        if self.world > 1:
            dist.all_reduce(self.flat_grad)
            self.flat_grad.mul_(1.0 / self.world)
        self.opt.step()
        self.loss.copy_(loss.detach().float())

    def load(self, feats: Dict[str, torch.Tensor]):
        """Copy a new window into the static input buffers (H2D when `feats` is pinned host memory)."""
        for k, v in feats.items():
            self.static[k].copy_(v, non_blocking=True)

    def __call__(self) -> torch.Tensor:
        if self.graph is not None:
            self.graph.replay()
            _kernels.invalidate_weight_cache()      # the replay moved the weights without touching their version counters
        else:
            self._eager()
        return self.loss
