"""CPU, world_size 2, gloo: the data-parallel path (one protein window per rank, DistributedDataParallel with
find_unused_parameters=True exactly as train_DFOLD_dynamics.py:612-616) averages the per-rank gradients."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    from dynamicpdb_b200 import kernels, synthetic as syn
    from dynamicpdb_b200.Dfold_network_dynamic import FullScoreNetwork
    from dynamicpdb_b200.score_epilogue import SE3ScoreDiffuser
    from oracle import ops
    for n in ops.ALL:
        setattr(kernels, n, getattr(ops, n))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    nf, N = 2, 10
    net = FullScoreNetwork(syn.model_conf(nf, **syn.PRESET_TINY), SE3ScoreDiffuser(syn.diffuser_conf(1.0)))
    sd = net.state_dict()
    syn.dezero_(sd)
    net.load_state_dict(sd)
    ddp = torch.nn.parallel.DistributedDataParallel(net, find_unused_parameters=True)
    loss = syn.surrogate_loss(ddp(syn.make_feats(nf, N, seed=rank)))
    loss.backward()
    g = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    # single-process reference: mean of the two windows' gradients
    if rank == 0:
        ref = {}
        for r in range(world):
            net.zero_grad(set_to_none=True)
            syn.surrogate_loss(net(syn.make_feats(nf, N, seed=r))).backward()
            for k, p in net.named_parameters():
                if p.grad is not None:
                    ref[k] = ref.get(k, 0) + p.grad / world
        worst = max(((g[k] - ref[k]).norm() / (ref[k].norm() + 1e-12)).item() for k in ref if ref[k].norm() > 1e-8)
        q.put(worst)
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_two_ranks_average_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    worst = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert worst < 1e-4, worst
