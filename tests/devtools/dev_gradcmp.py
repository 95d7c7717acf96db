"""Development probe: compare intermediate activations / gradients of the product on GPU (CUDA kernels) with the same
module code running on CPU over the oracle ops."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dynamicpdb_b200 import kernels as K, synthetic as syn
from dynamicpdb_b200.Dfold_network_dynamic import FullScoreNetwork
from dynamicpdb_b200.score_epilogue import SE3ScoreDiffuser
from oracle import ops as OPS

preset = getattr(syn, "PRESET_" + (sys.argv[1] if len(sys.argv) > 1 else "B"))
nf, N = int(sys.argv[2]) if len(sys.argv) > 2 else 2, int(sys.argv[3]) if len(sys.argv) > 3 else 24
torch.manual_seed(0)
conf = syn.model_conf(nf, **preset)
net = FullScoreNetwork(conf, SE3ScoreDiffuser(syn.diffuser_conf(1.0)))
sd = net.state_dict(); syn.dezero_(sd); net.load_state_dict(sd)
feats = syn.make_feats(nf, N, seed=11); feats["res_mask"][:, -2:] = 0

def run(net, feats, tag):
    acts, grads, order = {}, {}, []
    hooks = []
    counter = {}
    def mk(name):
        def hook(mod, inp, out):
            o = out[0] if isinstance(out, tuple) else out
            if not isinstance(o, torch.Tensor) or not o.requires_grad: return
            c = counter.get(name, 0); counter[name] = c + 1
            key = f"{name}#{c}"
            order.append(key); acts[key] = o.detach().double().cpu()
            o.register_hook(lambda g, key=key: grads.__setitem__(key, g.detach().double().cpu()))
        return hook
    for n, m in net.named_modules():
        if n and (n.count(".") <= 2) and not n.endswith("embedding_layer"):
            hooks.append(m.register_forward_hook(mk(n)))
    out = net(dict(feats)); loss = syn.surrogate_loss(out); loss.backward()
    for h in hooks: h.remove()
    pg = {k: p.grad.detach().double().cpu() for k, p in net.named_parameters() if p.grad is not None}
    return acts, grads, order, pg, loss.item()

saved = {n: getattr(K, n) for n in OPS.ALL}
for n in OPS.ALL: setattr(K, n, getattr(OPS, n))
a0, g0, order, pg0, l0 = run(copy.deepcopy(net), feats, "cpu")
for n in OPS.ALL: setattr(K, n, saved[n])
netg = copy.deepcopy(net).cuda()
a1, g1, order1, pg1, l1 = run(netg, {k: v.cuda() for k, v in feats.items()}, "gpu")
print("loss", l0, l1)
rel = lambda a, b: ((a - b).norm() / (a.norm() + 1e-30)).item()
for k in order:
    ga = f"{rel(g0[k], g1[k]):.2e}" if k in g0 and k in g1 else "-"
    print(f"{k:50s} act {rel(a0[k], a1[k]):.2e}   grad {ga}")
print("---- parameter gradients (rel L2) > 1e-3")
for k in pg0:
    if k in pg1 and pg0[k].norm() > 1e-9:
        e = rel(pg0[k], pg1[k])
        if e > 1e-3: print(f"{k:60s} {e:.2e}  maxrel {((pg0[k]-pg1[k]).abs().max()/pg0[k].abs().max()).item():.2e}")

if "--swap" not in sys.argv: sys.exit(0)
print("==== swap one CUDA op at a time for its torch statement (on GPU) and re-measure")
def param_err(pg1):
    worst = 0
    for k in pg0:
        if k in pg1 and pg0[k].norm() > 1e-9 and "linear_b.bias" not in k:
            worst = max(worst, rel(pg0[k], pg1[k]))
    return worst
for name in OPS.ALL:
    setattr(K, name, getattr(OPS, name))
    try:
        _, g2, _, pg2, l2 = run(copy.deepcopy(net).cuda(), {k: v.cuda() for k, v in feats.items()}, "gpu")
        kk = "score_model.trunk.bb_update_3#0"
        print(f"swap {name:18s} bb3-out grad err {rel(g0[kk], g2[kk]):.2e}   worst param grad err {param_err(pg2):.2e}")
    except Exception as e:
        print("swap", name, "failed", repr(e)[:200])
    setattr(K, name, saved[name])
