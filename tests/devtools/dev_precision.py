"""Development probe: error of the split-bf16 GEMM vs K, and per-output errors of the full network vs the oracle."""
import sys, os, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dynamicpdb_b200 import kernels as K, synthetic as syn
from dynamicpdb_b200.Dfold_network_dynamic import FullScoreNetwork
from dynamicpdb_b200.score_epilogue import SE3ScoreDiffuser
from oracle import dfold_oracle as O

torch.manual_seed(0)
for Kd in (256, 2048, 16384, 32000):
    x = torch.randn(256, Kd, device="cuda"); w = torch.randn(256, Kd, device="cuda") / math.sqrt(Kd)
    y = K.linear(x, w)
    ref = (x.double() @ w.double().T)
    e = (y.double() - ref)
    y32 = (x @ w.T).double()   # cuBLAS fp32 (may use TF32? default off)
    print(f"K={Kd:6d} rel_max={e.abs().max().item()/ref.abs().max().item():.2e} rms={e.pow(2).mean().sqrt().item()/ref.pow(2).mean().sqrt().item():.2e} "
          f"bias(mean e*sign(ref))={(e*ref.sign()).mean().item()/ref.abs().mean().item():+.2e}  torch_fp32_rms={(y32-ref).pow(2).mean().sqrt().item()/ref.pow(2).mean().sqrt().item():.2e}")
    # positive operands: exposes truncation bias in the accumulator
    xp, wp = x.abs(), w.abs()
    yp = K.linear(xp, wp); rp = xp.double() @ wp.double().T
    print(f"          positive operands: mean rel err={(yp.double()-rp).mean().item()/rp.mean().item():+.2e}  max={((yp.double()-rp).abs().max()/rp.abs().max()).item():.2e}")

for name, preset, nf, N in [("A", syn.PRESET_A, 4, 40), ("A", syn.PRESET_A, 8, 64)]:
    conf = syn.model_conf(nf, **preset)
    net = FullScoreNetwork(conf, SE3ScoreDiffuser(syn.diffuser_conf(1.0)))
    sd = net.state_dict(); syn.dezero_(sd); net.load_state_dict(sd)
    feats = syn.make_feats(nf, N, seed=11)
    p = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    t0 = time.time()
    out_o = O.full_forward(p, feats, O.default_conf(**preset), O.default_diffuser_conf(1.0))
    loss_o = O.surrogate_loss(out_o)
    names = [k for k in p if p[k].requires_grad]
    g_o = dict(zip(names, torch.autograd.grad(loss_o, [p[k] for k in names], allow_unused=True)))
    print(f"{name} nf={nf} N={N} oracle {time.time()-t0:.1f}s")
    net = net.cuda()
    out_g = net({k: v.cuda() for k, v in feats.items()})
    loss_g = syn.surrogate_loss(out_g); loss_g.backward(); torch.cuda.synchronize()
    for k in out_o:
        a, b = out_o[k].double(), out_g[k].detach().cpu().double()
        print(f"   {k:14s} per-res L2 max={(a-b).flatten(2).norm(dim=-1).max().item():.2e}  maxabs={(a-b).abs().max().item():.2e}  refmax={a.abs().max().item():.2e}")
    errs = []
    for k, prm in net.named_parameters():
        go = g_o.get(k)
        if go is None or prm.grad is None: continue
        sc = go.abs().max().item()
        if sc < 1e-7: continue
        errs.append(((go - prm.grad.cpu()).abs().max().item() / sc, k))
    errs.sort(reverse=True)
    print("   worst grads:", [(f"{e:.1e}", k) for e, k in errs[:6]])
