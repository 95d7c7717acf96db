import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dynamicpdb_b200 import kernels as K
from oracle import ops as O
import numpy as np
F, N = 3, 70
gen = torch.Generator().manual_seed(5)
q_pred = torch.randn(F, N, 4, generator=gen) * 1.2
q_t = torch.nn.functional.normalize(q_pred + q_pred.norm(dim=-1, keepdim=True) * torch.randn(F, N, 4, generator=gen) * torch.logspace(-1.5, 0.5, N)[None, :, None], dim=-1)
g = torch.from_numpy(np.log(np.linspace(0.0, 1.0, 1000) * np.exp(1.5) + (1 - np.linspace(0.0, 1.0, 1000)) * np.exp(0.1)))
kw = dict(max_sigma=1.5, min_sigma=0.1, min_b=0.1, max_b=20.0, r3_scale=0.1, ipa_scale=2.0, L=1000)
for tval in (0.03, 0.5):
    t = torch.tensor([tval], dtype=torch.float64)
    ro, _ = O.score_epilogue(q_pred, q_t, None, None, t, g, None, **kw)
    rk, _ = K.score_epilogue(q_pred.cuda(), q_t.cuda(), None, None, t.cuda(), g.cuda(), None, **kw)
    rk = rk.cpu()
    # fp64 truth
    ro64, _ = O.score_epilogue(q_pred.double(), q_t.double(), None, None, t, g, None, **kw)
    e = (ro - rk).norm(dim=-1) / ro.norm(dim=-1).clamp(min=1.0)
    e64o = (ro - ro64).norm(dim=-1) / ro64.norm(dim=-1).clamp(min=1.0)
    e64k = (rk - ro64).norm(dim=-1) / ro64.norm(dim=-1).clamp(min=1.0)
    i = int(e.flatten().argmax())
    f, n = i // N, i % N
    from oracle import dfold_oracle as OO
    q0inv = q_pred * q_pred.new_tensor([1., -1, -1, -1]) / (q_pred * q_pred).sum(-1, keepdim=True)
    vec = OO.quat_to_rotvec(OO.quat_mul(q0inv, q_t))
    om = vec.norm(dim=-1)
    print(f"t={tval}: max err kernel-vs-oracle32 {e.max():.3e} at ({f},{n}) omega={om[f,n]:.4f} ref={ro[f,n].tolist()} ker={rk[f,n].tolist()}")
    print(f"   oracle32 vs fp64 truth max {e64o.max():.3e};  kernel vs fp64 truth max {e64k.max():.3e}")
    worst = torch.topk(e.flatten(), 5)
    print("   top5 errs", worst.values.tolist(), "omegas", om.flatten()[worst.indices].tolist())
