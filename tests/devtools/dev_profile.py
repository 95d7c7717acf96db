"""Development probe: per-kernel time breakdown of one training step with torch.profiler."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
from dynamicpdb_b200 import synthetic as syn
from dynamicpdb_b200.Dfold_network_dynamic import FullScoreNetwork
from dynamicpdb_b200.score_epilogue import SE3ScoreDiffuser
nf, N = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
net = FullScoreNetwork(syn.model_conf(nf, **syn.PRESET_A), SE3ScoreDiffuser(syn.diffuser_conf(1.0)))
sd = net.state_dict(); syn.dezero_(sd); net.load_state_dict(sd); net = net.cuda()
opt = torch.optim.Adam(net.parameters(), lr=1e-4, amsgrad=True)
feats = {k: v.cuda() for k, v in syn.make_feats(nf, N, seed=0).items()}
def step():
    opt.zero_grad(set_to_none=True)
    loss = syn.surrogate_loss(net(dict(feats))); loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
