import sys, os; sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
import numpy as np, torch
from diverse_conventions_b200 import layouts
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet
g = np.load('tests/golden/policy_simple_h64.npz')
lp = layouts.load_layout('simple', 400)
nets = {}
for kind in ('actor','critic'):
    net = PolicyNet(kind, 5, 4, 20, 64)
    net.load_state_dict({k[len(kind)+1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(kind+'.')})
    nets[kind] = net
pol = FusedPolicy(lp, 64, 1); pol.set_weights(0, nets['actor'], nets['critic'])
obs = torch.from_numpy(g['obs']).cuda()
out = pol.act(obs, deterministic=True, want_logits=True); val = pol.value(obs)
torch.cuda.synchronize()
lg = out['logits'].cpu(); ref = torch.from_numpy(g['logits'])
print("logits rel err", float((lg-ref).abs().max()/ref.abs().max()), "max ref", float(ref.abs().max()))
print(lg[:3]); print(ref[:3])
v = val.cpu(); rv = torch.from_numpy(g['values'])[:,0]
print("values rel err", float((v-rv).abs().max()/rv.abs().max())); print(v[:6], rv[:6])
