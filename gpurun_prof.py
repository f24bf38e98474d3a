import os, sys
sys.path.insert(0, "/root/repo")
import torch
from bench import env_prop_for
from marl_demandresponse_b200 import BatchedEnv
R, N, layout = 4096, 100, "hand_engineered"
env = BatchedEnv(env_prop_for(N), R, obs_layout=layout, noise="philox", seed=1)
env.reset()
D = env.sim.D
fc = torch.nn.ModuleList([torch.nn.Linear(D, 100), torch.nn.Linear(100, 100), torch.nn.Linear(100, 2)]).cuda()
w = BatchedEnv.actor_weights(fc)
po = torch.zeros((R, env.sim.Ns), dtype=torch.float32, device="cuda")
for _ in range(3):
    env.sim.policy_step(w, seed=1, prob_on=po)
torch.cuda.synchronize()
v = po.flatten()[:16].cpu().numpy()
names = ["wait A1+barrier", "issue fetch", "MMA1 roundtrip", "relu1+barrier", "MMA2 roundtrip", "relu2+barrier", "MMA3 roundtrip", "softmax+draw"]
for s in range(2):
    tot = v[s*8:(s+1)*8].sum()
    print("slot", s, "total cycles", tot, {n: f"{100*x/tot:.0f}%" for n, x in zip(names, v[s*8:(s+1)*8])})
