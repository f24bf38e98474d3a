"""A few policy steps of the 3xTF32 actor on the C4 / C3 observation shape (target of the ncu capture)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import env_prop_for
from marl_demandresponse_b200 import BatchedEnv
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
R, N, layout = (4096, 100, "hand_engineered") if which == "c3" else (2048, 1000, "tarmac")
env = BatchedEnv(env_prop_for(N), R, obs_layout=layout, noise="philox", seed=1)
env.reset()
torch.manual_seed(0)
D = env.sim.D
fc = torch.nn.ModuleList([torch.nn.Linear(D, 100), torch.nn.Linear(100, 100), torch.nn.Linear(100, 2)]).cuda()
w = BatchedEnv.actor_weights(fc)
for _ in range(4):
    env.policy_step(w, precision=sys.argv[2] if len(sys.argv) > 2 else "tf32x3")
torch.cuda.synchronize()
