import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upkie_b200.envs import B200VectorEnv
from upkie_b200.model import Model
m = Model.standard_upkie()
n = 65536
env = B200VectorEnv(n, "servos", model=m)
env.reset(seed=0)
a = np.zeros((n,6,6), np.float32); a[:,:,0]=np.nan; a[:,:,5]=m.tau_max
acts=[torch.from_numpy(a).pin_memory().numpy() for _ in range(4)]
for k in range(20): env.step(acts[k%4])
torch.cuda.synchronize(); t0=time.perf_counter()
K=400
for k in range(K): env.step(acts[k%4])
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print(os.environ.get("UPKIE_B200_HOST_CHUNKS"), "ms/step %.4f"%(dt/K*1e3), "env-steps/s %.3e"%(n*K/dt))
