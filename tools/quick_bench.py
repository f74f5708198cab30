import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from upkie_b200.model import Model
from upkie_b200 import _abi
from upkie_b200.sim import UpkieSim, neutral_action
m = Model.standard_upkie()
for n in (4096, 65536, 262144):
    cfg = _abi.default_sim_config()
    sim = UpkieSim(n, model=m, config=cfg)
    sim.reset(seed=1)
    a = neutral_action(m, n, device='cuda')
    a[:, :, 0] = 0.0; a[:, [2,5], 0] = float('nan')
    for _ in range(5): sim.step_servos(a)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    K = 50
    e0.record()
    for _ in range(K): sim.step_servos(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"n={n} ms/step={ms:.4f} env-steps/s={n/ms*1e3:.3e}", flush=True)
    st = sim.get_state().cpu().numpy()
    print("  z mean", st[:,2].mean(), "contact frac", st[:,40].mean(), "nan", np.isnan(st).any())
