import sys, time, os; sys.path.insert(0,'/root/repo')
import torch
from gym_puzzles_b200 import abi
def run(env_id, N, settle=100, K=20):
    h = abi.Handle(env_id, N, seed=17)
    h.reset(); torch.cuda.synchronize()
    for t in range(settle):
        h.sample_actions(t); h.step()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    h.set_timing(True); h.get_timing(); h.get_phase_timing()
    e0.record()
    for t in range(K):
        h.sample_actions(1000+t); h.step()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/K
    ph=h.get_phase_timing()
    print(env_id, N, 'ms/step %.3f'%ms, 'env-steps/s %.3e'%(N/ms*1e3), {k:round(v/K,3) for k,v in ph.items()}, flush=True)
    h.close()
for env_id, N in [("MultiRobotPuzzleHeavy-v0", 1048576)] + ([("MultiRobotPuzzle-v0", 1048576), ("MultiRobotPuzzle-v2", 1048576)] if os.environ.get('ALL') else []):
    run(env_id, N)
