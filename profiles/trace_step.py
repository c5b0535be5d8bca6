import os, sys
sys.path.insert(0, "/root/repo")
import torch
from gym_puzzles_b200 import abi
N = int(os.environ.get("QB_ENVS", 1048576))
h = abi.Handle("MultiRobotPuzzleHeavy-v0", N, seed=17)
h.reset()
for t in range(100):
    h.sample_actions(t); h.step()
torch.cuda.synchronize()
os.environ["MRP_TRACE"] = "1"
for t in range(4):
    h.sample_actions(1000 + t); h.step()
torch.cuda.synchronize()
