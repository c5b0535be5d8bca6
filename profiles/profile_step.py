"""Short program for the ncu passes: MultiRobotPuzzleHeavy-v0 (or argv[1]), 524288 envs, 60 settle steps + 4 steps.
Launches per step: k_clear, k_broad, k_narrow, k_pre, k_solve_vel, k_solve_pos, k_post, k_post_events, k_reset_list
(+ k_sample_actions)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from gym_puzzles_b200 import abi

env_id = sys.argv[1] if len(sys.argv) > 1 else "MultiRobotPuzzleHeavy-v0"
N = int(os.environ.get("QB_ENVS", 524288))
h = abi.Handle(env_id, N, seed=17)
h.reset()
for t in range(64):
    h.sample_actions(t)
    h.step()
torch.cuda.synchronize()
print("ok", env_id, N, h.launch_count)
