"""Timeline of a step with a steady stream of auto-resets (TimeLimit 200, de-phased): MRP_TRACE=1 python profiles/trace_resets.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import gym_puzzles_b200 as gp
N, cap = int(os.environ.get("QB_ENVS", 1048576)), 200
env = gp.VectorEnv("MultiRobotPuzzleHeavy-v0", N, device="cuda:0", seed=17, max_episode_steps=cap)
env.reset()
phase = torch.arange(N, device="cuda:0") % cap
for t in range(2 * cap + 30):
    if t < cap:
        env.reset((phase == t).to(torch.uint8))
    env.sample_actions(step_index=t)
    env.step()
torch.cuda.synchronize()
