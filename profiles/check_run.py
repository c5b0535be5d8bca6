"""Every kernel family once, default and wide build, against the CHECKING build of the library (-DMRP_CHECK: in-kernel bounds
assertions on body fields vs. shared-memory layout, fixture / contact slots, pool records, queue slots; compute-sanitizer is
not available on this pool).  Under gpurun:
  bash profiles/build_probe.sh check && MRP_LIB_PATH=$PWD/gym_puzzles_b200/csrc/libmrp_check.so python profiles/check_run.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("MRP_CHUNKS_HOST", "8")
os.environ.setdefault("MRP_HOST_WAVES", "3")     # front-half waves of mrp_step_host
os.environ.setdefault("MRP_SPARES", "1")         # spare episodes: refill pass (list mode of the pipeline kernels) + reset by copy
os.environ.setdefault("MRP_REFILL_MIN", "1")
os.environ.setdefault("MRP_OVERLAP_POST", "1")
os.environ.setdefault("MRP_BIG", "1")
os.environ.setdefault("MRP_GRAPH", "0")
import numpy as np
import torch

import gym_puzzles_b200 as gp
from gym_puzzles_b200 import abi

rng = np.random.default_rng(0)
for env_id, n_agents, N in [("MultiRobotPuzzleHeavy-v0", 0, 1500), ("MultiRobotPuzzle-v0", 0, 700), ("MultiRobotPuzzle-v2", 0, 700),
                            ("MultiRobotPuzzleHeavy-v2", 5, 300), ("MultiRobotPuzzleSquare-v2", 0, 400)]:
    h = abi.Handle(env_id, N, seed=1, n_agents=n_agents, max_episode_steps=12)
    h.enable_terminal_info()
    if env_id.endswith("v2"):
        h.enable_curriculum()
    h.reset_host()
    for t in range(30):
        if t % 2:
            h.step_host(rng.uniform(-1, 1, (N, h.act_dim)).astype(np.float32))
        else:
            h.sample_actions(t)
            h.step()
    w = h.get_state()
    h.set_state(w)
    h.step()
    torch.cuda.synchronize()
    print(env_id, n_agents, h.stats())
    h.close()
env = gp.VecNormalize(gp.VectorEnv("MultiRobotPuzzleHeavy-v0", 2000, seed=2, max_episode_steps=10))
env.venv.enable_terminal_info()
env.reset()
for t in range(25):
    env.venv.sample_actions(t)
    env.step()
torch.cuda.synchronize()
print("vecnorm ok", env.state_dict()["obs_rms.count"])
venv = gp.SB3VecEnv("MultiRobotPuzzle-v0", 256, seed=3, max_episode_steps=8)
venv.reset()
for t in range(20):
    venv.step(rng.uniform(-1, 1, (256, 6)).astype(np.float32))
print("sb3 ok")
import ctypes as C
counts = (C.c_uint * 8)()
wide = (C.c_uint * 8)()
lib = abi.load().lib
assert lib.mrp_debug_check_counts(counts) == 0 and lib.mrp_debug_check_counts_wide(wide) == 0
print("MRP_CHECK failed assertions (body field, fixture, contact slot, record, queue, nc):", list(counts)[:6], "wide build:", list(wide)[:6])
assert sum(counts) == 0 and sum(wide) == 0
