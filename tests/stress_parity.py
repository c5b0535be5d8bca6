"""One-off large parity run on the GPU (a test tool, not collected by pytest): rollout of N envs against the oracle.
  python tests/stress_parity.py [ENV_ID] [N] [T]        e.g. 65536 envs x 80 steps = 5.2M env-steps, ~6k TOI events"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
from gym_puzzles_b200 import abi
from parity_util import rollout_compare

env_id = sys.argv[1] if len(sys.argv) > 1 else "MultiRobotPuzzleHeavy-v0"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 80
h = abi.Handle(env_id, N, seed=123, max_episode_steps=40)
rep = rollout_compare(h, env_id, N, T, seed=123, max_episode_steps=40, nthreads=os.cpu_count(), state_every=20,
                      device_path=os.environ.get("STRESS_HOST_PATH") is None)   # default: mrp_step (device-resident)
print(env_id, {k: v for k, v in rep.items() if k != "oracle_stats"}, "oracle toi_events", rep["oracle_stats"]["toi_events"])
assert rep["flag_mismatch"] == 0 and rep["done_mismatch"] == 0 and rep["state_tol_bad"] == 0 and rep["obs_not_close"] == 0
