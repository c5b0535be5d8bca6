"""Latency of one synchronous mrp_step_host call (numpy in / numpy out) at SB3-scale batch sizes.
  python profiles/small_batch.py [ENV_ID]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from gym_puzzles_b200 import abi

env_id = sys.argv[1] if len(sys.argv) > 1 else "MultiRobotPuzzle-v0"
for N in [int(x) for x in os.environ.get("SB_SIZES", "1,6,64,1024,16384,65536").split(",")]:
    h = abi.Handle(env_id, N, seed=1)
    pin = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt).pin_memory().numpy()  # noqa: E731
    a = pin(N, h.act_dim)
    a[:] = np.random.default_rng(0).uniform(-1, 1, a.shape)   # one action batch held for the whole run: steady pushing, a heavy case
    out = (pin(N, h.obs_dim), pin(N), pin(N, dt=torch.uint8), pin(N, dt=torch.uint8))
    h.reset_host()
    for _ in range(50):
        h.step_host(a, *out)
    t0 = time.perf_counter()
    K = 300
    for _ in range(K):
        h.step_host(a, *out)
    dt = (time.perf_counter() - t0) / K
    print(f"{env_id} N={N:6d}  {dt * 1e6:8.1f} us/step  {N / dt:12.0f} env-steps/s", flush=True)
    h.close()
