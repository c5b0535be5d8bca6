"""Quick device-side A/B helper (not the bench contract — see bench.py): ms/step of mrp_step for a few settings.

  python profiles/quickbench.py [ENV_ID ...]        env: MRP_CHUNKS, MRP_SOLVER_CTAS, QB_ENVS, QB_PHASES=1, QB_E2E=1, QB_MAXSTEPS (TimeLimit)
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from gym_puzzles_b200 import abi


def run(env_id, N, settle=100, K=20, n_agents=0):
    h = abi.Handle(env_id, N, seed=17, n_agents=n_agents, max_episode_steps=int(os.environ.get("QB_MAXSTEPS", 0)))
    h.reset()
    torch.cuda.synchronize()
    for t in range(settle):
        h.sample_actions(t)
        h.step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    phases = bool(os.environ.get("QB_PHASES"))
    if phases:
        h.set_timing(True)
        h.get_timing()
        h.get_phase_timing()
    e0.record()
    for t in range(K):
        h.sample_actions(1000 + t)
        h.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    ph = {k: round(v / K, 3) for k, v in h.get_phase_timing().items()} if phases else ""
    if phases:
        h.set_timing(False)
    line = f"{env_id} n_agents={n_agents} N={N} chunks={os.environ.get('MRP_CHUNKS', 'dflt')} ms/step {ms:.3f} env-steps/s {N / ms * 1e3:.3e} {ph}"
    if os.environ.get("QB_E2E"):
        A, O = h.act_dim, h.obs_dim
        pin = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt).pin_memory().numpy()
        acts = [pin(N, A) for _ in range(4)]
        rng = np.random.default_rng(0)   # a different batch per buffer: constant actions would be a heavier workload (steady pushing)
        for a in acts:
            a[:] = rng.uniform(-1, 1, a.shape)
        out = (pin(N, O), pin(N), pin(N, dt=torch.uint8), pin(N, dt=torch.uint8))
        h.step_host(acts[0], *out)
        torch.cuda.synchronize()
        e0.record()
        for t in range(10):
            h.step_host(acts[t % 4], *out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        line += f" | e2e host_chunks={os.environ.get('MRP_CHUNKS_HOST', 'dflt')} ms/step {ms:.3f} env-steps/s {N / ms * 1e3:.3e}"
    print(line, flush=True)
    h.close()


if __name__ == "__main__":
    ids = sys.argv[1:] or ["MultiRobotPuzzleHeavy-v0"]
    for env_id in ids:
        run(env_id, int(os.environ.get("QB_ENVS", 1048576)), n_agents=int(os.environ.get("QB_AGENTS", 0)))
