"""The only collective of the path: episode statistics summed over ranks.  world_size-2 gloo run on CPU of the same
reduce_stats() the NCCL path uses, plus shard-invariant stepping (each rank owns a contiguous env-id block)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emu_lib import emu_lib
    from gym_puzzles_b200 import abi
    from gym_puzzles_b200.vector_env import reduce_stats, shard_range

    total = 64
    base, n = shard_range(total, rank, world)
    h = abi.Handle("MultiRobotPuzzle-v0", n, seed=17, env_id_base=base, max_episode_steps=10, lib=emu_lib())
    h.reset_host()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (25, total, 6)).astype(np.float32)
    obs = None
    for t in range(25):
        obs = h.step_host(acts[t, base:base + n])[0]
    local = h.stats()
    st = torch.tensor([local[k] for k in abi.STAT_NAMES], dtype=torch.float64)
    red = reduce_stats(torch, st, reduce_across_ranks=True, reset=True)
    assert float(st.sum()) == 0.0
    gathered = [None] * world
    dist.all_gather_object(gathered, obs)
    if rank == 0:
        q.put((red, local, np.concatenate(gathered)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stats_allreduce_and_shard_invariance():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    red, local0, obs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # 64 envs, cap 10, 25 steps => every env finished exactly 2 episodes (plus rare completions)
    assert red["episodes"] >= 128 and red["episodes"] >= 2 * local0["episodes"] - 4
    assert red["mean_length"] == pytest.approx(red["sum_length"] / red["episodes"])
    # single-process run over all 64 envs gives the same final observations
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu_lib import emu_lib
    from gym_puzzles_b200 import abi
    h = abi.Handle("MultiRobotPuzzle-v0", 64, seed=17, max_episode_steps=10, lib=emu_lib())
    h.reset_host()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (25, 64, 6)).astype(np.float32)
    for t in range(25):
        o = h.step_host(acts[t])[0]
    assert np.array_equal(o, obs)
    assert h.stats()["episodes"] == red["episodes"]
