#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN PYTHON env code (imported unmodified from
/root/reference) over the Box2D stand-in of tests/refshim (the oracle's Box2D restatement behind pybox2d's API).

  python tests/golden/make_golden.py            # build container only: needs /root/reference

Each file holds, for one registered id, G independent envs (global env ids `gids`) driven for T steps with recorded
U(-1,1) float32 actions: the observation / reward / done / truncation the reference returned (float64 as the reference
computes them), its per-robot goal_contact flags and its dynamic-body state after every step.  Random draws
(np.random.uniform spawns, action_space.sample() hidden reset action) come from the Philox streams keyed by
(seed, gid, episode) that the oracle and the product use, so a replay needs nothing but seed, gids and the actions.
TimeLimit is `cap` steps (instead of the registered 2000/3000) so that resets are frequent.

What the fixtures pin: the env logic of mrp00 / mrp02 as executed by the reference itself.  What they do not pin:
Box2D's arithmetic (a restatement on both sides) — see DESIGN.md §2.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "refshim"))

import harness  # noqa: E402
from oracle_lib import OracleBatch  # noqa: E402  (only to pick interesting env ids; values come from the reference)

SEED = 17
# (steps, TimeLimit cap) per family.  The v2 robots spawn ~0.8 m left of the block facing it and crawl at <= 6 mm / step
# (mrp02:355-361,54): they need ~150 forward steps to reach it, so v2 uses a longer horizon and a driving action phase.
PLAN = {False: (120, 50), True: (300, 260)}


def interesting_gids(env_id, want_done):
    """a contiguous block of env ids plus (v0 family) a few whose block spawns inside the goal tolerance, so the
    completion branch (+10 +10000, mrp00:512-519) is part of the fixtures"""
    gids = list(range(6))
    if want_done:
        o = OracleBatch(env_id, 4096, seed=SEED, env_id_base=1000, nthreads=4)
        o.reset()
        _, _, done, _ = o.step(o.sample_actions(0))
        gids += [1000 + int(i) for i in np.nonzero(done)[0][:2]]
    return gids


def run(env_id):
    gids = interesting_gids(env_id, env_id.endswith("v0"))
    v2 = env_id.endswith("v2")
    T, CAP = PLAN[v2]
    envs = [harness.ReferenceEnv(env_id, seed=SEED, gid=g, max_episode_steps=CAP) for g in gids]
    G = len(envs)
    obs0 = np.stack([e.reset() for e in envs])
    n, O = len(envs[0].env.agents), obs0.shape[1]
    A = envs[0].env.action_space.shape[0]
    rng = np.random.default_rng(1234)
    actions = rng.uniform(-1, 1, (T, G, A)).astype(np.float32)
    if v2:   # (turn, vel) per robot: drive forward with gentle steering until the block is reached, then random
        actions[:230, :, 0::2] *= 0.3
        actions[:230, :, 1::2] = 0.85 + 0.15 * actions[:230, :, 1::2]
    else:    # a stretch of constant actions drives robots into walls / the block (contacts, TOI events)
        actions[20:45] = actions[20]
    obs = np.zeros((T, G, O))
    rew = np.zeros((T, G))
    done = np.zeros((T, G), np.uint8)
    trunc = np.zeros((T, G), np.uint8)
    contact = np.zeros((T, G, n), np.uint8)
    bodies = np.zeros((T, G, n + 1, 6), np.float32)
    for t in range(T):
        for g, e in enumerate(envs):
            obs[t, g], rew[t, g], done[t, g], trunc[t, g] = e.step(actions[t, g])
            contact[t, g] = e.goal_contacts
            bodies[t, g] = e.body_rows()
    out = os.path.join(HERE, env_id + ".npz")
    np.savez_compressed(out, seed=SEED, cap=CAP, gids=np.asarray(gids, np.int64), actions=actions, obs0=obs0, obs=obs, rew=rew,
                        done=done, trunc=trunc, contact=contact, bodies=bodies)
    print(env_id, "envs", G, "steps", T, "dones", int(done.sum()), "by env", int((done & ~trunc.astype(bool)).sum()),
          "contact steps", int(contact.any(axis=2).sum()), "->", os.path.relpath(out), os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    if not harness.available():
        sys.exit("needs the reference checkout at " + harness.REFERENCE_ROOT)
    for env_id in harness.REGISTRY:
        run(env_id)
