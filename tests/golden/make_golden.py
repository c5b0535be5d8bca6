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

Since round 2 every file also holds single steps FROM GIVEN STATES (`ss_*` arrays): canonical state records
(include/mrp_state.h) loaded into the reference env (tests/refshim harness.ReferenceEnv.set_state), one env.step with
a recorded action, and what the reference returned.  They cover mid-rollout states with touching contacts and — v2 ids —
the termination branches no rollout reaches: robot / block centre of mass inside the 0.1 band along the world edge
(mrp02:279-295,552-563), without and with `update_params(7, 0.93)`.

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
import state_cases  # noqa: E402
from oracle_lib import OracleBatch  # noqa: E402  (only to pick interesting env ids / states; values come from the reference)

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
    ss = single_steps(env_id)
    out = os.path.join(HERE, env_id + ".npz")
    np.savez_compressed(out, seed=SEED, cap=CAP, gids=np.asarray(gids, np.int64), actions=actions, obs0=obs0, obs=obs, rew=rew,
                        done=done, trunc=trunc, contact=contact, bodies=bodies, **ss)
    print("   single steps from given states:", len(ss["ss_states"]), "of which out-of-bounds", int((ss["ss_kind"] > 0).sum()),
          "done", int(ss["ss_done"].sum()))
    print(env_id, "envs", G, "steps", T, "dones", int(done.sum()), "by env", int((done & ~trunc.astype(bool)).sum()),
          "contact steps", int(contact.any(axis=2).sum()), "->", os.path.relpath(out), os.path.getsize(out) // 1024, "KiB")


DECAY = (7, 0.93)   # update_params(timestep, decay) of the decayed out-of-bounds cases


def single_steps(env_id):
    """one reference env.step from each given state -> ss_* arrays (kind 0 rollout state, 1 robot OOB, 2 block OOB, 3 both;
    ss_decay_pow = decay ** (-timestep) in force, 1.0 = update_params(0, 1.0))"""
    v2 = env_id.endswith("v2")
    states = [state_cases.rollout_states(env_id, 12)]
    kinds = [0] * 12
    decay_pow = [1.0] * 12
    if v2:
        oob, names = state_cases.oob_states(env_id)
        code = {"agent": 1, "block": 2, "both": 3}
        for dp in (1.0, DECAY[1] ** (-DECAY[0])):
            states.append(oob)
            kinds += [code[k] for k in names]
            decay_pow += [dp] * len(oob)
    states = np.concatenate(states)
    o = OracleBatch(env_id, len(states), seed=5)
    lay = o.layout
    acts = o.sample_actions(3)
    o.close()
    S = len(states)
    n = lay.n_agents
    obs = np.zeros((S, lay.obs_dim))
    rew = np.zeros(S)
    done = np.zeros(S, np.uint8)
    contact = np.zeros((S, n), np.uint8)
    bodies = np.zeros((S, n + 1, 6), np.float32)
    r = harness.ReferenceEnv(env_id, seed=state_cases.SEED, gid=0)
    r.reset()
    for k in range(S):
        if v2:
            r.env.update_params(*(DECAY if decay_pow[k] != 1.0 else (0, 1.0)))
        r.set_state(lay, states[k])
        with r._feeds():
            ob, rw, dn, _ = r.env.step(acts[k].astype(np.float64))
        obs[k], rew[k], done[k] = np.asarray(ob, dtype=np.float64), float(rw), bool(dn)
        contact[k], bodies[k] = r.goal_contacts, r.body_rows()
    return dict(ss_states=states, ss_actions=acts, ss_kind=np.asarray(kinds, np.int32), ss_decay_pow=np.asarray(decay_pow),
                ss_obs=obs, ss_rew=rew, ss_done=done, ss_contact=contact, ss_bodies=bodies)


if __name__ == "__main__":
    if not harness.available():
        sys.exit("needs the reference checkout at " + harness.REFERENCE_ROOT)
    for env_id in harness.REGISTRY:
        run(env_id)
