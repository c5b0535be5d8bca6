"""Between-steps states (canonical records, include/mrp_state.h) that the parity tests and tests/golden/make_golden.py
load into the reference's Python (tests/refshim), the oracle, the host build of the kernel source and the sm_100a
library alike — BASELINE.json north_star: "checked ... from identical states".

Two families:
 * rollout states: what an oracle rollout looks like after t steps (contacts, warm-start impulses, stale fat AABBs);
 * v2 out-of-bounds states (mrp02:279-295,552-563): a robot's / the block's centre of mass inside the 0.1 band along
   the world edge.  The walls (half-thickness BOUNDS = 0.1, mrp02:394-411) cover that band, so no rollout ever gets
   there: the branches are reachable only from a constructed state.
"""
import numpy as np

from oracle_lib import OracleBatch, StateView

SEED = 29


def rollout_states(env_id, count, n_agents=0, steps=(3, 11, 24, 40, 57), envs=16, cap=45):
    """`count` records taken from an oracle rollout of `envs` envs at the given step numbers (round robin)."""
    o = OracleBatch(env_id, envs, seed=SEED, nthreads=4, max_episode_steps=cap, n_agents=n_agents)
    o.reset()
    out, t = [], 0
    for stop in steps:
        while t < stop:
            o.step(o.sample_actions(t))
            t += 1
        w = o.get_state()
        nc = StateView(o.layout, w).n_contacts
        order = np.argsort(-nc, kind="stable")          # contact-rich envs first
        out.extend(w[i].copy() for i in order[:max(1, count // len(steps) + 1)])
    o.close()
    return np.stack(out[:count])


def _band_positions(W, H):
    """centre-of-mass positions inside the 0.1 band, one per world edge"""
    return [(0.05, 0.45 * H), (W - 0.04, 0.6 * H), (0.4 * W, 0.03), (0.55 * W, H - 0.06)]


def oob_states(env_id, n_agents=0):
    """-> (records, kinds): for every world edge one state with a robot inside the band (kind 'agent'), one with the
    block inside it ('block') and one with both ('both': the robot branch wins, mrp02:552-563).  Built from a settled
    rollout state so that the other bodies, the goal and the previous distances are ordinary."""
    assert env_id.endswith("v2")
    base = rollout_states(env_id, 1, n_agents=n_agents, steps=(6,), envs=4)[0]
    o = OracleBatch(env_id, 1, n_agents=n_agents)
    l = o.layout
    o.close()
    W, H = 1440 / 560.0, 810 / 560.0          # VIEWPORT / SCALE, mrp02:40-43
    recs, kinds = [], []
    for edge, (x, y) in enumerate(_band_positions(W, H)):
        for kind in ("agent", "block", "both"):
            w = base.copy()
            w[3] = 0                                              # no contacts carried over (bodies are teleported)
            w[l.off_contacts:] = 0
            w[l.off_goal_contact:l.off_goal_contact + l.n_agents] = 0
            f = w[l.off_bodies:l.off_bodies + 6 * l.n_dyn_bodies].view(np.float32).reshape(-1, 6)
            f[:, 3:] = 0.0
            if kind in ("agent", "both"):
                f[1 + edge % l.n_agents, 0:2] = (x, y)
            if kind in ("block", "both"):
                f[0, 0:2] = (x, y) if kind == "block" else (W - x, H - y)
            # previous distances as reset() / the previous step would have left them for these poses (mrp02:263-277)
            ratio = 560.0 / 1440.0
            goal = np.ascontiguousarray(w[l.off_goal:l.off_goal + 4]).view(np.float64)
            c = f[:, 0:2].astype(np.float64) * ratio
            d = [np.hypot(*(c[1 + i] - c[0])) for i in range(l.n_agents)] + [np.hypot(*(c[0] - goal))]
            w[l.off_dists:l.off_dists + 2 * (l.n_agents + 1)] = np.asarray(d, dtype=np.float64).view(np.uint32)
            recs.append(w)
            kinds.append(kind)
    return np.stack(recs), kinds
