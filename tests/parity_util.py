"""Shared parity harness: drives an mrp C-ABI handle and the CPU oracle from identical states / seeds and compares
contact flags, contact-point counts, done flags (bit-exact) and float state / obs / reward (tolerance)."""
import numpy as np

from oracle_lib import OracleBatch, StateView

RTOL = 1e-5   # BASELINE.json north_star: "within a stated float32 tolerance (e.g. 1e-5 relative per step)"
ATOL = 1e-6


def split_state(layout, words):
    """-> integer ('exact') and float ('tol') views of canonical state words"""
    sv = StateView(layout, words)
    l = layout
    w = sv.w
    con = sv.contacts
    ints = np.concatenate([w[:, 0:4], w[:, l.off_goal_contact:l.off_goal_contact + 8],
                           w[:, l.off_episode_acc + 2:l.off_episode_acc + 3],
                           con[:, :, 0].reshape(len(w), -1), con[:, :, 1].reshape(len(w), -1)], axis=1)
    f32 = np.concatenate([sv.bodies.reshape(len(w), -1), sv.aabbs.reshape(len(w), -1),
                          np.ascontiguousarray(con[:, :, 2:]).reshape(len(w), -1).view(np.float32)], axis=1)
    f64 = np.concatenate([sv.dists, np.ascontiguousarray(w[:, l.off_goal:l.off_goal + 4]).view(np.float64),
                          np.ascontiguousarray(w[:, l.off_episode_acc:l.off_episode_acc + 2]).view(np.float64)], axis=1)
    return ints, f32, f64


def compare_states(layout, wa, wb):
    """returns (n_int_mismatch_envs, n_float_out_of_tol_envs, n_envs_not_bit_identical, max_rel_err)"""
    ia, fa, da = split_state(layout, wa)
    ib, fb, db = split_state(layout, wb)
    int_bad = int((ia != ib).any(axis=1).sum())
    okf = np.isclose(fa, fb, rtol=RTOL, atol=ATOL)
    okd = np.isclose(da, db, rtol=RTOL, atol=ATOL)
    tol_bad = int(((~okf).any(axis=1) | (~okd).any(axis=1)).sum())
    # bitwise comparison up to the sign of zero: the kernels stop the 180 velocity sweeps once a sweep changes no
    # value, after which Box2D's remaining sweeps can still flip the sign of a zero-valued impulse (-0.0 vs +0.0);
    # +-0 never influences a non-zero value (no division / atan2 on these quantities), see DESIGN.md "early exit"
    za = np.where(wa == 0x80000000, 0, wa)
    zb = np.where(wb == 0x80000000, 0, wb)
    bit_bad = int((za != zb).any(axis=1).sum())
    denom = np.maximum(np.abs(fa), 1e-3)
    max_rel = float(np.max(np.abs(fa - fb) / denom)) if fa.size else 0.0
    return int_bad, tol_bad, bit_bad, max_rel


def device_stepper(handle):
    """step(actions) through the device-resident entry point mrp_step (actions uploaded into the library's action
    buffer, results read back from its device buffers): the path VectorEnv.step and bench.py's `value` use."""
    import torch

    from gym_puzzles_b200.vector_env import _wrap

    dev = torch.device("cuda", 0)
    b, N = handle.buffers, handle.num_envs
    act = _wrap(torch, b.action_dev, (N, handle.act_dim), "<f4", handle, dev)
    obs = _wrap(torch, b.obs_dev, (N, handle.obs_dim), "<f4", handle, dev)
    rew = _wrap(torch, b.reward_dev, (N,), "<f4", handle, dev)
    done = _wrap(torch, b.done_dev, (N,), "|u1", handle, dev)
    trunc = _wrap(torch, b.trunc_dev, (N,), "|u1", handle, dev)

    def step(a):
        act.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)))
        handle.step(None, torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize()
        return obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), trunc.cpu().numpy()

    return step


def rollout_compare(handle, variant, N, T, seed, max_episode_steps=0, n_agents=0, env_id_base=0, nthreads=8, state_every=25,
                    device_path=False, step_fn=None):
    """handle: gym_puzzles_b200.abi.Handle built with the same (variant, N, seed, ...).  device_path: step through
    mrp_step (device-resident) instead of mrp_step_host."""
    if step_fn is None:
        step_fn = device_stepper(handle) if device_path else handle.step_host
    o = OracleBatch(variant, N, seed=seed, nthreads=nthreads, max_episode_steps=max_episode_steps, n_agents=n_agents,
                    env_id_base=env_id_base)
    rep = dict(steps=0, flag_mismatch=0, done_mismatch=0, obs_not_close=0, obs_not_exact=0, rew_not_close=0,
               state_tol_bad=0, state_bit_bad=0, max_rel=0.0, dones=0)
    oo, ho = o.reset(), handle.reset_host()
    rep["obs_not_exact"] += int((oo.astype(np.float32) != ho).any(axis=1).sum())
    for t in range(T):
        a = o.sample_actions(t)
        obs_o, r_o, d_o, t_o = o.step(a)
        obs_h, r_h, d_h, t_h = step_fn(a)
        rep["steps"] += N
        rep["dones"] += int(d_o.sum())
        rep["done_mismatch"] += int((d_o != d_h).sum() + (t_o != t_h).sum())
        o32 = obs_o.astype(np.float32)
        rep["obs_not_exact"] += int((o32 != obs_h).any(axis=1).sum())
        rep["obs_not_close"] += int((~np.isclose(o32, obs_h, rtol=RTOL, atol=1e-4)).any(axis=1).sum())
        rep["rew_not_close"] += int((~np.isclose(r_o.astype(np.float32), r_h, rtol=1e-4, atol=1e-3)).sum())
        if t % state_every == state_every - 1 or t == T - 1:
            ib, tb, bb, mr = compare_states(o.layout, o.get_state(), handle.get_state())
            rep["flag_mismatch"] += ib
            rep["state_tol_bad"] += tb
            rep["state_bit_bad"] += bb
            rep["max_rel"] = max(rep["max_rel"], mr)
    rep["oracle_stats"] = o.stats()
    return rep


def single_step_compare(handle, oracle, states, actions):
    """Load identical states into both, one step, compare (the north star's one-step equivalence)."""
    oracle.set_state(states)
    handle.set_state(states)
    obs_o, r_o, d_o, t_o = oracle.step(actions)
    obs_h, r_h, d_h, t_h = handle.step_host(actions)
    ib, tb, bb, mr = compare_states(oracle.layout, oracle.get_state(), handle.get_state())
    o32 = obs_o.astype(np.float32)
    return dict(flag_mismatch=ib, state_tol_bad=tb, state_bit_bad=bb, max_rel=mr,
                done_mismatch=int((d_o != d_h).sum() + (t_o != t_h).sum()),
                obs_not_exact=int((o32 != obs_h).any(axis=1).sum()),
                obs_not_close=int((~np.isclose(o32, obs_h, rtol=RTOL, atol=1e-4)).any(axis=1).sum()),
                rew_not_close=int((~np.isclose(r_o.astype(np.float32), r_h, rtol=1e-4, atol=1e-3)).sum()))
