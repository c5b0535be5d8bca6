"""Diagnostic: where do GPU and oracle part ways on the square variant?  (first differing step per env, which words)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from gym_puzzles_b200 import abi
from oracle_lib import OracleBatch, StateView
from parity_util import split_state

SQ = sys.argv[1] if len(sys.argv) > 1 else "MultiRobotPuzzleSquare-v2"
N, T = 4096, 30
o = OracleBatch(SQ, N, seed=17, nthreads=8, max_episode_steps=50)
h = abi.Handle(SQ, N, seed=17, max_episode_steps=50)
o.reset(); h.reset_host()
first = {}
for t in range(T):
    a = o.sample_actions(t)
    o.step(a); h.step_host(a)
    wo, wh = o.get_state(), h.get_state()
    za = np.where(wo == 0x80000000, 0, wo); zb = np.where(wh == 0x80000000, 0, wh)
    bad = np.nonzero((za != zb).any(axis=1))[0]
    for e in bad:
        if e not in first:
            cols = np.nonzero(za[e] != zb[e])[0]
            first[e] = (t, cols)
L = o.layout
print(SQ, "envs that ever differ:", len(first), "of", N, "after", T, "steps")
names = [(0, "hdr"), (L.off_goal_contact, "goalc"), (L.off_bodies, "bodies"), (L.off_dists, "dists"), (L.off_goal, "goal"), (L.off_episode_acc, "epacc"), (L.off_aabb, "aabb"), (L.off_contacts, "contacts")]
def sect(c):
    nm = "hdr"
    for off, n in names:
        if c >= off: nm = n
    return nm
for e, (t, cols) in list(first.items())[:12]:
    sec = sorted({sect(c) for c in cols})
    det = ""
    if "contacts" in sec:
        cc = [c for c in cols if c >= L.off_contacts]
        det = " contact words " + str(sorted({(c - L.off_contacts) % 14 for c in cc}))
    print(f"env {e}: first differs at step {t}: {len(cols)} words in {sec}{det}")
