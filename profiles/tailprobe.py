"""Tail anatomy of the persistent kernels (profiling build with -DMRP_TAILPROBE, see profiles/build_probe.sh):
when do the warps of k_solve_vel / k_solve_pos / k_post_events finish, and how long did the longest single task run?
  bash profiles/build_probe.sh && MRP_LIB_PATH=$PWD/gym_puzzles_b200/csrc/libmrp_probe.so python profiles/tailprobe.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from gym_puzzles_b200 import abi

N = int(os.environ.get("QB_ENVS", 1048576))
env_id = sys.argv[1] if len(sys.argv) > 1 else "MultiRobotPuzzleHeavy-v0"
os.environ.setdefault("MRP_OVERLAP_POST", "0")   # no side-stream k_post: kernel boundaries are clean
h = abi.Handle(env_id, N, seed=17)
lib = h.lib.lib
h.reset()
for t in range(int(os.environ.get("QB_SETTLE", 100))):
    h.sample_actions(t)
    h.step()
torch.cuda.synchronize()
KW = 8192
buf = np.zeros((8, 2 + KW), dtype=np.uint64)
lib.mrp_debug_tailprobe.argtypes = [C.c_int, C.c_void_p]
names = ["k_solve_vel", "k_solve_pos", "k_post_events", "k_post_events(free)", "vel class 0", "vel class 1", "vel class 2", "vel class 3 / k_solve_big"]
for rep in range(3):
    lib.mrp_debug_tailprobe(0, None)
    h.sample_actions(1000 + rep)
    h.step()
    torch.cuda.synchronize()
    lib.mrp_debug_tailprobe(1, buf.ctypes.data_as(C.c_void_p))
    print(f"--- step {rep} ({env_id}, {N} envs)")
    for k, name in enumerate(names):
        st, mx, ends = int(buf[k, 0]), int(buf[k, 1]), buf[k, 2:]
        ends = ends[ends > 0].astype(np.int64)
        if len(ends) == 0 or st == 2 ** 64 - 1:
            continue
        rel = np.sort(ends - st) / 1e3
        q = lambda p: rel[min(len(rel) - 1, int(p * len(rel)))]   # noqa: E731
        info = mx & 0xffffff
        print(f"{name:20s} warps {len(rel):5d}  finish us: p10 {q(.10):7.1f} p50 {q(.5):7.1f} p90 {q(.9):7.1f} p99 {q(.99):7.1f} max {rel[-1]:7.1f}"
              f" | longest task {(mx >> 24) / 1e3:7.1f} us (T {info >> 16}, sweeps {info & 0xffff})")
    # event pass: per-env breakdown (SM cycles)
    rec = np.zeros((65536, 5), dtype=np.uint64)
    lib.mrp_debug_event_records.argtypes = [C.c_void_p, C.c_int]
    n = lib.mrp_debug_event_records(rec.ctypes.data_as(C.c_void_p), 65536)
    r = rec[:min(n, 65536)].astype(np.float64)
    if len(r):
        mhz = 1965.0
        ev = r[:, 4] > 0
        for name, sel in (("candidates only", ~ev), ("with events", ev)):
            x = r[sel]
            if len(x) == 0:
                continue
            order = np.argsort(x[:, 0])
            top = x[order[-max(1, len(x) // 100):]]
            print(f"  events pass, envs {name}: {len(x)}  mean total {x[:, 0].mean() / mhz:6.1f} us (TOI {x[:, 1].mean() / mhz:6.1f} us in {x[:, 3].mean():.1f} calls, "
                  f"toi_event {x[:, 2].mean() / mhz:6.1f} us in {x[:, 4].mean():.1f} events) | slowest 1 %: total {top[:, 0].mean() / mhz:6.1f} us, TOI {top[:, 1].mean() / mhz:6.1f} us "
                  f"in {top[:, 3].mean():.1f} calls, toi_event {top[:, 2].mean() / mhz:6.1f} us in {top[:, 4].mean():.1f} events")
h.close()
