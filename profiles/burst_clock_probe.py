"""Does a step run slower when the GPU idles between steps (as in the synchronous host-buffer loop)?
Device-resident mrp_step timed with CUDA events: back to back, then with a host sync + sleep between steps, with the SM clock
sampled through NVML during both.   python profiles/burst_test.py   (under gpurun)"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pynvml
import torch

from gym_puzzles_b200 import abi

N = int(os.environ.get("QB_ENVS", 1048576))
h = abi.Handle("MultiRobotPuzzleHeavy-v0", N, seed=17)
h.reset()
for t in range(100):
    h.sample_actions(t)
    h.step()
torch.cuda.synchronize()
pynvml.nvmlInit()
dev = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False


def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(dev) / 1e3))
        time.sleep(0.002)


def run(gap_s, steps=60):
    global samples, stop
    samples, stop = [], False
    th = threading.Thread(target=sampler)
    th.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for t in range(steps):
        h.sample_actions(500 + t)
        ev[t][0].record()
        h.step()
        ev[t][1].record()
        if gap_s is not None:
            torch.cuda.synchronize()
            time.sleep(gap_s)
    torch.cuda.synchronize()
    stop = True
    th.join()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    clk = sorted(s[0] for s in samples)
    pw = sorted(s[1] for s in samples)
    print(f"gap {gap_s}: step ms median {ms[len(ms) // 2]:.3f} min {ms[0]:.3f} max {ms[-1]:.3f} | sm clock median {clk[len(clk) // 2]} min {clk[0]} "
          f"| power median {pw[len(pw) // 2]:.0f} W max {pw[-1]:.0f} W", flush=True)


run(None)
run(0.0)
run(0.003)
run(0.010)
run(None)
