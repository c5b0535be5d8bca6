"""HBM roofline of the VecNormalize passes (csrc/mrp_vecnorm.cu): CUDA-event timing of each kernel alone.
  python profiles/bench_vecnorm.py [N] [O]      default 4,194,304 envs x 40 obs (671 MB per obs tensor >> 126 MB L2)"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from gym_puzzles_b200.vec_normalize import VecNormHandle

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4 << 20
O = int(sys.argv[2]) if len(sys.argv) > 2 else 40
peak = 6551.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
vn = VecNormHandle(N, O, device=0)
obs = torch.randn(N, O, device="cuda") * 50 + 100
out = torch.empty_like(obs)
rew = torch.randn(N, device="cuda")
rout = torch.empty_like(rew)
done = (torch.rand(N, device="cuda") < 0.01).to(torch.uint8)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_m = timed(lambda: vn.moments(obs.data_ptr(), rew.data_ptr()))
t_a = timed(lambda: vn.apply(obs.data_ptr(), rew.data_ptr(), done.data_ptr(), out.data_ptr(), rout.data_ptr()))
t_copy = timed(lambda: out.copy_(obs))
b_m = N * (4 * O + 4 + 16)            # obs read; reward read, returns read+write (f64)
b_a = N * (8 * O + 4 + 4 + 1)         # obs read + write; reward read + write; done read
res = {"N": N, "O": O, "peak_gbs": peak,
       "k_vn_moments": {"ms": t_m, "algorithmic_bytes": b_m, "gbs": b_m / t_m / 1e6, "frac": b_m / t_m / 1e6 / peak},
       "k_vn_merge+k_vn_apply": {"ms": t_a, "algorithmic_bytes": b_a, "gbs": b_a / t_a / 1e6, "frac": b_a / t_a / 1e6 / peak},
       "torch_copy_same_tensor": {"ms": t_copy, "gbs": 8 * N * O / t_copy / 1e6}}
print(json.dumps(res))
