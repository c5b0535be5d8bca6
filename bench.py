#!/usr/bin/env python
"""bench.py — env-steps/sec of the MultiRobotPuzzle hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # native arm (sm_100a kernels through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU restatement on the host cores

Workload (config.workload): MultiRobotPuzzleHeavy-v0 (5 robots, 2x block), 1,048,576 envs per GPU, random actions
U(-1,1) from the Philox ACTION stream, auto-reset on (BASELINE.json configs[2]; the registered TimeLimit of 3000
steps applies).  One "step" = one env.step of every env of the batch.

  value     whole-job env-steps/s with actions already resident in HBM (pre-generated ring of action buffers),
            K steps bracketed by barrier + synchronize, CUDA-event timed, max over ranks.  Production configuration of
            mrp_step: library timers off, k_post of the envs without solver tasks on a side stream beside the solvers.
  e2e       the same through mrp_step_host(): pinned HOST action buffer -> H2D, step, obs/reward/done/trunc D2H,
            every step, copies inside the timed region.
  roofline  dominant phase of the step's pipeline (k_broad+k_narrow+k_pre / k_solve_vel / k_solve_pos / k_post / k_post_events):
            algorithmic bytes (513 B per env-step, SURVEY.md §8d / DESIGN.md) over its mean launch time, measured with
            CUDA events recorded between the kernels inside the library (mrp_set_timing / mrp_get_phase_timing) in a
            separate pass of K steps (the timers keep every kernel on the launching stream, one after the other).
  cpu_baseline  the oracle ("port": pybox2d is not installable here) on all host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ENV_ID = "MultiRobotPuzzleHeavy-v0"
ALGO_BYTES_PER_ENV_STEP = 513  # 4*(A+O+1)+1 + 2*24*(n+1) with A=15, O=40, n=5
METRIC = "env-steps/sec MultiRobotPuzzleHeavy-v0"
UNIT = "env-steps/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_run(n_envs, steps, warmup, nthreads):
    """Oracle (CPU restatement of the reference path) on the host cores; returns env-steps/s."""
    from oracle_lib import OracleBatch

    o = OracleBatch(ENV_ID, n_envs, seed=17, nthreads=nthreads)
    o.reset()
    acts = [o.sample_actions(t) for t in range(4)]
    for t in range(warmup):
        o.step(acts[t % 4])
    t0 = time.perf_counter()
    for t in range(steps):
        o.step(acts[t % 4])
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = args.ref_envs
    v, dt = cpu_baseline_run(n_envs, args.steps, max(args.warmup, 1), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{ENV_ID}, random actions U(-1,1), auto-reset; each step = one env.step of a bounded sample of "
                               f"{n_envs} envs on the host cores"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_envs} envs x {args.steps} steps, {cores} threads (pybox2d not installable: oracle/ C++ restatement, "
                                   "no Python/SWIG overhead => upper bound on the reference)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    GUARD.emit(json.dumps(line))


def run_native(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group(backend="nccl", device_id=dev)

    import gym_puzzles_b200 as gp

    N = args.envs
    K, W = args.steps, max(args.warmup, 3)
    env = gp.VectorEnv(ENV_ID, N, device=dev, seed=17, env_id_base=rank * N)
    h = env.handle
    A, O = h.act_dim, h.obs_dim
    env.reset()
    # settle: all envs are reset at t=0, so the first steps resolve spawn overlaps (heavier than steady state);
    # run them untimed (actions sampled on the fly) before the W warm-up steps so the timed region sees the
    # rollout's stationary mix
    for t in range(args.settle + W):
        env.sample_actions(step_index=t)
        env.step()
    # inputs of the timed region are resident in HBM before it starts: one distinct pre-generated U(-1,1) action
    # buffer per timed step (up to 64, then the ring repeats; a short ring would give every robot a net drift)
    R = min(K, 64)
    acts = torch.empty((R, N, A), dtype=torch.float32, device=dev)
    for r in range(R):
        env.sample_actions(step_index=args.settle + W + r, out=acts[r])
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = None
    if rank == 0:
        try:
            gpu_sel = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", "")
        except Exception:
            gpu_sel = str(local_rank)
        sampler = ClockSampler(gpu_sel)
    # ---------------- device-resident throughput (library timers off: production configuration)
    launches0 = h.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for t in range(K):
        env.step(acts[t % R])
    e1.record()
    barrier()
    elapsed_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed_ms.item())
    launches = h.launch_count - launches0
    value = world * N * K / (elapsed_ms / 1e3)

    # ---------------- per-kernel durations for the roofline: the same steps again with the library's CUDA-event timers
    # on (events between the phase kernels on the launching stream; the timers force a single chunk / single stream)
    h.set_timing(True)
    h.get_timing(reset_after=True)
    h.get_phase_timing(reset_after=True)
    for t in range(K):
        env.step(acts[(K + t) % R])
    torch.cuda.synchronize()
    k_ms, k_cnt = h.get_timing(reset_after=True)
    phase_ms = h.get_phase_timing(reset_after=True)
    h.set_timing(False)

    # ---------------- end to end through the host-buffer C-ABI call (pinned host memory)
    Ke = max(2, min(K, args.e2e_steps))
    host_acts = [acts[(2 * K + i) % R].cpu().pin_memory() for i in range(Ke + 1)]   # a different action batch per step
    host_obs = torch.empty((N, O), dtype=torch.float32).pin_memory()
    host_rew = torch.empty((N,), dtype=torch.float32).pin_memory()
    host_done = torch.empty((N,), dtype=torch.uint8).pin_memory()
    host_trunc = torch.empty((N,), dtype=torch.uint8).pin_memory()
    np_acts = [x.numpy() for x in host_acts]
    np_args = [x.numpy() for x in (host_obs, host_rew, host_done, host_trunc)]
    h.step_host(np_acts[Ke], *np_args)  # warm
    barrier()
    e0.record()
    for t in range(Ke):
        h.step_host(np_acts[t], *np_args)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * N * Ke / (float(e2e_ms.item()) / 1e3)
    clocks = sampler.stop() if sampler is not None else None

    # the one collective of this path: episode statistics, 8 doubles summed over ranks (NCCL)
    stats = env.episode_stats(reduce_across_ranks=True, reset=False)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        k_mean_ms = k_ms / max(k_cnt, 1)
        per_kernel = {k: v / max(k_cnt, 1) for k, v in phase_ms.items()}
        dom = max(per_kernel, key=per_kernel.get)
        dom_ms = per_kernel[dom]
        achieved = ALGO_BYTES_PER_ENV_STEP * N / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else None
        pipeline_gbs = ALGO_BYTES_PER_ENV_STEP * N / (k_mean_ms / 1e3) / 1e9 if k_mean_ms > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if os.path.exists(tp):
            try:
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture,
                # scaled from the profiled batch size to this run's
                # the "k_pre" timer spans the three collide / setup launches k_broad + k_narrow + k_pre
                recs = json.load(open(tp))
                names = ["k_broad", "k_narrow", "k_pre"] if dom == "k_pre" else [dom, dom + "#2"]
                traffic = sum(recs[k]["dram_bytes_per_launch"] * (N / recs[k]["envs"]) for k in names if k in recs)
            except Exception:
                traffic = None
        # FP32-issue view of the same kernels (north star: "fraction of the FP32 and HBM roofline"): issue-slot and FMA-pipe
        # utilisation, active lanes per instruction and the leading stall reasons from the committed ncu --set full capture
        issue = None
        ip = os.path.join(ROOT, "profiles", "kernel_issue.json")
        if os.path.exists(ip):
            try:
                rec = json.load(open(ip))
                issue = {k: rec[k] for k in (["k_broad", "k_narrow", "k_pre"] if dom == "k_pre" else [dom, dom + "#2"]) if k in rec}
            except Exception:
                issue = None
        # Lane-issue roofline of the whole step: thread-level instructions per env-step (warp instructions x active lanes per
        # instruction of every kernel, from the committed ncu capture) x this run's env-steps/s, against
        # 148 SMs x 4 schedulers x 32 lanes x the SM clock sampled during the timed region.
        lane_issue = None
        try:
            rec = json.load(open(ip))
            tinst = sum(r["warp_inst_per_launch"] * r["active_lanes_per_inst"] / r["envs"] for r in rec.values())
            mhz = (clocks or {}).get("sm_mhz") or 1965.0
            peak_ti = 148 * 4 * 32 * mhz * 1e6
            ach_ti = tinst * N / (k_mean_ms / 1e3) if k_mean_ms > 0 else None
            lane_issue = {"thread_inst_per_env_step": tinst, "achieved": ach_ti, "peak": peak_ti, "unit": "thread-inst/s",
                          "frac": ach_ti / peak_ti if ach_ti else None, "sm_mhz": mhz}
        except Exception:
            lane_issue = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{ENV_ID}, {N} envs per GPU ({world * N} total), random actions U(-1,1) (Philox), auto-reset, "
                                   "TimeLimit 3000", "envs_per_gpu": N, "settle_steps": args.settle, "parallelism": f"env-sharded x{world}, no data-path collective",
                       "l2": "per-step working set (state 2.3 GB + obs/actions 0.23 GB per GPU) >> 126 MB L2, no flush needed"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N * A * 4, "d2h_bytes_per_step": N * (O * 4 + 4 + 1 + 1),
                    "steps": Ke, "ms_per_step": float(e2e_ms.item()) / Ke},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                         "frac": (achieved / peaks["hbm_gbs"]) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k_broad+k_narrow+k_pre" if dom == "k_pre" else dom, "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / k_mean_ms if k_mean_ms else None,
                         "kernels_ms": per_kernel, "pipeline_ms": k_mean_ms, "pipeline_achieved": pipeline_gbs,
                         "fp32_issue_from_ncu": issue, "lane_issue": lane_issue,
                         "note": "path is FP32-issue/latency bound by nature (SURVEY.md §8d): HBM fraction is expected << 1"},
            "episode_stats": {k: stats[k] for k in ("episodes", "done_by_env", "truncated", "mean_return", "mean_length", "overflow")},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, dt = cpu_baseline_run(args.ref_envs, args.cpu_steps, 2, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.ref_envs} envs x {args.cpu_steps} steps of the same workload, {cores} threads, {dt:.1f} s"}
        GUARD.emit(json.dumps(line))
    env.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


class StdoutGuard:
    """Everything libraries write to fd 1 while the benchmark runs (e.g. NCCL's version banner) goes to stderr, so that
    stdout carries exactly ONE line: the JSON result printed through emit()."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)


GUARD = None


def main():
    global GUARD
    GUARD = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--envs", type=int, default=1048576, help="envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--ref-envs", type=int, default=32768, help="bounded sample of the workload for the CPU arm")
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--settle", type=int, default=100, help="untimed steps after the initial reset, before warm-up")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
