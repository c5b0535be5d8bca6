#!/usr/bin/env python
"""bench.py — env-steps/sec of the MultiRobotPuzzle hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # native arm (sm_100a kernels through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU restatement on the host cores
  python bench.py --config c3|c3-strong|c3-resets|c2|c4|c4-heavy|c5 ...   # the other BASELINE.json configs (default c3)

Default workload (config.workload): MultiRobotPuzzleHeavy-v0 (5 robots, 2x block), 1,048,576 envs per GPU, random actions
U(-1,1) from the Philox ACTION stream, auto-reset on (BASELINE.json configs[2]; the registered TimeLimit of 3000
steps applies).  One "step" = one env.step of every env of the batch.

  value     whole-job env-steps/s with actions already resident in HBM (pre-generated ring of action buffers),
            K steps bracketed by barrier + synchronize, CUDA-event timed, max over ranks.  Production configuration of
            mrp_step: library timers off.
  e2e       the same through mrp_step_host(): pinned HOST action buffer -> H2D, step, obs/reward/done/trunc D2H,
            every step, copies inside the timed region.
  roofline  the binding roofline of this path is instruction issue, not HBM: `bound: "issue"` reports
            thread-level instructions per second against 148 SMs x 4 schedulers x 32 lanes x f_clk; `hbm` beside it holds
            the algorithmic bytes over the whole step and over the dominant phase, and the DRAM traffic
            ncu measured; `fp32` is the FLOP view from the device counters of the timed run (MRP_STAT_PAIRS / M1 / M2 / ...).
            Per-phase durations come from CUDA events recorded between the kernels inside the library
            (mrp_set_timing / mrp_get_phase_timing) in a separate pass of K steps.
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

METRIC = "env-steps/sec MultiRobotPuzzleHeavy-v0"
UNIT = "env-steps/s"

# BASELINE.json configs.  envs: per GPU ("weak") or in total, sharded over the ranks ("strong").
CONFIGS = {
    "c3": dict(env_id="MultiRobotPuzzleHeavy-v0", envs=1048576, scaling="weak", cap=0, n=5, A=15, O=40,
               what="configs[2]: Heavy-v0, 1M envs per GPU"),
    "c3-strong": dict(env_id="MultiRobotPuzzleHeavy-v0", envs=1048576, scaling="strong", cap=0, n=5, A=15, O=40,
                      what="configs[2]: Heavy-v0, 1M envs in total sharded over the GPUs"),
    "c3-resets": dict(env_id="MultiRobotPuzzleHeavy-v0", envs=1048576, scaling="weak", cap=200, n=5, A=15, O=40,
                      what="configs[2] with TimeLimit 200: the auto-reset path fires every step"),
    "c2": dict(env_id="MultiRobotPuzzle-v0", envs=65536, scaling="weak", cap=0, n=2, A=6, O=28,
               what="configs[1]: v0, 65,536 envs on one GPU"),
    "c4": dict(env_id="MultiRobotPuzzle-v2", envs=1048576, scaling="weak", cap=0, n=2, A=4, O=39,
               what="configs[3]: v2, 1M envs per GPU"),
    "c4-heavy": dict(env_id="MultiRobotPuzzleHeavy-v2", envs=1048576, scaling="weak", cap=0, n=2, A=4, O=39,
                     what="configs[3]: Heavy-v2, 1M envs per GPU"),
    # configs[4] "4M envs on 8 GPUs" = 524,288 per GPU; n counts the bodies besides the first block (2 robots + 2 more blocks)
    "c5": dict(env_id="MultiRobotPuzzleSquare-v2", envs=524288, scaling="weak", cap=0, n=4, A=4, O=72,
               what="configs[4]: 3-block square-forming variant (extension; oracle-defined semantics), Heavy-v2 dynamics, 524,288 envs per GPU"),
}


def algo_bytes(cfg):
    """algorithmic bytes per env-step, SURVEY.md §8d: action in + obs / reward / done out + minimal body state read + write"""
    return 4 * (cfg["A"] + cfg["O"] + 1) + 1 + 2 * 24 * (cfg["n"] + 1)


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


def cpu_baseline_run(env_id, n_envs, steps, warmup, nthreads, cap=0):
    """Oracle (CPU restatement of the reference path) on the host cores; returns env-steps/s.  Same action stream layout as the
    GPU arm: one distinct U(-1,1) buffer per step from a ring of up to 64."""
    from oracle_lib import OracleBatch

    o = OracleBatch(env_id, n_envs, seed=17, nthreads=nthreads, max_episode_steps=cap)
    o.reset()
    R = min(max(steps, 1), 64)
    acts = [o.sample_actions(t) for t in range(R)]
    for t in range(warmup):
        o.step(acts[t % R])
    t0 = time.perf_counter()
    for t in range(steps):
        o.step(acts[(warmup + t) % R])
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    n_envs = args.ref_envs
    v, dt = cpu_baseline_run(cfg["env_id"], n_envs, args.steps, max(args.warmup, 1), cores, cap=cfg["cap"])
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cfg['env_id']}, random actions U(-1,1), auto-reset; each step = one env.step of a bounded sample of "
                               f"{n_envs} envs on the host cores", "bench_config": args.config},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_envs} envs x {args.steps} steps, {cores} threads (pybox2d not installable: oracle/ C++ restatement, "
                                   "no Python/SWIG overhead => upper bound on the reference)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    GUARD.emit(json.dumps(line))


def _profile_tables(env_id):
    """committed ncu summaries (profiles/kernel_traffic.json, kernel_issue.json: `python profiles/summarize.py`), keyed by
    kernel name; they describe the MultiRobotPuzzleHeavy-v0 step and are used for that workload only"""
    out = {}
    for name in ("kernel_traffic", "kernel_issue"):
        path = os.path.join(ROOT, "profiles", name + ".json")
        try:
            rec = json.load(open(path))
        except Exception:
            rec = {}
        out[name] = {k.split("::")[-1]: v for k, v in rec.items()} if env_id == "MultiRobotPuzzleHeavy-v0" else {}
    return out["kernel_traffic"], out["kernel_issue"]


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
    from gym_puzzles_b200 import abi

    cfg = CONFIGS[args.config]
    env_id = cfg["env_id"]
    total = args.envs if args.envs > 0 else cfg["envs"]
    if cfg["scaling"] == "strong":
        if total % world:
            raise SystemExit(f"--config {args.config}: {total} envs do not divide over {world} ranks")
        N = total // world          # this rank's shard of the fixed total
    else:
        N = total                   # per GPU
    K, W = args.steps, max(args.warmup, 3)
    env = gp.VectorEnv(env_id, N, device=dev, seed=17, env_id_base=rank * N, max_episode_steps=cfg["cap"])
    h = env.handle
    A, O = h.act_dim, h.obs_dim
    assert (A, O) == (cfg["A"], cfg["O"])
    env.reset()
    # settle: all envs are reset at t=0, so the first steps resolve spawn overlaps (heavier than steady state);
    # run them untimed (actions sampled on the fly) before the W warm-up steps so the timed region sees the
    # rollout's stationary mix.  With a TimeLimit cap the episodes are de-phased first: during the first `cap` steps env i is
    # reset again at step i mod cap, so that from then on N / cap envs reach the limit and respawn in EVERY step (a steady
    # stream of auto-resets, as in training) instead of all at once every `cap` steps.
    settle = args.settle if not cfg["cap"] else max(args.settle, 2 * cfg["cap"] + 20)
    phase = torch.arange(N, device=dev) % cfg["cap"] if cfg["cap"] else None
    for t in range(settle + W):
        if cfg["cap"] and t < cfg["cap"]:
            env.reset((phase == t).to(torch.uint8))
        env.sample_actions(step_index=t)
        env.step()
    # inputs of the timed region are resident in HBM before it starts: one distinct pre-generated U(-1,1) action
    # buffer per timed step (up to 64, then the ring repeats; a short ring would give every robot a net drift)
    R = min(K, 64)
    acts = torch.empty((R, N, A), dtype=torch.float32, device=dev)
    for r in range(R):
        env.sample_actions(step_index=settle + W + r, out=acts[r])
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
    stats0 = env.stats_tensor.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for t in range(K):
        env.step(acts[t % R])
    e1.record()
    barrier()
    counters = (env.stats_tensor.clone() - stats0).tolist()    # workload counters of exactly the K timed steps (this rank)
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

    # the one collective of this path: episode statistics, 16 doubles summed over ranks (NCCL)
    stats = env.episode_stats(reduce_across_ranks=True, reset=False)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        hbm_peak = peaks.get("hbm_gbs")
        step_ms = elapsed_ms / K
        k_mean_ms = k_ms / max(k_cnt, 1)
        per_kernel = {k: v / max(k_cnt, 1) for k, v in phase_ms.items()}
        dom = max(per_kernel, key=per_kernel.get)
        dom_ms = per_kernel[dom]
        ab = algo_bytes(cfg)
        traffic_tab, issue_tab = _profile_tables(env_id)
        # DRAM traffic per step (dram__bytes_read.sum + dram__bytes_write.sum of every kernel of the committed ncu --set full
        # capture, scaled from the profiled batch size to this run's) and of the dominant phase alone
        traffic_step = sum(r["dram_bytes_per_launch"] * (N / r["envs"]) for r in traffic_tab.values()) or None
        dom_names = ["k_broad", "k_narrow", "k_pre", "k_front"] if dom == "k_pre" else [dom, dom + "#2"]
        traffic_dom = sum(traffic_tab[k]["dram_bytes_per_launch"] * (N / traffic_tab[k]["envs"]) for k in dom_names if k in traffic_tab) or None
        mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        # ---- issue roofline (binding): thread-level instructions per env-step (warp instructions x active lanes per instruction
        # of every kernel of the committed capture) x this run's env-steps/s against SMs x 4 schedulers x 32 lanes x f_clk
        tinst = sum(r["warp_inst_per_launch"] * r["active_lanes_per_inst"] / r["envs"] for r in issue_tab.values()) or None
        winst = sum(r["warp_inst_per_launch"] / r["envs"] for r in issue_tab.values()) or None
        peak_ti = sms * 4 * 32 * mhz * 1e6
        ach_ti = tinst * N / (step_ms / 1e3) if tinst else None
        # ---- FP32 roofline from the device counters of the timed run (SURVEY.md §8d / Appendix D cost model, 1 flop per
        # mul / add: the kernels are built -fmad=false, so the FP32 peak is SMs x 128 lanes x f_clk x 1)
        cn = dict(zip(abi.STAT_NAMES, counters))
        env_steps = float(N * K)
        f_fix = 30.0 * (cfg["n"] + 1) + 60.0 + 12.0 * cfg["n"]
        f_exec = f_fix * env_steps + 700.0 * cn["pairs"] + cn["vel_flops"] + 64.0 * cn["pos_points"] + 2000.0 * cn["toi_calls"]
        f_algo = f_fix * env_steps + 700.0 * cn["pairs"] + 180.0 * (81.0 * cn["m1"] + 160.0 * cn["m2"]) + 64.0 * cn["pos_points"] + 2000.0 * cn["toi_calls"]
        t_s = elapsed_ms / 1e3
        fp32_peak = sms * 128 * mhz * 1e6 / 1e12
        fp32 = {"executed_tflops": f_exec / t_s / 1e12, "algorithmic_tflops": f_algo / t_s / 1e12, "peak_tflops": fp32_peak,
                "frac": f_exec / t_s / 1e12 / fp32_peak, "algorithmic_frac": f_algo / t_s / 1e12 / fp32_peak,
                "flops_per_env_step_executed": f_exec / env_steps, "flops_per_env_step_algorithmic": f_algo / env_steps,
                "per_env_step": {k: cn[k] / env_steps for k in ("pairs", "m1", "m2", "vel_flops", "pos_points", "toi_calls")},
                "model": "F = F_fix + 700 P + vel + 64 pos_points + 2000 toi_calls; vel = executed sweeps (81 / 160 per contact) or 180 (81 m1 + 160 m2) "
                         "for the algorithmic figure (Box2D runs all 180; the kernels stop at the exact fixed point); peak = SMs x 128 x f_clk, no FMA"}
        hbm = {"algorithmic_bytes_per_env_step": ab, "step_achieved": ab * N / (step_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "step_frac": ab * N / (step_ms / 1e3) / 1e9 / hbm_peak if hbm_peak else None,
               "traffic_bytes_per_step": traffic_step, "traffic_bytes_per_env_step": traffic_step / N if traffic_step else None,
               "traffic_over_algorithmic": traffic_step / (ab * N) if traffic_step else None,
               "traffic_frac": traffic_step / (step_ms / 1e3) / 1e9 / hbm_peak if traffic_step and hbm_peak else None,
               "dominant_phase": {"kernel": "k_broad+k_narrow+k_pre" if dom == "k_pre" else dom, "ms": dom_ms,
                                  "share_of_step": dom_ms / k_mean_ms if k_mean_ms else None, "traffic_bytes": traffic_dom,
                                  "traffic_gbs": traffic_dom / (dom_ms / 1e3) / 1e9 if traffic_dom and dom_ms else None},
               "peak_source": peak_src}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{env_id}, {N} envs per GPU ({world * N} total), random actions U(-1,1) (Philox), auto-reset, "
                                   f"TimeLimit {h.layout.max_episode_steps}", "bench_config": args.config, "what": cfg["what"],
                       "envs_per_gpu": N, "settle_steps": settle, "parallelism": f"env-sharded x{world}, no data-path collective",
                       "l2": f"per-step working set (state {N * h.state_bytes / 1e9:.2f} GB + obs/actions {N * (A + O) * 4 / 1e9:.2f} GB per GPU) "
                             ">> 126 MB L2, no flush needed" if N * h.state_bytes > 4 * 126e6 else
                             f"per-step working set {N * (h.state_bytes + (A + O) * 4) / 1e6:.0f} MB per GPU: partly L2-resident between steps, as in use"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N * A * 4, "d2h_bytes_per_step": N * (O * 4 + 4 + 1 + 1),
                    "steps": Ke, "ms_per_step": float(e2e_ms.item()) / Ke},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "issue", "achieved": ach_ti, "peak": peak_ti, "unit": "thread-inst/s",
                         "frac": ach_ti / peak_ti if ach_ti else None, "traffic": traffic_step,
                         "thread_inst_per_env_step": tinst, "warp_inst_per_env_step": winst,
                         "active_lanes_per_inst": tinst / winst if tinst and winst else None, "sm_mhz": mhz,
                         "kernel": "whole step (phase pipeline)", "kernels_ms": per_kernel, "pipeline_ms": k_mean_ms,
                         "hbm": hbm, "fp32": fp32, "issue_from_ncu": issue_tab or None,
                         "note": "instruction issue binds this path (branchy FP32 geometry + a serial Gauss-Seidel solver), not HBM; "
                                 "thread-instruction counts and DRAM traffic come from the committed ncu capture of the same step "
                                 "(profiles/kernel_issue.json, kernel_traffic.json), durations and counters from this run"},
            "episode_stats": {k: stats[k] for k in ("episodes", "done_by_env", "truncated", "mean_return", "mean_length", "overflow", "nan_resets")},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, dt = cpu_baseline_run(env_id, args.ref_envs, args.cpu_steps, 2, cores, cap=cfg["cap"])
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
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS), help="BASELINE.json config (default c3 = configs[2], the headline)")
    ap.add_argument("--envs", type=int, default=0, help="override the config's env count (per GPU; in total for c3-strong)")
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
