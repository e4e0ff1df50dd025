#!/usr/bin/env python
"""Benchmark of the LTM timestep (BASELINE.json metric: link-timesteps/sec, fp64 state).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid SIZE] [--no-env] [--no-cpu-baseline]

Workload at N=1 (`config.workload`): BASELINE config 4, the synthetic 512x512-node lattice of
data/create_grid.py's rule (1 046 528 directed links, default_link of data/45_intersections, 35
origins = 4 corners + every 64th boundary node, gaussian-peak Poisson demand pre-drawn on the host (lattice),
uniform turning fractions), single replica, on-device Philox draws.  configs[1] (nine_intersections,
24 links) is a parity-test case: at 24 links a step is pure launch latency and says nothing about the
HBM roofline the metric is quoted against; the 512x512 grid is the largest single-GPU configuration.
A "step" is one `network_loading(t)` over the whole network.  With N>1 every rank runs an
independent replica of the grid (replicas only, no data-path collective; weak scaling).

One JSON line on stdout (rank 0).  See the tier contract in the task brief for the keys.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_ALG = 176.0            # algorithmic bytes per link-timestep, SURVEY.md section 8(d)
# split of B_ALG by kernel (DESIGN.md "kernels"; sums to 176): the fused link-pair kernel carries the
# 96 B of lagged / previous-row reads and 48 B of link-state + sending/receiving writes, the node kernel
# the 32 B of flow / cumulative-count writes.  Intermediate re-reads between kernels are not algorithmic.
B_ALG_PASS = {"link_pair": 144.0, "route_probs": 0.0, "node_flows": 32.0}
GRID_SIZE = 512
REF_SAMPLE_SIZE = 32


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this build
    (profiles/r1_final_traffic.json; measured offline -- never under this run), or None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r1_final_traffic.json")))
        return rec[{"link_pair": "k_link_lane", "node_flows": "k_node_flows"}[kernel]]["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # first calls are slow (tens of ms): pay for them before the timed region starts
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=1)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------------- reference arm
def oracle_sample(steps, warmup):
    """The reference's algorithm on the host: the Python oracle (kind 'port'; the reference itself is
    Python and does not travel to the GPU box) on a 32x32 lattice with the workload's link
    parameters, origin rule and demand pattern.  Per-link-step cost of this code is size-independent
    (12.2k vs 12.7k link-steps/s at 16x16 / 32x32)."""
    import numpy as np
    from oracle.ltm_oracle import LtmOracle
    from pednstream_b200 import Network
    from pednstream_b200.grid import DEFAULT_LINK, default_origins, grid_adjacency
    size = REF_SAMPLE_SIZE
    origins = default_origins(size, stride=64)
    S = max(1000, warmup + steps + 1)
    params = {"unit_time": 10, "simulation_steps": S, "default_link": dict(DEFAULT_LINK),
              "demand": {f"origin_{o}": {"peak_lambda": 50, "base_lambda": 30} for o in origins}}
    np.random.seed(0)
    net = Network(grid_adjacency(size), params, origin_nodes=list(origins), verbose=False)
    o = LtmOracle(net)
    for t in range(1, warmup + 1):
        o.network_loading(t)
    t0 = time.perf_counter()
    for t in range(warmup + 1, warmup + steps + 1):
        o.network_loading(t)
    dt = time.perf_counter() - t0
    L = len(net.links)
    return L * steps / dt, dt, L


def _oracle_worker(job):
    steps, warmup = job
    t0 = time.perf_counter()
    value, dt, L = oracle_sample(steps, warmup)
    return L * steps, dt, L, time.perf_counter() - t0


def oracle_sample_all_cores(steps, warmup):
    """The reference algorithm is single-threaded; its only parallelism is independent replicas
    (the reference's RLlib workers), so the best case on the host is one replica per core."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    with mp.get_context("spawn").Pool(cores) as pool:
        out = pool.map(_oracle_worker, [(steps, warmup)] * cores)
    work = sum(o[0] for o in out)
    slowest = max(o[1] for o in out)
    return work / slowest, slowest, out[0][2], cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 60)
    value, dt, L, cores = oracle_sample_all_cores(steps, min(args.warmup, 3))
    sample = (f"{cores} independent {REF_SAMPLE_SIZE}x{REF_SAMPLE_SIZE} lattices ({L} links each), {steps} steps, same "
              f"link parameters / origin rule / demand pattern as the workload; one Python process per host core")
    line = {"impl": "reference", "metric": "link-timesteps/sec", "value": value, "unit": "link-timesteps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"grid{GRID_SIZE} (config 4: synthetic {GRID_SIZE}x{GRID_SIZE}-node lattice, 1046528 "
                                   f"directed links, 1 replica per GPU, philox draws, uniform turning fractions)",
                       "reference_sample": sample},
            "cpu_baseline": {"value": value, "unit": "link-timesteps/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "link-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------- batched env (config 5)
def bench_env45(torch, dist, world, rank, dev, replicas, steps, warmup):
    """BASELINE config 5: data/45_intersections, batched replicas with random gate actions in [0, 4] m,
    obs_mode option3, action_gap 1.  Returns (env-steps/s device-resident, env-steps/s end-to-end with the
    actions coming from pinned host memory and observations + rewards copied back every step)."""
    from pednstream_b200.rl import BatchedPedNetEnv
    env = BatchedPedNetEnv("45_intersections", replicas=replicas, obs_mode="option3", seed=1000 + rank,
                           replica_base=rank * replicas, device=dev)
    R, A = env.R, env.n_act
    gen = torch.Generator(device=dev)
    gen.manual_seed(rank)
    actions = torch.rand((warmup + 2 * steps, R, A), generator=gen, device=dev, dtype=torch.float32) * 4.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(warmup):
        env.step(actions[k])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        env.step(actions[warmup + k])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    host_actions = actions[warmup + steps:].cpu().pin_memory()
    host_obs = torch.zeros((R, env.n_obs), dtype=torch.float32).pin_memory()
    host_rew = torch.zeros((R,), dtype=torch.float32).pin_memory()
    dev_act = torch.zeros((R, A), dtype=torch.float32, device=dev)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for k in range(steps):
        dev_act.copy_(host_actions[k], non_blocking=True)
        obs, rew, done, _ = env.step(dev_act)
        host_obs.copy_(obs, non_blocking=True)
        host_rew.copy_(rew, non_blocking=False)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    env.engine.check_errors()
    # episode turnover: reset = state init + demand of the next episode drawn on the device (wall clock, it
    # has host work in it)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    env.reset()
    torch.cuda.synchronize()
    reset_ms = (time.perf_counter() - w0) * 1e3
    t = torch.tensor([ms, ms_e2e, reset_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, reset_ms = float(t[0]), float(t[1]), float(t[2])
    S_ep = env.simulation_steps
    episode_ms = S_ep * ms / steps + reset_ms
    links = env.engine.L
    return {"env_steps_per_s": world * R * steps / (ms * 1e-3),
            "env_steps_per_s_e2e": world * R * steps / (ms_e2e * 1e-3),
            "env_steps_per_s_with_resets": world * R * S_ep / (episode_ms * 1e-3),
            "reset_ms": reset_ms, "steps_per_episode": S_ep,
            "link_timesteps_per_s": world * R * links * steps / (ms * 1e-3),
            "replicas_per_gpu": R, "replicas_total": world * R, "env_steps": steps, "ms_per_env_step": ms / steps,
            "launches_per_env_step": env.launches_per_step(), "links": links,
            "h2d_bytes_per_step": R * A * 4, "d2h_bytes_per_step": R * (env.n_obs + 1) * 4,
            "workload": "config 5: data/45_intersections, obs option3, uniform random gate actions, philox draws, "
                        "per-replica demand drawn on the device at reset",
            "mean_reward_last_step": float(host_rew.mean())}


# --------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pednstream_b200 import _native
    from pednstream_b200.engine import Engine
    from pednstream_b200.grid import build_grid_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    K, W = args.steps, max(args.warmup, 3)
    size = args.grid
    Kp, Ke = min(K, 50), min(K, 100)                     # profiled / end-to-end region lengths
    S = W + K + Kp + Ke + 2                              # history rows: 80 B x links x (S+1) of HBM
    if 80.0 * (S + 1) * (2 * 2 * size * (size - 1)) > 150e9:
        raise SystemExit(f"--steps {K}: the {size}x{size} history would not fit in HBM; use <= 1500 steps")
    link_over = {"speed_noise_std": args.sigma} if args.sigma is not None else None
    from pednstream_b200.grid import default_origins
    plan, gate, tf, demand = build_grid_plan(size, S, demand_seed=rank, locality_order=True, link=link_over,
                                             origins=default_origins(size, args.origin_stride))
    L = plan["n_links"]
    eng = Engine(plan, replicas=1, rng="philox", seed=rank, device=dev)
    eng.initialise(gate, None, tf, demand, None)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the timed region: inputs resident in HBM, K steps in one native call ----
    eng.run(1, W)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run(W + 1, K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t_next = W + K + 1

    # ---- per-kernel durations (CUDA events on the launching stream, inside the library) --------
    k_ms, k_cnt = eng.run_profiled(t_next, Kp)
    t_next += Kp
    names = ("link_pair", "route_probs", "node_flows")
    per_kernel = {n: (k_ms[i] / k_cnt[i] if k_cnt[i] else None) for i, n in enumerate(names)}

    # ---- end to end through the public call, host buffers: every step copies its demand row from
    # pinned host memory, launches the step, and reads the network-wide pedestrian count back ----
    pinned = torch.zeros((eng.demand.shape[0], eng.demand.shape[1]), dtype=torch.float64).pin_memory()
    pinned[: demand.shape[0], : demand.shape[1]] = torch.from_numpy(np.ascontiguousarray(demand))
    from pednstream_b200 import _native as _nat
    result_host = torch.zeros((Ke, _nat.METRIC_ROW), dtype=torch.float64).pin_memory()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    eng.run_streamed(t_next, Ke, pinned, result_host)      # public API: per-step H2D demand row + D2H step metric
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    num_hist = None
    eng.check_errors()
    total_peds = float(eng.streamed_metric(result_host, Ke)[-1])

    t_ms = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t_ms[0]), float(t_ms[1])

    # free the lattice history before the batched environment allocates its own
    h2d_bytes = int(plan["n_demand_rows"] * 8)
    del num_hist, pinned
    eng = None
    torch.cuda.empty_cache()
    env_stats = None
    if not args.no_env:
        env_stats = bench_env45(torch, dist, world, rank, dev, args.env_replicas, min(K, 300), 10)

    if rank == 0:
        peak, peak_src = measured_peak()
        value = world * L * K / (ms * 1e-3)
        e2e = world * L * Ke / (ms_e2e * 1e-3)
        dom = max((n for n in names if per_kernel[n]), key=lambda n: per_kernel[n])
        step_gbs = B_ALG * L / (ms / K * 1e-3) / 1e9
        dom_gbs = B_ALG_PASS[dom] * L / (per_kernel[dom] * 1e-3) / 1e9
        cpu_v, cpu_dt, cpu_L = (None, None, None)
        if world == 1 and not args.no_cpu_baseline:
            cpu_steps = 150                   # ~13 s of single-core work (the brief asks for 10-30 s)
            cpu_v, cpu_dt, cpu_L = oracle_sample(cpu_steps, 2)
        line = {
            "metric": "link-timesteps/sec", "value": value, "unit": "link-timesteps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"grid{size} (config 4: synthetic {size}x{size}-node lattice, {L} directed links, "
                                   f"1 replica per GPU, philox draws, uniform turning fractions)",
                       "links": L, "nodes": plan["n_nodes"], "origins": plan["n_demand_rows"],
                       "l2_policy": "inputs larger than L2: one step touches >= 176 B x links = "
                                    f"{B_ALG * L / 1e6:.0f} MB of history",
                       "multi_gpu": "replicas only: one independent grid per rank, no data-path collective"},
            "e2e": {"value": e2e, "unit": "link-timesteps/s", "steps": Ke,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8 * _nat.METRIC_ROW,
                    "note": "Engine.run_streamed (C-ABI pns_step_streamed): every step copies its demand row from "
                            "pinned host memory, runs the step (whose link kernel also reduces the network-wide "
                            "pedestrian count to 64 partial sums) and copies those to pinned host memory; "
                            "stream-ordered, one host wait at the end"},
            "gpu_launches": int(2 * K + 1 + (K if plan["rt_grp_node"].size else 0)),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_link_lane" if dom == "link_pair" else "k_" + dom, "achieved": dom_gbs, "peak": peak,
                         "unit": "GB/s", "frac": dom_gbs / peak, "traffic": measured_traffic(dom),
                         "peak_source": peak_src,
                         "alg_bytes_per_link_step": B_ALG_PASS[dom],
                         "kernel_ms": {k: v for k, v in per_kernel.items() if v},
                         "step": {"achieved": step_gbs, "frac": step_gbs / peak, "alg_bytes_per_link_step": B_ALG}},
            "check": {"pedestrians_on_links_last_step": total_peds},
        }
        if env_stats is not None:
            line["batched_env"] = env_stats
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": "link-timesteps/s", "cores": 1, "kind": "port",
                                    "sample": f"Python oracle, {REF_SAMPLE_SIZE}x{REF_SAMPLE_SIZE} lattice ({cpu_L} links), "
                                              f"{cpu_steps} steps, {cpu_dt:.1f} s; host has {os.cpu_count()} cores, the "
                                              f"reference algorithm is single-threaded"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=GRID_SIZE)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sigma", type=float, default=None, help="override speed_noise_std of the lattice (experiments)")
    ap.add_argument("--origin-stride", type=int, default=64, help="every n-th boundary node is an origin")
    ap.add_argument("--no-env", action="store_true", help="skip the batched 45_intersections environment section")
    ap.add_argument("--env-replicas", type=int, default=1024, help="replicas per GPU of the batched environment")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
