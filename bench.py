#!/usr/bin/env python
"""Benchmark of the LTM timestep (BASELINE.json metric: link-timesteps/sec, fp64 state).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid SIZE]
                    [--no-env] [--no-cpu-baseline] [--no-variants] [--no-small]

Workload at N=1 (`config.workload`): BASELINE config 4, the synthetic 512x512-node lattice of
data/create_grid.py's rule (1 046 528 directed links, default_link of data/45_intersections, 35
origins = 4 corners + every 64th boundary node, gaussian-peak Poisson demand pre-drawn on the host,
uniform turning fractions), single replica, on-device Philox draws.  A "step" is one
`network_loading(t)` over the whole network.  With N>1 every rank runs an independent replica of the
grid (replicas only, no data-path collective; weak scaling).

Further blocks of the JSON line (rank 0):
  roofline.variants   the dominant kernel in the timed region, late in the run (t ~ 950) and on the dense-boundary
                      workload (every boundary node an origin)
  config4b_routed_lattice   the routed variant of config 4 (8 OD pairs, k_paths 3): sparse setup time, step time, roofline
  batched_env         BASELINE config 5 at its stated size: data/45_intersections, 8192 replicas in total
                      (8192/N per GPU), random gate actions, with its own roofline and end-to-end figures
  small_configs       BASELINE configs 1-3 (long_corridor, nine_intersections, melbourne@2000): GPU wall time per
                      step in numpy-compatible and philox draw modes
  optimal_node_model  assign_flows_type 'optimal' (one linear program per node, replica and step): the batched
                      environment at 1024 replicas with the device simplex, beside scipy.optimize.linprog (the
                      reference's solver) on the host for a sample of recorded programs
  cpu_baseline        the reference's own code (baseline/_ref, kind "reference"; the oracle port when the
                      reference is not installed) on one host core, bounded sample

`--impl reference` times the reference's own CPU implementation on all host cores: the config-4 sample (one 32x32
lattice per core -- the reference cannot build a 512x512 network from a dense adjacency matrix) as the line's
value, plus configs 1, 2, 3 (sample) and 5 (one episode per core) in `reference_configs`.
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
# split of B_ALG by kernel (DESIGN.md "kernels"; sums to 176): the fused link kernel carries the
# 96 B of lagged / previous-row reads and 48 B of link-state + sending/receiving writes, the node kernel
# the 32 B of flow / cumulative-count writes.  Intermediate re-reads between kernels are not algorithmic.
B_ALG_PASS = {"link_pair": 144.0, "route_probs": 0.0, "node_flows": 32.0}
GRID_SIZE = 512
REF_SAMPLE_SIZE = 32
ENV_REPLICAS_TOTAL = 8192          # BASELINE.json configs[4]
LATE_T, LATE_K = 950, 40           # late-time window of the lattice run
DENSE_WARM, DENSE_K = 600, 40      # dense-boundary variant: steps before / inside its window
ROUTED_WARM = 200                  # config 4b: steps before its timed window
ROUTED_ORIGIN_COLS = (112, 128, 144, 160)      # config 4b: origins on row 0 ...
ROUTED_DEST_ROW, ROUTED_DEST_COLS = 32, (120, 152)   # ... destinations on row 32: 8 OD pairs, 40-72 links apart


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed single-pass ncu capture of this build
    (measured offline -- never under this run), or None."""
    for name in ("r2_final_traffic.json", "r1_final_traffic.json"):
        try:
            rec = json.load(open(os.path.join(ROOT, "profiles", name)))
            return rec[{"link_pair": "k_link_lane", "node_flows": "k_node_flows"}[kernel]]["dram_bytes_per_launch"]
        except Exception:
            continue
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


# ================================================================================= reference arm
def _lattice_params(size, S):
    from pednstream_b200.grid import DEFAULT_LINK, default_origins
    origins = default_origins(size, stride=64)
    params = {"unit_time": 10, "simulation_steps": S, "default_link": dict(DEFAULT_LINK),
              "demand": {f"origin_{o}": {"peak_lambda": 50, "base_lambda": 30} for o in origins}}
    return origins, params


def reference_installed():
    from oracle import ref_harness as rh
    return rh.reference_available()


def lattice_sample(steps, warmup, kind, min_seconds=0.0):
    """`steps` steps (more, in chunks of `steps`, until `min_seconds` have passed) of a 32x32 lattice with the workload's link parameters, origin rule and demand pattern on the
    host: kind "reference" = the reference's own Network (baseline/_ref), kind "port" = the Python oracle.
    Per-link-step cost of this code is size-independent.  Returns
    (link-timesteps/s, seconds, links)."""
    import numpy as np
    from pednstream_b200.grid import grid_adjacency
    size = REF_SAMPLE_SIZE
    S = max(1000, warmup + steps + 1)
    origins, params = _lattice_params(size, S)
    np.random.seed(0)
    if kind == "reference":
        from oracle import ref_harness as rh
        Network, _ = rh.import_reference()
        net = Network(grid_adjacency(size), params, origin_nodes=list(origins), verbose=False)
        step = net.network_loading
    else:
        from oracle.ltm_oracle import LtmOracle
        from pednstream_b200 import Network
        net = Network(grid_adjacency(size), params, origin_nodes=list(origins), verbose=False)
        step = LtmOracle(net).network_loading
    for t in range(1, warmup + 1):
        step(t)
    t0 = time.perf_counter()
    done, t = 0, warmup + 1
    while True:
        for _ in range(steps):
            step(t)
            t += 1
        done += steps
        dt = time.perf_counter() - t0
        if dt >= min_seconds or t + steps > S:
            break
    L = len(net.links)
    return L * done / dt, dt, L, done


def _lattice_worker(job):
    steps, warmup, kind = job
    value, dt, L, done = lattice_sample(steps, warmup, kind)
    return L * done, dt, L


def reference_dataset_run(job):
    """One reference run of a shipped scenario: (links * steps, seconds, links, steps).  `env`: through the
    reference's own PedNetParallelEnv with random gate actions (config 5), else network_loading directly."""
    import numpy as np
    from oracle import ref_harness as rh
    name, steps, steps_override, env = job
    np.random.seed(0)
    if env:
        e = rh.make_reference_env(name, obs_mode="option3", seed=0)
        e.reset()
        L = len(e.network.links)
        t0 = time.perf_counter()
        for _ in range(steps):
            e.step({a: e.action_space(a).sample() for a in e.possible_agents})
        dt = time.perf_counter() - t0
    else:
        net, _ = rh.create_network(name, steps_override=steps_override, verbose=False)
        L = len(net.links)
        t0 = time.perf_counter()
        for t in range(1, steps + 1):
            net.network_loading(t)
        dt = time.perf_counter() - t0
    return L * steps, dt, L, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    kind = "reference" if reference_installed() else "port"
    steps = max(5, min(args.steps, 40))
    warmup = min(args.warmup, 3)
    cores = os.cpu_count() or 1
    ref_cfg = {}
    with mp.get_context("spawn").Pool(cores) as pool:
        # config 4 sample: the reference algorithm is single-threaded; its only parallelism is independent
        # replicas (the reference's RLlib workers), so the best case on the host is one lattice per core
        out = pool.map(_lattice_worker, [(steps, warmup, kind)] * cores)
        work, slowest, L = sum(o[0] for o in out), max(o[1] for o in out), out[0][2]
        if kind == "reference" and not args.no_small:
            jobs = {"config1_long_corridor": ("long_corridor", 500, None, False),
                    "config2_nine_intersections": ("nine_intersections", 499, None, False),
                    "config3_melbourne_2000_sample": ("melbourne", 150, 2000, False)}
            res = pool.map(reference_dataset_run, list(jobs.values()))
            for key, (w, dt, l, n) in zip(jobs, res):
                ref_cfg[key] = {"links": l, "steps": n, "seconds": dt, "ms_per_step": 1e3 * dt / n,
                                "link_timesteps_per_s": w / dt, "cores": 1}
            # config 5: one full episode of the reference's own env per core
            ep = pool.map(reference_dataset_run, [("45_intersections", 700, None, True)] * cores)
            dt5 = max(o[1] for o in ep)
            ref_cfg["config5_45_intersections_env"] = {
                "links": ep[0][2], "steps": 700, "episodes": cores, "cores": cores, "seconds": dt5,
                "env_steps_per_s": cores * 700 / dt5, "link_timesteps_per_s": sum(o[0] for o in ep) / dt5,
                "note": "reference rl/pz_pednet_env.PedNetParallelEnv, obs option3, action_space.sample() every step; "
                        "one episode per host core"}
    value = work / slowest
    src = ("the reference's own Network (baseline/_ref, unmodified)" if kind == "reference"
           else "the Python oracle port (baseline/_ref not installed)")
    sample = (f"{cores} independent {REF_SAMPLE_SIZE}x{REF_SAMPLE_SIZE} lattices ({L} links each), {steps} steps, same "
              f"link parameters / origin rule / demand pattern as the workload; one Python process per host core; {src}")
    line = {"impl": "reference", "metric": "link-timesteps/sec", "value": value, "unit": "link-timesteps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * slowest / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"grid{GRID_SIZE} (config 4: synthetic {GRID_SIZE}x{GRID_SIZE}-node lattice, 1046528 "
                                   f"directed links, 1 replica per GPU, philox draws, uniform turning fractions)",
                       "reference_sample": sample},
            "cpu_baseline": {"value": value, "unit": "link-timesteps/s", "cores": cores, "kind": kind,
                             "sample": sample},
            "e2e": {"value": value, "unit": "link-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if ref_cfg:
        line["reference_configs"] = ref_cfg
    print(json.dumps(line), flush=True)


# ================================================================================= batched env (config 5)
def bench_env45(torch, dist, world, rank, dev, replicas, steps, warmup, peak, congested=True):
    """BASELINE config 5: data/45_intersections, batched replicas with random gate actions in [0, 4] m,
    obs_mode option3, action_gap 1.  Device-resident and end-to-end (actions from pinned host memory, observations
    and rewards back to pinned host memory every step, `BatchedPedNetEnv.rollout_host`) env-steps/s at the start of
    an episode, and device-resident again in the congested state around step 300."""
    from pednstream_b200.rl import BatchedPedNetEnv
    env = BatchedPedNetEnv("45_intersections", replicas=replicas, obs_mode="option3", seed=1000,
                           replica_base=rank * replicas, device=dev)
    R, A = env.R, env.n_act
    links = env.engine.L
    gen = torch.Generator(device=dev)
    gen.manual_seed(rank)
    pool = torch.rand((16, R, A), generator=gen, device=dev, dtype=torch.float32) * 4.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    roll_obs = torch.empty((steps, R, env.n_obs), dtype=torch.float32, device=dev)
    roll_rew = torch.empty((steps, R), dtype=torch.float32, device=dev)
    roll_act = pool.repeat((steps + 15) // 16, 1, 1)[:steps].contiguous()

    def timed(n):
        """n env steps with device-resident random actions in one native call (BatchedPedNetEnv.rollout)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        env.rollout(roll_act[:n], roll_obs[:n], roll_rew[:n])
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    for k in range(warmup):
        env.step(pool[k & 15])
    ms = timed(steps)
    # end to end: every step its own H2D (actions) and D2H (observations + rewards), pipelined on two copy streams
    e2e_steps = max(steps, 100)
    host_actions = (torch.rand((e2e_steps, R, A), dtype=torch.float32) * 4.0).pin_memory()
    host_obs = torch.zeros((e2e_steps, R, env.n_obs), dtype=torch.float32).pin_memory()
    host_rew = torch.zeros((e2e_steps, R), dtype=torch.float32).pin_memory()
    env.rollout_host(host_actions[:2], host_obs[:2], host_rew[:2])          # creates the streams / buffers
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    env.rollout_host(host_actions, host_obs, host_rew)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    # congested state (queues at the gater and the origins; blockers draws of jammed links)
    ms_late, late_from = None, None
    if congested and env.sim_step + steps < 300:
        while env.sim_step < 300:
            env.step(pool[env.sim_step & 15])
        late_from = env.sim_step
        ms_late = timed(steps)
    env.engine.check_errors()
    # cross-GPU gather of the per-replica episode statistic (NCCL all-gather over NVLink), timed on the device
    gather_us = None
    if world > 1:
        from pednstream_b200 import parallel
        parallel.gather_replica_values(env.cumulative_reward)               # warm-up (communicator set-up)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        for _ in range(10):
            allr = parallel.gather_replica_values(env.cumulative_reward)
        g1.record()
        barrier()
        gather_us = 1e3 * g0.elapsed_time(g1) / 10
        assert allr.shape[0] == world * R
    # episode turnover: reset = state init + demand of the next episode drawn on the device (wall clock, it
    # has host work in it)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    env.reset()
    torch.cuda.synchronize()
    reset_ms = (time.perf_counter() - w0) * 1e3
    # domain randomisation drawn on the device: episode turnover with a new scenario per replica
    rand_reset_ms = None
    S_ep, launches, n_obs = env.simulation_steps, env.launches_per_step(), env.n_obs
    if world == 1:
        del env
        torch.cuda.empty_cache()
        env = BatchedPedNetEnv("45_intersections", replicas=replicas, obs_mode="option3", seed=1000,
                               replica_base=rank * replicas, device=dev, randomize="device")
        for k in range(3):
            env.step(pool[k])
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        env.reset()
        torch.cuda.synchronize()
        rand_reset_ms = (time.perf_counter() - w0) * 1e3
        env.engine.check_errors()
        # scenario groups: per-group origin / destination nodes (generate_random_od_nodes), per-replica everything else
        del env
        torch.cuda.empty_cache()
        env = None
        try:        # an extra: its failure must not take the line's main figures with it
            from pednstream_b200.rl import GroupedPedNetEnv
            n_groups = 8
            genv = GroupedPedNetEnv("45_intersections", replicas=replicas, groups=n_groups, obs_mode="option3", seed=1000,
                                    device=dev, randomize="device")
            for k in range(warmup):
                genv.step(pool[k % len(pool)])
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for k in range(steps):
                genv.step(pool[k % len(pool)])
            g1.record()
            torch.cuda.synchronize()
            genv.check_errors()
            grouped = {"groups": n_groups, "replicas": replicas, "ms_per_env_step": g0.elapsed_time(g1) / steps,
                       "env_steps_per_s": replicas / (g0.elapsed_time(g1) / steps * 1e-3),
                       "distinct_od_node_sets": len({str(od) for od in genv.od_nodes}),
                       "note": "GroupedPedNetEnv: every group perturbs the origin / destination nodes with its own seed "
                               "(reference generate_random_od_nodes) and runs its own plan on its own stream; scenarios "
                               "inside a group are per replica, drawn on the device"}
        except Exception as exc:                     # noqa: BLE001
            grouped = {"error": repr(exc)[:300]}
    vals = [ms, ms_e2e, reset_ms, ms_late if ms_late is not None else 0.0, gather_us or 0.0]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, reset_ms, ms_late_max, gather_us_max = [float(v) for v in t]
    episode_ms = S_ep * ms / steps + reset_ms

    def roof(ms_window):
        gbs = B_ALG * world * R * links * steps / (ms_window * 1e-3) / 1e9
        return {"achieved": gbs, "frac": gbs / (peak * world), "link_timesteps_per_s": world * R * links * steps / (ms_window * 1e-3),
                "us_per_env_step": 1e3 * ms_window / steps}

    out = {"env_steps_per_s": world * R * steps / (ms * 1e-3),
           "env_steps_per_s_e2e": world * R * e2e_steps / (ms_e2e * 1e-3), "e2e_env_steps": e2e_steps,
           "env_steps_per_s_with_resets": world * R * S_ep / (episode_ms * 1e-3),
           "reset_ms": reset_ms, "steps_per_episode": S_ep,
           "link_timesteps_per_s": world * R * links * steps / (ms * 1e-3),
           "replicas_per_gpu": R, "replicas_total": world * R, "scaling": "strong (8192 replicas in total)",
           "env_steps": steps, "ms_per_env_step": ms / steps,
           "launches_per_env_step": launches, "links": links,
           "h2d_bytes_per_step": R * A * 4, "d2h_bytes_per_step": R * (n_obs + 1) * 4,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak * world, "alg_bytes_per_link_step": B_ALG,
                        "note": "whole env step (link flows + route choice + node model + link update with actions, "
                                "observations and reward riding in the link kernels): 176 B x links x replicas / step time",
                        "episode_start": roof(ms)},
           "workload": "config 5: data/45_intersections, obs option3, uniform random gate actions, philox draws, "
                       "per-replica demand drawn on the device at reset",
           "e2e_note": "BatchedPedNetEnv.rollout_host: per step H2D of the actions and D2H of observations + rewards "
                       "(pinned host memory), double-buffered on two copy streams",
           "mean_reward_last_step": float(host_rew[-1].mean())}
    if ms_late is not None:
        out["roofline"]["congested_t%d" % late_from] = roof(ms_late_max)
    if rand_reset_ms is not None:
        out["randomized_reset_ms"] = rand_reset_ms
        out["randomized_reset_note"] = ("BatchedPedNetEnv(randomize='device'): per-replica link bottlenecks, OD weights and "
                                        "demand patterns drawn by one kernel launch (pns_env_randomize), then state init and "
                                        "demand draw; wall clock")
    if rand_reset_ms is not None:
        out["od_node_groups"] = grouped
    if world > 1:
        out["reward_gather_us"] = gather_us_max
        out["reward_gather_note"] = "all-gather of one float32 per replica over NCCL (per call, device-timed)"
    del env
    return out


# ================================================================================= 'optimal' node model
def bench_optimal(torch, dev, steps=20):
    """45_intersections x 1024 replicas with assign_flows_type 'optimal' (reference node.py:249-271): step time with
    the warp-level simplex (k_node_lp), and the reference's solver call on the host for recorded programs."""
    import numpy as np
    from pednstream_b200.rl import BatchedPedNetEnv
    R = 1024
    out = {}
    for model in ("classic", "optimal"):
        env = BatchedPedNetEnv("45_intersections", replicas=R, obs_mode="option3", seed=1000, device=dev,
                               params={"assign_flows_type": model})
        acts = torch.rand((60 + steps, R, env.n_act), device=dev, dtype=torch.float32) * 4.0
        env.rollout(acts[:60])                                  # into the loaded state: most nodes carry flow
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.rollout(acts[60:])
        e1.record()
        torch.cuda.synchronize()
        env.engine.check_errors()
        out[model] = {"us_per_env_step": 1e3 * e0.elapsed_time(e1) / steps, "lp_nodes": int(env.engine.net.n_lp_nodes)}
        del env
        torch.cuda.empty_cache()
    n_prog = out["optimal"]["lp_nodes"] * R
    extra_us = out["optimal"]["us_per_env_step"] - out["classic"]["us_per_env_step"]
    out["programs_per_step"] = n_prog
    out["ns_per_program"] = 1e3 * extra_us / n_prog
    try:        # the reference's solver on the host (test infrastructure; the one place bench.py may execute oracle/)
        from oracle.ltm_oracle import scipy_lp
        gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "lp_programs.npz"))
        n, t0 = 0, time.perf_counter()
        for m in (3, 4, 5):
            for k in range(40):
                scipy_lp(m, gold[f"s_{m}"][k], gold[f"r_{m}"][k], gold[f"phi_{m}"][k])
                n += 1
        out["host_linprog_us_per_program"] = 1e6 * (time.perf_counter() - t0) / n
        out["host_note"] = ("scipy.optimize.linprog as the reference calls it (node.py:262) on 120 programs recorded "
                            "from the reference's runs, one host core")
    except Exception as exc:                    # noqa: BLE001
        out["host_note"] = f"host solver not timed: {exc!r}"[:200]
    return out


# ================================================================================= small configs (1-3)
def bench_small_configs(torch, dev):
    """BASELINE configs 1-3 on the GPU through the facade: wall time per network_loading(t) in the numpy-compatible
    draw mode (bit-identical to the reference; one host round trip per step) and in philox mode (no host in the
    loop; one native call for the whole run)."""
    import numpy as np
    from pednstream_b200 import NetworkEnvGenerator, _native
    out = {}
    for key, name, steps, override in (("config1_long_corridor", "long_corridor", 500, None),
                                       ("config2_nine_intersections", "nine_intersections", 499, None),
                                       ("config3_melbourne_2000", "melbourne", 1999, 2000)):
        rec = {}
        for mode in ("numpy", "philox"):
            np.random.seed(0)
            g = NetworkEnvGenerator()
            if override:
                g.network_data = g.load_network_data(name)
                g.config["params"]["simulation_steps"] = override
            net = g.create_network(name, verbose=False, rng=mode, seed=0, device=dev)
            eng = net.engine
            net.network_loading(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if mode == "numpy":
                for t in range(2, steps + 1):
                    net.network_loading(t)
            else:
                eng._push_all_demand(net)
                eng.run(2, steps - 1, _native.RNG_PHILOX)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            eng.check_errors()
            L = len(net.links)
            rec[mode] = {"ms_per_step": 1e3 * dt / (steps - 1), "link_timesteps_per_s": L * (steps - 1) / dt}
            rec["links"], rec["steps"] = L, steps
        out[key] = rec
    return out


# ================================================================================= our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pednstream_b200 import _native
    from pednstream_b200.engine import Engine
    from pednstream_b200.grid import build_grid_plan, default_origins

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = measured_peak()

    K, W = args.steps, max(args.warmup, 3)
    size = args.grid
    Kp, Ke = min(K, 50), 100                             # profiled / end-to-end region lengths (the end-to-end
                                                         # region is 100 steps whatever K is: at 20 steps host jitter
                                                         # moves it by +-15 %)
    want_late = not args.no_variants and W + K + Kp + Ke + 8 < LATE_T
    S = max(W + K + Kp + Ke + 10, LATE_T + LATE_K + 2 if want_late else 0)   # history rows: 80 B x links x (S+1)
    n_links_est = 2 * 2 * size * (size - 1)
    if 80.0 * (S + 1) * n_links_est > 150e9:
        if want_late:
            want_late = False
            S = W + K + Kp + Ke + 10
        if 80.0 * (S + 1) * n_links_est > 150e9:
            raise SystemExit(f"--steps {K}: the {size}x{size} history would not fit in HBM; use <= 1500 steps")
    link_over = {"speed_noise_std": args.sigma} if args.sigma is not None else None
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

    names = ("link_pair", "route_probs", "node_flows")

    def profiled(engine, t0, n):
        k_ms, k_cnt = engine.run_profiled(t0, n)
        return {nm: (k_ms[i] / k_cnt[i] if k_cnt[i] else None) for i, nm in enumerate(names)}

    def chained(engine, t0, n):
        """ms per step of n steps in one native call (the programmatic-launch chain, as in the timed region)."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        engine.run(t0, n)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    def roof_of(per_kernel, links, chained_ms=None):
        dom = max((n for n in names if per_kernel[n]), key=lambda n: per_kernel[n])
        gbs = B_ALG_PASS[dom] * links / (per_kernel[dom] * 1e-3) / 1e9
        step_ms = sum(v for v in per_kernel.values() if v)
        out = {"kernel": "k_link_lane" if dom == "link_pair" else "k_" + dom, "achieved": gbs, "frac": gbs / peak,
               "kernel_ms": {k: v for k, v in per_kernel.items() if v},
               "step_serialised": {"ms": step_ms, "frac": B_ALG * links / (step_ms * 1e-3) / 1e9 / peak}}
        if chained_ms is not None:
            out["step"] = {"ms": chained_ms, "frac": B_ALG * links / (chained_ms * 1e-3) / 1e9 / peak}
        return dom, out

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
    per_kernel = profiled(eng, t_next, Kp)
    t_next += Kp

    # stores the exact shortcuts skipped in the last step (idle nodes leave the zero inflow / outflow rows alone):
    # what the step really moved, next to the 176 algorithmic bytes it is credited with
    stride = int(plan["nd_stride"])
    active = (eng.nm_s.view(-1, stride) != 0).any(dim=1)
    slot_node = torch.from_numpy(np.ascontiguousarray(plan["lk_slots"]) // stride).to(dev)
    skipped = int((~active[slot_node[:, 0]]).sum()) + int((~active[slot_node[:, 1]]).sum())
    moved_bytes = B_ALG - 8.0 * skipped / L

    # ---- end to end through the public call, host buffers: every step copies its demand row from
    # pinned host memory, launches the step, and reads the network-wide pedestrian count back ----
    pinned = torch.zeros((eng.demand.shape[0], eng.demand.shape[1]), dtype=torch.float64).pin_memory()
    pinned[: demand.shape[0], : demand.shape[1]] = torch.from_numpy(np.ascontiguousarray(demand))
    from pednstream_b200 import _native as _nat
    result_host = torch.zeros((Ke + 4, _nat.METRIC_ROW), dtype=torch.float64).pin_memory()
    eng.run_streamed(t_next, 4, pinned, result_host)      # untimed: creates the copy stream and its events
    t_next += 4
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    eng.run_streamed(t_next, Ke, pinned, result_host)      # public API: per-step H2D demand row + D2H step metric
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    t_next += Ke
    eng.check_errors()
    total_peds = float(eng.streamed_metric(result_host, Ke)[-1])

    # ---- roofline variants: late in the run, and the dense-boundary workload ----------------------------------
    variants = {}
    dom, rf = roof_of(per_kernel, L, ms / K)
    rf["window"] = f"t = {W + 1} .. {W + K} (chained step = the timed region), {W + K + 1} .. {W + K + Kp} (per kernel)"
    variants["timed_region"] = rf
    if want_late and rank == 0:
        eng.run(t_next, LATE_T - LATE_K - t_next)
        ms_late_step = chained(eng, LATE_T - LATE_K, LATE_K)
        _, rl = roof_of(profiled(eng, LATE_T, LATE_K), L, ms_late_step)
        rl["window"] = f"t = {LATE_T - LATE_K} .. {LATE_T - 1} (chained step), {LATE_T} .. {LATE_T + LATE_K - 1} (per kernel)"
        rl["pedestrians_on_links"] = float(eng.history("num_pedestrians")[LATE_T + LATE_K - 1].sum())
        variants["late"] = rl

    t_ms = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t_ms[0]), float(t_ms[1])

    # free the lattice history before the next allocation
    h2d_bytes = int(plan["n_demand_rows"] * 8)
    n_nodes, n_origins = plan["n_nodes"], plan["n_demand_rows"]
    routed = bool(plan["rt_grp_node"].size)
    del pinned, active, slot_node
    eng = None
    torch.cuda.empty_cache()

    if not args.no_variants and rank == 0 and world == 1:
        Sd = DENSE_WARM + 2 * DENSE_K + 4
        if 80.0 * (Sd + 1) * n_links_est < 150e9:
            pl2, g2, tf2, d2 = build_grid_plan(size, Sd, demand_seed=0, locality_order=True, link=link_over,
                                               origins=default_origins(size, 1))
            e2_ = Engine(pl2, replicas=1, rng="philox", seed=0, device=dev)
            e2_.initialise(g2, None, tf2, d2, None)
            e2_.run(1, DENSE_WARM)
            ms_dense_step = chained(e2_, DENSE_WARM + 1, DENSE_K)
            _, rd = roof_of(profiled(e2_, DENSE_WARM + DENSE_K + 1, DENSE_K), pl2["n_links"], ms_dense_step)
            rd["window"] = (f"every boundary node an origin ({pl2['n_demand_rows']} origins), t = {DENSE_WARM + 1} .. "
                            f"{DENSE_WARM + DENSE_K} (chained step), then {DENSE_K} steps per kernel")
            rd["pedestrians_on_links"] = float(e2_.history("num_pedestrians")[DENSE_WARM + 2 * DENSE_K].sum())
            variants["dense_boundary"] = rd
            e2_ = None
            del pl2, g2, tf2, d2
            torch.cuda.empty_cache()

    # ---- config 4b: the routed lattice (8 OD pairs, k_paths = 3): route-choice kernel + routed node kernel ----
    cfg4b = None
    if not args.no_variants and rank == 0 and world == 1:
        from pednstream_b200.grid import build_routed_grid_plan
        W4, K4 = ROUTED_WARM, max(20, min(K, 100))
        S4 = W4 + 2 * K4 + 8
        if 80.0 * (S4 + 1) * n_links_est < 150e9:
            w0 = time.perf_counter()
            orig4 = [c for c in ROUTED_ORIGIN_COLS if c < size]
            dest4 = [ROUTED_DEST_ROW * size + c for c in ROUTED_DEST_COLS if c < size and ROUTED_DEST_ROW < size]
            pl4, g4, tf4, d4, od4 = build_routed_grid_plan(size, S4, orig4, dest4, demand_seed=0, locality_order=True,
                                                          link=link_over)
            setup_s = time.perf_counter() - w0
            e4 = Engine(pl4, replicas=1, rng="philox", seed=0, device=dev)
            e4.initialise(g4, None, tf4, d4, od4)
            e4.run(1, W4)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            e4.run(W4 + 1, K4)
            a1.record()
            torch.cuda.synchronize()
            ms4 = a0.elapsed_time(a1)
            _, r4 = roof_of(profiled(e4, W4 + K4 + 1, K4), pl4["n_links"])
            e4.check_errors()
            gbs4 = B_ALG * pl4["n_links"] / (ms4 / K4 * 1e-3) / 1e9
            cfg4b = {"workload": f"config 4b: {size}x{size} routed lattice, {len(pl4['od_keys'])} OD pairs "
                                 f"(origins row 0, columns {list(ROUTED_ORIGIN_COLS)}; destinations row {ROUTED_DEST_ROW}, "
                                 f"columns {list(ROUTED_DEST_COLS)}), k_paths 3, {len(pl4['rt_routed_nodes'])} routed nodes, "
                                 f"{len(pl4['rt_opt_link'])} route options; steps {W4 + 1}..{W4 + K4}",
                     "value": pl4["n_links"] * K4 / (ms4 * 1e-3), "unit": "link-timesteps/s", "ms_per_step": ms4 / K4,
                     "steps": K4, "setup_seconds": setup_s,
                     "setup_note": "sparse constructor: lattice plan + networkx k-shortest simple paths and turn structures "
                                   "(grid.build_routed_grid_plan), no dense adjacency matrix",
                     "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "step": {"achieved": gbs4, "frac": gbs4 / peak},
                                  **r4},
                     "arrivals_at_destinations": float(e4.history("cumulative_outflow")[W4 + K4].sum()),
                     "pedestrians_on_links": float(e4.history("num_pedestrians")[W4 + K4].sum())}
            e4 = None
            del pl4, g4, tf4, d4, od4
            torch.cuda.empty_cache()

    env_stats = None
    if not args.no_env:
        per_gpu = max(1, args.env_replicas // world)
        env_stats = bench_env45(torch, dist, world, rank, dev, per_gpu, max(10, min(K, 200)), 10, peak)
        torch.cuda.empty_cache()

    small = None
    if rank == 0 and world == 1 and not args.no_small:
        try:                                        # extras: a failure must not cost the line its main figures
            small = bench_small_configs(torch, dev)
        except Exception as exc:                    # noqa: BLE001
            small = {"error": repr(exc)[:300]}
    optimal = None
    if rank == 0 and world == 1 and not args.no_small:
        try:
            optimal = bench_optimal(torch, dev)
        except Exception as exc:                    # noqa: BLE001
            optimal = {"error": repr(exc)[:300]}

    if rank == 0:
        value = world * L * K / (ms * 1e-3)
        e2e = world * L * Ke / (ms_e2e * 1e-3)
        step_gbs = B_ALG * L / (ms / K * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            kind = "reference" if reference_installed() else "port"
            # chunks of 20 steps until ~12 s of single-core work have passed (the brief asks for 10-30 s)
            cpu_v, cpu_dt, cpu_L, cpu_steps = lattice_sample(20, 2, kind, min_seconds=12.0)
            cpu = {"value": cpu_v, "unit": "link-timesteps/s", "cores": 1, "kind": kind,
                   "sample": f"{'the reference (baseline/_ref, unmodified)' if kind == 'reference' else 'Python oracle port'}, "
                             f"{REF_SAMPLE_SIZE}x{REF_SAMPLE_SIZE} lattice ({cpu_L} links), {cpu_steps} steps, {cpu_dt:.1f} s; "
                             f"host has {os.cpu_count()} cores, the reference algorithm is single-threaded"}
        line = {
            "metric": "link-timesteps/sec", "value": value, "unit": "link-timesteps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"grid{size} (config 4: synthetic {size}x{size}-node lattice, {L} directed links, "
                                   f"1 replica per GPU, philox draws, uniform turning fractions)",
                       "links": L, "nodes": n_nodes, "origins": n_origins,
                       "l2_policy": "inputs larger than L2: one step touches >= 176 B x links = "
                                    f"{B_ALG * L / 1e6:.0f} MB of history",
                       "multi_gpu": "replicas only: one independent grid per rank, no data-path collective"},
            "e2e": {"value": e2e, "unit": "link-timesteps/s", "steps": Ke,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8 * _nat.METRIC_ROW,
                    "note": "Engine.run_streamed (C-ABI pns_step_streamed): every step copies its demand row from "
                            "pinned host memory, runs the step (whose link kernel also reduces the network-wide "
                            "pedestrian count to 64 partial sums) and copies those to pinned host memory; "
                            "stream-ordered, one host wait at the end"},
            "gpu_launches": int(2 * K + 1 + (K if routed else 0)),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": variants["timed_region"]["kernel"],
                         "achieved": variants["timed_region"]["achieved"], "peak": peak,
                         "unit": "GB/s", "frac": variants["timed_region"]["frac"], "traffic": measured_traffic(dom),
                         "peak_source": peak_src,
                         "alg_bytes_per_link_step": B_ALG_PASS[dom],
                         "kernel_ms": variants["timed_region"]["kernel_ms"],
                         "step": {"achieved": step_gbs, "frac": step_gbs / peak, "alg_bytes_per_link_step": B_ALG,
                                  "moved_bytes_per_link_step": moved_bytes,
                                  "moved_frac": step_gbs / peak * moved_bytes / B_ALG,
                                  "note": "moved = 176 B minus the inflow/outflow stores that idle nodes skip (rows are "
                                          "zero by contract), counted in the last profiled step"},
                         "variants": variants},
            "check": {"pedestrians_on_links_last_step": total_peds},
        }
        if cfg4b is not None:
            line["config4b_routed_lattice"] = cfg4b
        if env_stats is not None:
            line["batched_env"] = env_stats
        if small is not None:
            line["small_configs"] = small
        if optimal is not None:
            line["optimal_node_model"] = optimal
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=GRID_SIZE)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sigma", type=float, default=None, help="override speed_noise_std of the lattice (experiments)")
    ap.add_argument("--origin-stride", type=int, default=64, help="every n-th boundary node is an origin")
    ap.add_argument("--no-env", action="store_true", help="skip the batched 45_intersections environment section")
    ap.add_argument("--no-variants", action="store_true", help="skip the late-time / dense-boundary roofline variants")
    ap.add_argument("--no-small", action="store_true", help="skip BASELINE configs 1-3")
    ap.add_argument("--env-replicas", type=int, default=ENV_REPLICAS_TOTAL,
                    help="replicas of the batched environment in total (divided over the GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
