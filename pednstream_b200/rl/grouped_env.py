"""Scenario groups: per-group origin / destination nodes on top of the per-replica scenarios.

The reference's `NetworkEnvGenerator.randomize_network(seed)` (src/utils/env_loader.py:160-181) perturbs, besides
link bottlenecks, OD weights and demand patterns, the *set of origin and destination nodes*
(`generate_random_od_nodes`, :261-361).  That changes the topology the step runs on -- which nodes carry virtual
O/D links, their slot counts, the k-shortest paths and every route structure -- so replicas with different OD nodes
cannot share one plan.  `GroupedPedNetEnv` therefore splits the replicas into G groups; group g perturbs the OD
nodes with its own seed and compiles its own plan, and inside a group every replica still draws its own
bottlenecks, OD weights and demand (`BatchedPedNetEnv(randomize=...)`).  The controllers (agents) are the
scenario's, never O/D nodes, so actions and observations have the same layout in every group and the groups'
tensors are slices of one [R, ...] tensor.  Each group steps on its own CUDA stream.

The first replica of every group is, in episode 0 with randomize="host", exactly the network
`randomize_network(dataset, seed=scenario_seed)` builds -- OD nodes included (tests/test_env.py).
"""
from __future__ import annotations

import ctypes as C

import torch

from .batched_env import BatchedPedNetEnv


class GroupedPedNetEnv:
    def __init__(self, dataset: str, replicas: int, groups: int, obs_mode: str = "option3",
                 normalize_obs: bool = False, seed: int = 0, replica_base: int = 0, device=None, data_dir="data",
                 randomize=True, perturb_first_group: bool = True, action_gap: int = 1, _lib=None,
                 _emulation: bool = False):
        if not 1 <= groups <= replicas:
            raise ValueError("1 <= groups <= replicas")
        self.R, self.G, self.action_gap = int(replicas), int(groups), int(action_gap)
        sizes = [self.R // self.G + (1 if g < self.R % self.G else 0) for g in range(self.G)]
        self.envs, self.slices = [], []
        offset = 0
        for g, n in enumerate(sizes):
            base = int(replica_base) + offset
            # = BatchedPedNetEnv.scenario_seed(0, 0) of the group: its first replica is randomize_network(od_seed)
            od_seed = (int(seed) + 7919 * base + 1) % (2 ** 32)
            env = BatchedPedNetEnv(dataset, n, obs_mode=obs_mode, normalize_obs=normalize_obs, seed=seed,
                                   replica_base=base, device=device, data_dir=data_dir, randomize=randomize,
                                   od_nodes_seed=od_seed if (g > 0 or perturb_first_group) else None,
                                   action_gap=action_gap,
                                   _lib=_lib, _emulation=_emulation)
            self.envs.append(env)
            self.slices.append(slice(offset, offset + n))
            offset += n
        first = self.envs[0]
        for env in self.envs[1:]:
            if (env.n_act, env.n_obs, env.possible_agents) != (first.n_act, first.n_obs, first.possible_agents):
                raise RuntimeError("scenario groups disagree on the agents' action / observation layout")
        self.n_act, self.n_obs = first.n_act, first.n_obs
        self.possible_agents = first.possible_agents
        self.action_slices, self.obs_slices = first.action_slices, first.obs_slices
        self.simulation_steps = first.simulation_steps
        self.device = first.device
        self.emulation = first.engine.emulation
        self.streams = None if self.emulation else [torch.cuda.Stream(device=self.device) for _ in self.envs]
        if self.streams is not None:          # re-recorded every step (an earlier wait keeps the record it saw)
            self._ready = torch.cuda.Event()
            self._done = [torch.cuda.Event() for _ in self.envs]
        self.obs = torch.zeros((self.R, self.n_obs), dtype=torch.float32, device=self.device)
        self.reward = torch.zeros((self.R,), dtype=torch.float32, device=self.device)
        self._obs_views = {sl.start: self.obs[sl] for sl in self.slices}
        self._reward_views = {sl.start: self.reward[sl] for sl in self.slices}
        self._gather_obs()

    @property
    def od_nodes(self):
        """Per group: {'origin_nodes': [...], 'destination_nodes': [...]} (None = the scenario's own)."""
        return [env.od_nodes for env in self.envs]

    @property
    def sim_step(self):
        return self.envs[0].sim_step

    def _gather_obs(self):
        for env, sl in zip(self.envs, self.slices):
            self.obs[sl] = env.obs

    def _fork(self):
        if self.streams is None:
            return [None] * self.G
        self._ready.record(torch.cuda.current_stream(self.device))
        for st in self.streams:
            st.wait_event(self._ready)
        return self.streams

    def _join(self):
        if self.streams is None:
            return
        cur = torch.cuda.current_stream(self.device)
        for st, done in zip(self.streams, self._done):
            done.record(st)
            cur.wait_event(done)

    def reset(self):
        for env in self.envs:
            env.reset()
        self._gather_obs()
        return self.obs

    def step(self, actions: torch.Tensor):
        """actions: float32 [R, n_act] on the device -> (obs [R, n_obs], reward [R], done, info).
        One native call per group (pns_env_step on the group's stream), without the per-environment Python layers:
        with eight groups those cost more than the kernels."""
        if tuple(actions.shape) != (self.R, self.n_act) or actions.dtype != torch.float32:
            raise ValueError(f"actions must be float32 [{self.R}, {self.n_act}]")
        if not actions.is_contiguous():
            actions = actions.contiguous()
        t = self.envs[0].sim_step
        if t > self.simulation_steps:
            raise RuntimeError("episode finished: call reset()")
        if self.action_gap > 1:                 # several simulation steps per decision: the groups' own step()
            done = False
            with self.envs[0].engine._guard():
                streams = self._fork()
                for env, sl, st in zip(self.envs, self.slices, streams):
                    if st is None:
                        _, _, done, _ = env.step(actions[sl], self._obs_views[sl.start], self._reward_views[sl.start])
                    else:
                        with torch.cuda.stream(st):
                            _, _, done, _ = env.step(actions[sl], self._obs_views[sl.start], self._reward_views[sl.start])
                self._join()
            return self.obs, self.reward, done, {"step": self.envs[0].sim_step - 1}
        with self.envs[0].engine._guard():
            streams = self._fork()
            for env, sl, st in zip(self.envs, self.slices, streams):
                eng = env.engine
                eng._begin_steps(t, 1)
                eng._native_env_step(actions[sl] if self.n_act else None, self._obs_views[sl.start],
                                     self._reward_views[sl.start], env.cumulative_reward, t,
                                     stream=None if st is None else C.c_void_p(st.cuda_stream))
                eng.t_done = t
                env.sim_step = t + 1
            self._join()
        return self.obs, self.reward, t >= self.simulation_steps, {"step": t}

    def rollout(self, actions: torch.Tensor):
        """K decisions with given actions [K, R, n_act]: (obs [K, R, n_obs], reward [K, R], done)."""
        K = int(actions.shape[0])
        obs = torch.empty((K, self.R, self.n_obs), dtype=torch.float32, device=self.device)
        rew = torch.empty((K, self.R), dtype=torch.float32, device=self.device)
        done = False
        for k in range(K):
            o, r, done, _ = self.step(actions[k])
            obs[k].copy_(o)
            rew[k].copy_(r)
        return obs, rew, done

    def kpis(self, t_last: int = None) -> torch.Tensor:
        return torch.cat([env.kpis(t_last) for env in self.envs], dim=0)

    def check_errors(self):
        for env in self.envs:
            env.engine.check_errors()
