"""PyTorch custom ops over the C-ABI (include/pns_b200.h): `torch.ops.pednstream.*`.

The ops take the state tensors they mutate (so torch's dispatcher knows about the aliasing) plus
an integer handle to the engine that owns the prepared `pns_net` / `pns_state` / `pns_step_io`
structs; the implementation is one ctypes call into libpns_b200.so on torch's current CUDA stream.
They are declared with `torch.library.Library` schemas: the dispatcher adds ~5 us to a call, where the
`torch.library.custom_op` decorator costs 20-40 us -- as much as a whole environment step of 1024 replicas.
"""
from __future__ import annotations

import weakref

import torch

_ENGINES = weakref.WeakValueDictionary()
_NEXT = [1]


def register_engine(engine) -> int:
    h = _NEXT[0]
    _NEXT[0] += 1
    _ENGINES[h] = engine
    return h


def _engine(handle: int):
    eng = _ENGINES.get(handle)
    if eng is None:
        raise RuntimeError(f"pednstream: engine handle {handle} is no longer alive")
    return eng


_LIB = torch.library.Library("pednstream", "DEF")


def _define(schema, fn):
    _LIB.define(schema)
    name = schema.split("(", 1)[0]
    _LIB.impl(name, fn, "CompositeExplicitAutograd")
    return getattr(torch.ops.pednstream, name)


def _ltm_step(hist64, hist32, runsum, tf_routed, probs, err, handle, t0, n_steps, rng_mode):
    """network_loading(t) for t0 <= t < t0+n_steps (reference src/LTM/network.py:266-287)."""
    _engine(handle)._native_step(t0, n_steps, rng_mode)


def _ltm_draw_requests(hist64, hist32, requests, err, handle, t):
    """Pass 1 of numpy-compatible stepping: which binomials step t will draw, with which n (and the arguments
    of the step's route-choice exponentials)."""
    _engine(handle)._native_requests(t)


def _ltm_state_init(hist64, hist32, runsum, err, handle):
    """Initial history/width state of every link (reference src/LTM/link.py:12-17, 32-100)."""
    _engine(handle)._native_init()


def _ltm_step_streamed(hist64, hist32, runsum, tf_routed, probs, err, dev_metric, handle, t0, n_steps, rng_mode):
    """network_loading(t) for t0 <= t < t0+n_steps with the per-step host traffic of a driving loop folded in
    (C-ABI pns_step_streamed): the pinned host demand / result tensors were handed to the engine beforehand."""
    _engine(handle)._native_step_streamed(t0, n_steps, rng_mode)


def _env_step(hist64, hist32, runsum, tf_routed, probs, err, gate, actions, obs, reward, cum_reward, handle, t,
              has_actions):
    """One control-environment step over all replicas (reference rl/pz_pednet_env.py:194-254, action_gap 1):
    actions, network_loading(t), observations and reward (C-ABI pns_env_step)."""
    _engine(handle)._native_env_step(actions if has_actions else None, obs, reward, cum_reward, t)


def _env_rollout(hist64, hist32, runsum, tf_routed, probs, err, gate, actions, obs, reward, cum_reward, handle, t0,
                 n_steps, has_actions):
    """n_steps environment steps with given device-resident actions [n_steps, R, n_act] in one native call
    (C-ABI pns_env_rollout): obs [n_steps, R, n_obs], reward [n_steps, R]."""
    _engine(handle)._native_env_rollout(actions if has_actions else None, obs, reward, cum_reward, t0, n_steps)


def _episode_kpis(hist64, hist32, demand, role, scratch, out, handle, t_last, any_od_path):
    """Per-replica episode KPIs from history rows 0..t_last (reference rl/rl_utils.py:770-1512; C-ABI pns_kpi)."""
    _engine(handle)._native_kpi(role, scratch, out, t_last, any_od_path)


ltm_step = _define(
    "ltm_step(Tensor(a!) hist64, Tensor(b!) hist32, Tensor(c!) runsum, Tensor(d!) tf_routed, Tensor(e!) probs, "
    "Tensor(f!) err, int handle, int t0, int n_steps, int rng_mode) -> ()", _ltm_step)
ltm_draw_requests = _define(
    "ltm_draw_requests(Tensor hist64, Tensor hist32, Tensor(a!) requests, Tensor(b!) err, int handle, int t) -> ()",
    _ltm_draw_requests)
ltm_state_init = _define(
    "ltm_state_init(Tensor(a!) hist64, Tensor(b!) hist32, Tensor(c!) runsum, Tensor(d!) err, int handle) -> ()",
    _ltm_state_init)
ltm_step_streamed = _define(
    "ltm_step_streamed(Tensor(a!) hist64, Tensor(b!) hist32, Tensor(c!) runsum, Tensor(d!) tf_routed, "
    "Tensor(e!) probs, Tensor(f!) err, Tensor(g!) dev_metric, int handle, int t0, int n_steps, int rng_mode) -> ()",
    _ltm_step_streamed)
env_step = _define(
    "env_step(Tensor(a!) hist64, Tensor(b!) hist32, Tensor(c!) runsum, Tensor(d!) tf_routed, Tensor(e!) probs, "
    "Tensor(f!) err, Tensor(g!) gate, Tensor actions, Tensor(h!) obs, Tensor(i!) reward, Tensor(j!) cum_reward, "
    "int handle, int t, bool has_actions) -> ()", _env_step)
env_rollout = _define(
    "env_rollout(Tensor(a!) hist64, Tensor(b!) hist32, Tensor(c!) runsum, Tensor(d!) tf_routed, Tensor(e!) probs, "
    "Tensor(f!) err, Tensor(g!) gate, Tensor actions, Tensor(h!) obs, Tensor(i!) reward, Tensor(j!) cum_reward, "
    "int handle, int t0, int n_steps, bool has_actions) -> ()", _env_rollout)
episode_kpis = _define(
    "episode_kpis(Tensor hist64, Tensor hist32, Tensor demand, Tensor role, Tensor(a!) scratch, Tensor(b!) out, "
    "int handle, int t_last, bool any_od_path) -> ()", _episode_kpis)
