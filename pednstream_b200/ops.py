"""PyTorch custom ops over the C-ABI (include/pns_b200.h).

The ops take the state tensors they mutate (so torch's dispatcher knows about the aliasing) plus
an integer handle to the engine that owns the prepared `pns_net` / `pns_state` / `pns_step_io`
structs; the implementation is one ctypes call into libpns_b200.so on torch's current CUDA stream.
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


@torch.library.custom_op("pednstream::ltm_step",
                         mutates_args=("hist64", "hist32", "runsum", "tf_routed", "probs", "err"))
def ltm_step(hist64: torch.Tensor, hist32: torch.Tensor, runsum: torch.Tensor, tf_routed: torch.Tensor,
             probs: torch.Tensor, err: torch.Tensor, handle: int, t0: int, n_steps: int,
             rng_mode: int) -> None:
    """network_loading(t) for t0 <= t < t0+n_steps (reference src/LTM/network.py:266-287)."""
    _engine(handle)._native_step(t0, n_steps, rng_mode)


@torch.library.custom_op("pednstream::ltm_draw_requests", mutates_args=("requests", "err"))
def ltm_draw_requests(hist64: torch.Tensor, hist32: torch.Tensor, requests: torch.Tensor,
                      err: torch.Tensor, handle: int, t: int) -> None:
    """Pass 1 of numpy-compatible stepping: which binomials step t will draw, with which n."""
    _engine(handle)._native_requests(t)


@torch.library.custom_op("pednstream::ltm_state_init",
                         mutates_args=("hist64", "hist32", "runsum", "err"))
def ltm_state_init(hist64: torch.Tensor, hist32: torch.Tensor, runsum: torch.Tensor, err: torch.Tensor,
                   handle: int) -> None:
    """Initial history/width state of every link (reference src/LTM/link.py:12-17, 32-100)."""
    _engine(handle)._native_init()
