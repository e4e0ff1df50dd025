"""Action / observation spaces (reference: rl/spaces.py:20-104).

gymnasium is optional: when it is not installed a minimal `Box` with the same attributes
(`low`, `high`, `shape`, `dtype`, `sample()`, `contains()`) is used.
"""
from __future__ import annotations

import numpy as np

try:                                   # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box
except Exception:                      # gymnasium absent: duck-typed stand-in
    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return np.random.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class SpaceBuilder:
    """gate agent: Box([0]*k, [link.width]*k); separator agent: Box(min_w, W - min_w, (1,))."""

    def __init__(self, agent_manager, obs_mode: str, min_sep_width: float = 1.0):
        self.agent_manager = agent_manager
        self.min_sep_width = min_sep_width
        self.sep_obs_dim = 4
        self.gat_obs_dim_per_link = None

    def build_action_spaces(self):
        am, out = self.agent_manager, {}
        for aid in am.get_separator_agents():
            fwd, _ = am.get_separator_links(aid)
            out[aid] = Box(low=self.min_sep_width, high=fwd.width - self.min_sep_width, shape=(1,),
                           dtype=np.float32)
        for aid in am.get_gater_agents():
            links = am.get_gater_outgoing_links(aid)
            out[aid] = Box(low=np.zeros(len(links), dtype=np.float32),
                           high=np.array([l.width for l in links], dtype=np.float32),
                           shape=(len(links),), dtype=np.float32)
        return out

    def build_observation_spaces(self, features_per_link: int):
        am, out = self.agent_manager, {}
        self.gat_obs_dim_per_link = features_per_link
        for aid in am.get_separator_agents():
            out[aid] = Box(low=-np.inf, high=np.inf, shape=(self.sep_obs_dim,), dtype=np.float32)
        for aid in am.get_gater_agents():
            out[aid] = Box(low=-np.inf, high=np.inf,
                           shape=(am.get_max_outdegree(aid) * features_per_link,), dtype=np.float32)
        return out

    def get_separator_obs_dim(self):
        return self.sep_obs_dim

    def get_gater_obs_dim_per_link(self):
        return self.gat_obs_dim_per_link
