"""OD weights and origin demand (host-side setup; pre-drawn so the device step is deterministic).

Semantics follow the reference (src/LTM/od_manager.py): `ODManager` holds one weight array of
length S+1 per (origin, destination) (:22-50, all-ones when none are given); `DemandGenerator`
draws Poisson demand around two gaussian peaks at S/4 and 3S/4 with sigma S/20 (:145-155),
reseeding the *global legacy* numpy RNG per origin when `simulation.seed` is set (:153-154).
The draws go through `np.random.*` in the reference's order because the stream position is
observable in results (SURVEY.md "RNG ledger").
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Callable, Dict

import numpy as np


@dataclass
class DemandConfig:
    peak_lambda: float = 10.0
    base_lambda: float = 5.0
    seed: int = 42
    pattern: str = "gaussian_peaks"


class ODManager:
    """{(o, d): weight[S+1]} table; row t feeds the per-step route-choice kernel."""

    def __init__(self, simulation_steps: int, logger: logging.Logger = None):
        self.logger = logger or logging.getLogger(__name__)
        self.simulation_steps = simulation_steps
        self.od_flows: Dict[tuple, np.ndarray] = {}
        self._default_zero_flow = np.zeros(simulation_steps + 1)

    def init_od_flows(self, origin_nodes: list, destination_nodes: list, od_flows: dict = None):
        if od_flows:
            self._set_predefined_flows(od_flows)
            return
        self.logger.info("No OD flows provided, initializing with ones")
        for o in origin_nodes:
            for d in destination_nodes:
                if o != d:
                    self.od_flows[(o, d)] = np.ones(self.simulation_steps + 1)

    def _set_predefined_flows(self, od_flows: dict):
        n = self.simulation_steps + 1
        for (o, d), flow in od_flows.items():
            if isinstance(flow, (int, float)):
                self.od_flows[(o, d)] = np.full(n, flow)
            elif len(flow) != n:
                raise ValueError(f"Flow array length for OD pair ({o},{d}) must match simulation_steps")
            else:
                self.od_flows[(o, d)] = np.array(flow)

    def get_od_flow(self, origin: int, destination: int, time_step: int) -> float:
        return self.od_flows.get((origin, destination), self._default_zero_flow)[time_step]


class DemandGenerator:
    """Named demand patterns per origin: gaussian_peaks | constant | sudden_demand | custom."""

    def __init__(self, simulation_steps: int, params: dict, logger: logging.Logger):
        self.logger = logger
        self.simulation_steps = simulation_steps
        self.params = params
        self.time = np.arange(simulation_steps)
        self.seed = params.get("seed", None)
        self.demand_patterns: Dict[str, Callable] = {
            "gaussian_peaks": self.generate_gaussian_peaks,
            "constant": self.generate_constant,
            "sudden_demand": self.generate_sudden_demand,
        }

    def register_pattern(self, pattern_name: str, pattern_func: Callable):
        if not callable(pattern_func):
            raise ValueError("pattern_func must be callable")
        self.demand_patterns[pattern_name] = pattern_func

    def _get_demand_config(self, origin_id: int) -> DemandConfig:
        block = self.params.get("demand", {}).get(f"origin_{origin_id}")
        if block is None:
            # The reference logs through a logger that is None when verbose=False and dies with
            # AttributeError; falling back to the defaults is the evident intent.
            if self.logger is not None:
                self.logger.info(f"No demand configuration found for origin {origin_id}, using defaults")
            return DemandConfig()
        return DemandConfig(peak_lambda=block.get("peak_lambda", 10.0),
                            base_lambda=block.get("base_lambda", 5.0),
                            seed=self.seed,
                            pattern=block.get("pattern", "gaussian_peaks"))

    def generate_gaussian_peaks(self, origin_id: int, params=None) -> np.ndarray:
        return self._poisson_two_peaks(self._get_demand_config(origin_id))

    def generate_constant(self, origin_id: int, params=None) -> np.ndarray:
        return np.full(self.simulation_steps + 1, self._get_demand_config(origin_id).base_lambda)

    def generate_sudden_demand(self, origin_id: int, params=None) -> np.ndarray:
        demand = self._poisson_two_peaks(self._get_demand_config(origin_id))
        # three global-RNG integer draws, in this order (od_manager.py:119-121)
        period = np.random.randint(10, 20)
        start = np.random.randint(0, max(1, self.simulation_steps - period))
        demand[start:start + period] += np.random.randint(20, 50)
        return demand

    def generate_custom(self, origin_id: int, pattern: str) -> np.ndarray:
        if pattern not in self.demand_patterns:
            raise ValueError(f"Unknown demand pattern: {pattern}. "
                             f"Available patterns: {list(self.demand_patterns.keys())}")
        return self.demand_patterns[pattern](origin_id, params=self.params)

    def _poisson_two_peaks(self, cfg: DemandConfig) -> np.ndarray:
        S = self.simulation_steps
        width = 2 * (S / 20) ** 2
        lam = (cfg.base_lambda
               + cfg.peak_lambda * np.exp(-(self.time - S / 4) ** 2 / width)
               + cfg.peak_lambda * np.exp(-(self.time - 3 * S / 4) ** 2 / width))
        if self.seed is not None:
            np.random.seed(self.seed)
        return np.random.poisson(lam=lam)
