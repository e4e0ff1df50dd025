"""OD weights and origin demand: host-side setup, pre-drawn so that the device step is deterministic.

Behaviour is that of the reference's `src/LTM/od_manager.py` (public names kept so that callers can
switch): one weight series of length S+1 per (origin, destination), all ones when none are supplied
(:22-50); demand is Poisson around a base rate plus two gaussian bumps at S/4 and 3S/4 of width S/20
(:145-155).  Every random number goes through the *global legacy* numpy generator in the reference's
order -- including the per-origin reseeding when `params['seed']` is set (:153-154) and the three
integer draws of a sudden-demand burst (:119-121) -- because the position of that stream is visible in
the results (SURVEY.md "RNG ledger").
"""
from __future__ import annotations

import itertools
import logging
from dataclasses import dataclass
from typing import Callable, Dict

import numpy as np

BUILTIN_PATTERNS = ("gaussian_peaks", "constant", "sudden_demand")


@dataclass
class DemandConfig:
    peak_lambda: float = 10.0
    base_lambda: float = 5.0
    seed: int = 42
    pattern: str = "gaussian_peaks"


def two_peak_rate(steps: int, base: float, peak: float) -> np.ndarray:
    """Poisson rate per step: base + peak * (bump at S/4 + bump at 3S/4), bumps exp(-(t-c)^2 / (2 (S/20)^2))."""
    t = np.arange(steps)
    spread = 2 * (steps / 20) ** 2
    first = peak * np.exp(-(t - steps / 4) ** 2 / spread)
    second = peak * np.exp(-(t - 3 * steps / 4) ** 2 / spread)
    return base + first + second


class ODManager:
    """`od_flows[(o, d)]` = weight series [S+1]; row t of the stacked table feeds the route-choice kernel."""

    def __init__(self, simulation_steps: int, logger: logging.Logger = None):
        self.simulation_steps = simulation_steps
        self.logger = logger if logger is not None else logging.getLogger(__name__)
        self.od_flows: Dict[tuple, np.ndarray] = {}
        self._no_flow = np.zeros(simulation_steps + 1)

    def init_od_flows(self, origin_nodes: list, destination_nodes: list, od_flows: dict = None):
        length = self.simulation_steps + 1
        if not od_flows:
            self.logger.info("No OD flows provided, initializing with ones")
            for pair in itertools.product(origin_nodes, destination_nodes):
                if pair[0] != pair[1]:
                    self.od_flows[pair] = np.ones(length)
            return
        for pair, series in od_flows.items():
            if isinstance(series, (int, float)):                 # a constant weight
                self.od_flows[pair] = np.full(length, series)
                continue
            if len(series) != length:
                raise ValueError(f"Flow array length for OD pair ({pair[0]},{pair[1]}) must match simulation_steps")
            self.od_flows[pair] = np.array(series)

    def get_od_flow(self, origin: int, destination: int, time_step: int) -> float:
        series = self.od_flows.get((origin, destination))
        return (self._no_flow if series is None else series)[time_step]


class DemandGenerator:
    """Demand series of an origin by pattern name: the three built-ins or a registered callable
    `f(origin_id, params=...)`."""

    def __init__(self, simulation_steps: int, params: dict, logger: logging.Logger):
        self.simulation_steps = simulation_steps
        self.params = params
        self.logger = logger
        self.seed = params.get("seed", None)
        self.time = np.arange(simulation_steps)
        self.demand_patterns: Dict[str, Callable] = {name: getattr(self, "generate_" + name)
                                                     for name in BUILTIN_PATTERNS}

    def register_pattern(self, pattern_name: str, pattern_func: Callable):
        if not callable(pattern_func):
            raise ValueError("pattern_func must be callable")
        self.demand_patterns[pattern_name] = pattern_func

    def generate_custom(self, origin_id: int, pattern: str) -> np.ndarray:
        try:
            make = self.demand_patterns[pattern]
        except KeyError:
            raise ValueError(f"Unknown demand pattern: {pattern}. "
                             f"Available patterns: {list(self.demand_patterns.keys())}") from None
        return make(origin_id, params=self.params)

    # -- configuration -------------------------------------------------------------------------
    def _get_demand_config(self, origin_id: int) -> DemandConfig:
        entry = self.params.get("demand", {}).get(f"origin_{origin_id}")
        if entry is not None:
            return DemandConfig(entry.get("peak_lambda", 10.0), entry.get("base_lambda", 5.0), self.seed,
                                entry.get("pattern", "gaussian_peaks"))
        # The reference logs through a logger that is None when verbose=False and dies with
        # AttributeError; falling back to the defaults is the evident intent.
        if self.logger is not None:
            self.logger.info(f"No demand configuration found for origin {origin_id}, using defaults")
        return DemandConfig()

    # -- built-in patterns ---------------------------------------------------------------------
    def _draw(self, cfg: DemandConfig) -> np.ndarray:
        rate = two_peak_rate(self.simulation_steps, cfg.base_lambda, cfg.peak_lambda)
        if self.seed is not None:
            np.random.seed(self.seed)
        return np.random.poisson(lam=rate)

    def generate_gaussian_peaks(self, origin_id: int, params=None) -> np.ndarray:
        return self._draw(self._get_demand_config(origin_id))

    def generate_constant(self, origin_id: int, params=None) -> np.ndarray:
        level = self._get_demand_config(origin_id).base_lambda
        return np.full(self.simulation_steps + 1, level)

    def generate_sudden_demand(self, origin_id: int, params=None) -> np.ndarray:
        series = self._draw(self._get_demand_config(origin_id))
        # burst length, start and height: three integer draws from the global stream, in this order
        length = np.random.randint(10, 20)
        start = np.random.randint(0, max(1, self.simulation_steps - length))
        series[start:start + length] += np.random.randint(20, 50)
        return series
