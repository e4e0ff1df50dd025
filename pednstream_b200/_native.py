"""ctypes binding of the C-ABI in include/pns_b200.h (library: pednstream_b200/lib/libpns_b200.so).

There is no fallback: if the CUDA library has not been built (`python -c "import
__graft_entry__ as g; g.build()"`) loading raises, and so does every simulation call.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpns_b200.so")
ABI_VERSION = 8
METRIC_SLOTS, METRIC_STRIDE = 64, 4          # PNS_METRIC_SLOTS / PNS_METRIC_STRIDE of pns_step_streamed
METRIC_ROW = METRIC_SLOTS * METRIC_STRIDE
KPI_NAMES = ("total_demand", "total_outflow", "total_inflow", "person_time", "person_time_moving", "total_delay",
             "congestion_time", "area_time", "steps", "congested_steps", "avg_travel_time")   # PNS_KPI_* order

RNG_TABLE, RNG_PHILOX, RNG_REQUEST = 0, 1, 2
ERR_BITS = {1: "negative sending flow (reference link.py:346,366 ValueError)",
            2: "negative node flow (reference node.py:194,219,238 Warning)",
            4: "history index out of range (numpy IndexError in reference link.py:210-212)",
            8: "zero travel-time lag: the reference result depends on node visiting order",
            16: "'optimal' node model: the linear program of a node did not reach an optimum (reference node.py:265)"}

_p = C.c_void_p
_i32 = C.c_int32


class PnsLinkClass(C.Structure):
    _fields_ = ([(k, C.c_double) for k in ("length", "area", "space", "kc", "vf", "kj", "act", "sigma")]
                + [(k, C.c_float) for k in ("kc32", "kj32", "kj_minus_kc32", "gamma32", "bi32", "area32", "length32",
                                            "yp_coef32", "neg_vf32", "vf32", "sm_gamma32", "inv_kj32", "max_tt32", "tt0")]
                + [(k, _i32) for k in ("fftau", "swtau", "flags", "pad_")])


class PnsNet(C.Structure):
    _fields_ = (
        [(n, _i32) for n in ("abi_version", "n_links", "n_nodes", "n_cols64", "sim_steps", "replicas",
                             "window", "n_edges", "n_od", "n_demand_rows",
                             "n_routed", "n_groups", "n_opts", "n_rows", "n_terms", "n_classes", "max_degree", "nd_stride")]
        + [("unit_time", C.c_double)]
        + [("classes", _p), ("class0", PnsLinkClass)]
        + [(n, _p) for n in ("lk_class", "lk_width", "nd_meta", "nd_routed", "lk_slots", "nd_in_link",
                             "rt_routed_nodes", "rt_routed_edge0", "rt_routed_row0", "rt_row_routed", "rt_row_grp_ptr", "rt_row_grp", "rt_term_od", "rt_dyn_rows",
                             "rt_grp_node", "rt_grp_up", "rt_grp_od", "rt_grp_has_virtual", "rt_opt_ptr",
                             "rt_opt_link", "rt_opt_slot", "rt_opt_dist",
                             "rt_row_ptr", "rt_row_od", "rt_term_ptr", "rt_term_opt", "rt_term_row_entry")]
        + [(n, C.c_double) for n in ("rt_temp", "rt_alpha", "rt_beta", "rt_omega", "rt_eps")]
        + [("lane_order", _p), ("lane_order_block", _i32), ("n_lane_blocks", _i32),
           ("per_replica_scenario", _i32), ("n_dyn_rows", _i32),
           ("lp_nodes", _p), ("n_lp_nodes", _i32), ("lp_max_m", _i32), ("lp_w", C.c_double),
           ("nd_cols_order", _p), ("n_nodes_small", _i32), ("pad3_", _i32)]
    )


class PnsState(C.Structure):
    _fields_ = [(n, _p) for n in ("hist64", "hist32", "gate", "sep_np64", "runsum", "tf_static",
                                  "tf_routed", "probs", "nm_s", "nm_r", "err")] + [("n_f64", _i32)]


class PnsStepIO(C.Structure):
    _fields_ = [("demand", _p), ("od_w", _p), ("draw_b", _p), ("draw_n", _p),
                ("draw_row_stride", C.c_int64),
                ("req_kind", _p), ("req_n1", _p), ("req_rf", _p), ("req_sval", _p), ("req_n3", _p),
                ("req_exp", _p), ("draw_exp", _p),
                ("seed", C.c_uint64), ("replica_base", C.c_uint32), ("route_all_rows", C.c_uint32),
                ("lp_x", _p)]


class PnsEnv(C.Structure):
    _fields_ = [("n_act", _i32), ("n_obs", _i32), ("n_reward_links", _i32), ("pad_", _i32)] + \
               [(k, _p) for k in ("act_link", "act_sep", "act_lo", "act_hi", "act_max_delta", "act_total_width",
                                  "obs_link", "obs_src", "obs_div", "reward_link",
                                  "lk_act", "lk_obs_ptr", "lk_obs_col", "lk_obs_src", "lk_obs_div", "lk_reward",
                                  "reward_count")]


OBS_SRC = {"inflow": 0, "outflow": 1, "rev.inflow": 2, "rev.outflow": 3, "gdens": 4, "gdens/kjam": 5,
           "speed": 6, "gate": 7}


EXPORTS = ("pns_abi_version", "pns_last_error", "pns_state_init", "pns_link_flows", "pns_route_fractions",
           "pns_node_flows", "pns_link_update", "pns_step", "pns_step_profiled", "pns_step_streamed", "pns_env_apply_actions", "pns_env_observe", "pns_env_step", "pns_env_rollout", "pns_env_draw_demand", "pns_env_randomize", "pns_kpi",
           "pns_lp_solve", "pns_lane_block_size", "pns_rng_selftest")

_LIB = None


def _declare(lib):
    net_p, st_p, io_p = C.POINTER(PnsNet), C.POINTER(PnsState), C.POINTER(PnsStepIO)
    lib.pns_abi_version.restype = C.c_int
    lib.pns_last_error.restype = C.c_char_p
    lib.pns_state_init.argtypes = [net_p, st_p, _p]
    lib.pns_link_flows.argtypes = [net_p, st_p, io_p, C.c_int, C.c_int, _p]
    lib.pns_route_fractions.argtypes = [net_p, st_p, io_p, C.c_int, C.c_int, _p]
    lib.pns_node_flows.argtypes = [net_p, st_p, io_p, C.c_int, _p]
    lib.pns_link_update.argtypes = [net_p, st_p, io_p, C.c_int, C.c_int, _p]
    lib.pns_step.argtypes = [net_p, st_p, io_p, C.c_int, C.c_int, C.c_int, _p]
    lib.pns_step_profiled.argtypes = [net_p, st_p, io_p, C.c_int, C.c_int, C.c_int, _p, _p, _p]
    lib.pns_step_streamed.argtypes = [net_p, st_p, io_p, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p]
    env_p = C.POINTER(PnsEnv)
    lib.pns_env_apply_actions.argtypes = [net_p, st_p, env_p, _p, _p]
    lib.pns_env_draw_demand.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint64, _p, _p, _p, _p, _p, _p, _p]
    lib.pns_env_randomize.argtypes = [net_p, _p, C.c_int, _p, C.c_int, _p, _p, C.c_int, _p, _p, _p, _p, C.c_uint64,
                                      C.c_uint32, _p]
    lib.pns_kpi.argtypes = [net_p, st_p, io_p, C.c_int, _p, C.c_int, _p, _p, _p]
    lib.pns_env_step.argtypes = [net_p, st_p, io_p, env_p, _p, C.c_int, C.c_int, _p, _p, _p, _p]
    lib.pns_env_rollout.argtypes = [net_p, st_p, io_p, env_p, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p, _p, _p, _p, _p]
    lib.pns_env_observe.argtypes = [net_p, st_p, env_p, C.c_int, _p, _p, _p]
    lib.pns_lp_solve.argtypes = [C.c_int, C.c_int, _p, _p, _p, C.c_double, _p, _p, _p, _p]
    lib.pns_rng_selftest.argtypes = [C.c_int, C.c_int, _p, _p, C.c_uint64, C.c_int, C.c_int, _p, _p, _p]
    for name in EXPORTS[2:]:
        getattr(lib, name).restype = C.c_int
    return lib


def load(path: str = None):
    """Load (once) and return the native library; raises if it is missing or ABI-incompatible."""
    global _LIB
    if path is None and _LIB is not None:
        return _LIB
    target = path or os.environ.get("PNS_B200_LIB") or LIB_PATH     # env override: kernel tuning builds
    if not os.path.exists(target):
        raise RuntimeError(
            f"pednstream_b200: native CUDA library not found at {target}. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            f"There is no CPU fallback for the simulation step.")
    lib = _declare(C.CDLL(target))
    if lib.pns_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{target}: ABI version {lib.pns_abi_version()} != expected {ABI_VERSION}")
    if path is None:
        _LIB = lib
    return lib


def check(lib, rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib.pns_last_error().decode()}")
