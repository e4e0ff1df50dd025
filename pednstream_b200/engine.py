"""Device runtime: owns the torch tensors behind `pns_net` / `pns_state` and drives the step.

Three draw modes (SURVEY.md "RNG ledger"):

numpy   per step: REQUEST pass on the GPU -> the host draws the step's binomials/normals from
        numpy's *global legacy* RNG in the reference's visiting order -> TABLE step on the GPU.
        Same `np.random.seed` => bit-identical trajectory to the reference.  One host sync per step.
table   outcomes for many steps supplied up front (replay); no host involvement per step.
philox  counter-based sampling on the device; no host involvement per step.

All physics runs in the CUDA kernels; the host only supplies random numbers (numpy mode), the
demand row and width/fraction edits made between steps.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native, ops
from .state import F32_INDEX, F64_FIELDS, F64_INDEX

_NET_ARRAYS = ("lk_class", "lk_width", "nd_meta", "nd_routed", "lk_slots", "nd_in_link",
               "rt_routed_nodes", "rt_routed_edge0", "rt_routed_row0", "rt_row_routed", "rt_row_grp_ptr", "rt_row_grp", "rt_term_od", "rt_dyn_rows", "rt_grp_node", "rt_grp_up",
               "rt_grp_od", "rt_grp_has_virtual", "rt_opt_ptr", "rt_opt_link", "rt_opt_slot",
               "rt_opt_dist", "rt_row_ptr", "rt_row_od", "rt_term_ptr", "rt_term_opt",
               "rt_term_row_entry")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)


class Engine:
    def __init__(self, plan: dict, replicas: int = 1, device=None, rng: str = "numpy", seed: int = 0,
                 lib=None, emulation: bool = False):
        self.plan = plan
        self.R = int(replicas)
        self.rng = rng
        self.seed = int(seed)
        if rng not in ("numpy", "philox", "table"):
            raise ValueError(f"unknown rng mode {rng!r}")
        self.lib = lib if lib is not None else _native.load()
        if emulation:
            self.device = torch.device("cpu")
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("pednstream_b200 needs a CUDA device: the LTM step has no CPU path")
            self.device = torch.device(device if device is not None else "cuda")
            if self.device.type != "cuda":
                raise RuntimeError("pednstream_b200: device must be a CUDA device")
            if self.device.index is None:
                self.device = torch.device("cuda", torch.cuda.current_device())
        self.emulation = emulation
        dev = self.device
        p = plan
        self.L, self.N, self.S = p["n_links"], p["n_nodes"], p["sim_steps"]
        self.C64 = self.L + p["n_virtual"]
        self.n_f64 = 8 if p["has_separators"] else 7
        R, L, S = self.R, self.L, self.S

        # ---- immutable plan ------------------------------------------------------------------
        self._net_t = {k: torch.from_numpy(np.ascontiguousarray(p[k])).to(dev) for k in _NET_ARRAYS}
        cls_bytes = np.ascontiguousarray(p["classes"]).view(np.uint8).reshape(-1)
        self._net_t["classes"] = torch.from_numpy(cls_bytes.copy()).to(dev)
        self._meta_host = np.ascontiguousarray(p["nd_meta"]).copy()
        net = _native.PnsNet()
        net.abi_version = _native.ABI_VERSION
        net.n_links, net.n_nodes, net.n_cols64 = L, self.N, self.C64
        net.sim_steps, net.replicas, net.window = S, R, p["window"]
        net.n_edges, net.n_od, net.n_demand_rows = p["n_edges"], p["n_od"], p["n_demand_rows"]
        net.n_routed = len(p["rt_routed_nodes"])
        net.n_groups = len(p["rt_grp_node"])
        net.n_opts = len(p["rt_opt_link"])
        net.n_rows = len(p["rt_row_ptr"]) - 1
        net.n_terms = len(p["rt_term_opt"])
        net.n_dyn_rows = len(p["rt_dyn_rows"])
        net.n_classes = len(p["classes"])
        net.max_degree = int(p["max_degree"])
        net.nd_stride = int(p["nd_stride"])
        net.unit_time = p["unit_time"]
        for k in _NET_ARRAYS + ("classes",):
            setattr(net, k, _ptr(self._net_t[k]))
        C.memmove(C.byref(net.class0), np.ascontiguousarray(p["classes"][:1]).ctypes.data, C.sizeof(net.class0))
        net.rt_temp, net.rt_alpha, net.rt_beta, net.rt_omega, net.rt_eps = [float(x) for x in p["rt_scalars"]]
        # assign_flows_type 'optimal': the nodes that solve the linear program (kind 2 in nd_meta)
        kinds = (np.asarray(p["nd_meta"])[:, 1] >> 8) & 0xff if self.N else np.zeros(0, dtype=np.int64)
        lp_nodes = np.nonzero(kinds == 2)[0].astype(np.int32)
        self._lp_nodes = torch.from_numpy(lp_nodes if len(lp_nodes) else np.zeros(1, np.int32)).to(dev)
        net.lp_nodes, net.n_lp_nodes = _ptr(self._lp_nodes), len(lp_nodes)
        net.lp_max_m = int((np.asarray(p["nd_meta"])[lp_nodes, 1] & 0xff).max()) if len(lp_nodes) else 0
        net.lp_w = float(p.get("lp_w", 0.01))
        self.lp_x = None                                  # keep_lp_solutions(): the step's turn flows before the floor
        # schedule of the batched node kernel (nd_stride 8): nodes of <= 4 slots first, they share a CTA in twos
        self._cols_order = None
        if R > 1 and not emulation and int(p["nd_stride"]) > 4 and self.N:
            slots = np.asarray(p["nd_meta"])[:, 1] & 0xff
            order = np.concatenate([np.nonzero(slots <= 4)[0], np.nonzero(slots > 4)[0]]).astype(np.int32)
            self._cols_order = torch.from_numpy(order).to(dev)
            net.nd_cols_order, net.n_nodes_small = _ptr(self._cols_order), int((slots <= 4).sum())
        # launch order of the single-replica link kernel (a schedule; results do not depend on it)
        self._lane_order = None
        if R == 1 and not emulation and L > 0:
            from .plan import lane_block_order
            block = int(self.lib.pns_lane_block_size())
            order = lane_block_order(p["nd_meta"], p["nd_in_link"], L, int(p["nd_stride"]), block)
            if order is not None:
                self._lane_order = torch.from_numpy(order).to(dev)
                net.lane_order, net.lane_order_block, net.n_lane_blocks = _ptr(self._lane_order), block, len(order)
        self.net = net

        # ---- mutable state ---------------------------------------------------------------------
        self.hist64 = torch.empty((self.n_f64, S + 1, self.C64 * R), dtype=torch.float64, device=dev)
        self.hist32 = torch.empty((6, S + 1, L * R), dtype=torch.float32, device=dev)
        self.gate = torch.zeros((L * R,), dtype=torch.float64, device=dev)
        self.sep_np64 = torch.zeros((L * R,), dtype=torch.int32, device=dev)
        self.runsum = torch.zeros((L * R,), dtype=torch.float32, device=dev)
        self.tf_static = torch.zeros((max(1, p["n_edges"]),), dtype=torch.float64, device=dev)
        self.tf_routed = torch.zeros((max(1, p["n_edges"]) * R,), dtype=torch.float64, device=dev)
        self.probs = torch.zeros((max(1, net.n_opts) * R,), dtype=torch.float64, device=dev)
        self.err = torch.zeros((R,), dtype=torch.int32, device=dev)
        nm = max(1, self.N * int(p["nd_stride"]) * R)       # node-major exchange arrays (link pass -> node pass)
        self.nm_s = torch.zeros((nm,), dtype=torch.float64, device=dev)
        self.nm_r = torch.zeros((nm,), dtype=torch.float64, device=dev)
        st = _native.PnsState()
        for k in ("hist64", "hist32", "gate", "sep_np64", "runsum", "tf_static", "tf_routed", "probs",
                  "nm_s", "nm_r", "err"):
            setattr(st, k, _ptr(getattr(self, k)))
        st.n_f64 = self.n_f64
        self.state = st

        # ---- per-step inputs --------------------------------------------------------------------
        rows = p["n_demand_rows"]
        self.demand = torch.zeros((S + 1, max(1, rows) * R), dtype=torch.float64, device=dev)
        self.od_w = torch.zeros((S + 1, max(1, p["n_od"])), dtype=torch.float64, device=dev)
        n32 = L * R
        # request block (device + pinned host mirror): kind, n1, n3 (int32) | rf (float32) | sval (float64)
        self._req = torch.zeros((6 * n32,), dtype=torch.int32, device=dev)
        self._req_host = torch.zeros((6 * n32,), dtype=torch.int32, pin_memory=not emulation)
        # draw block: R1, R2, R3 outcomes (int32) + one pad row | noise (float64)
        self._draw = torch.zeros((6 * n32,), dtype=torch.int32, device=dev)
        self._draw_host = torch.zeros((6 * n32,), dtype=torch.int32, pin_memory=not emulation)
        io = _native.PnsStepIO()
        io.demand, io.od_w = _ptr(self.demand), _ptr(self.od_w)
        b = self._req.data_ptr()
        io.req_kind, io.req_n1, io.req_n3 = b, b + 4 * n32, b + 8 * n32
        io.req_rf, io.req_sval = b + 12 * n32, b + 16 * n32
        d = self._draw.data_ptr()
        io.draw_b, io.draw_n = d, d + 16 * n32
        io.draw_row_stride = 0
        io.seed = self.seed
        # route-choice exponentials of numpy-compatible stepping: arguments out, numpy's exp back in
        self._exp_arg = self._exp_val = self._exp_arg_host = self._exp_val_host = None
        if rng == "numpy" and net.n_opts > 0:
            n_exp = net.n_opts * R
            self._exp_arg = torch.zeros((n_exp,), dtype=torch.float64, device=dev)
            self._exp_val = torch.zeros((n_exp,), dtype=torch.float64, device=dev)
            self._exp_arg_host = torch.zeros((n_exp,), dtype=torch.float64, pin_memory=not emulation)
            self._exp_val_host = torch.zeros((n_exp,), dtype=torch.float64, pin_memory=not emulation)
            io.req_exp, io.draw_exp = _ptr(self._exp_arg), _ptr(self._exp_val)
        self.io = io
        self._table_io = None
        self._od_per_replica = False
        self._last_row = 0

        self.handle = ops.register_engine(self)
        self.t_done = 0
        self._net_ref = None
        self._initialised = False

    # ------------------------------------------------------------------ native calls
    def _stream(self):
        if self.emulation:
            return C.c_void_p(0)
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _guard(self):
        """Make the engine's device current for the native launches (no-op when it already is)."""
        if self.emulation or torch.cuda.current_device() == self.device.index:
            return _NULL
        return torch.cuda.device(self.device)

    def _native_init(self):
        _native.check(self.lib, self.lib.pns_state_init(C.byref(self.net), C.byref(self.state), self._stream()),
                      "pns_state_init")

    def _native_requests(self, t):
        _native.check(self.lib, self.lib.pns_link_flows(C.byref(self.net), C.byref(self.state), C.byref(self.io),
                                                        t, _native.RNG_REQUEST, self._stream()), "pns_link_flows")

    def _native_step(self, t0, n_steps, rng_mode):
        io = self._table_io if (rng_mode == _native.RNG_TABLE and self._table_io is not None) else self.io
        _native.check(self.lib, self.lib.pns_step(C.byref(self.net), C.byref(self.state), C.byref(io),
                                                  t0, n_steps, rng_mode, self._stream()), "pns_step")
        self._route_all(False)

    def _native_step_streamed(self, t0, n_steps, rng_mode):
        host_demand, host_metric = self._streamed_host
        _native.check(self.lib, self.lib.pns_step_streamed(
            C.byref(self.net), C.byref(self.state), C.byref(self.io), t0, n_steps, rng_mode,
            _ptr(host_demand), _ptr(self._dev_metric), _ptr(host_metric), self._stream()), "pns_step_streamed")
        self._route_all(False)

    def _native_env_step(self, actions, obs, reward, cum_reward, t, stream=None):
        _native.check(self.lib, self.lib.pns_env_step(
            C.byref(self.net), C.byref(self.state), C.byref(self.io), C.byref(self._env_struct),
            _ptr(actions) if actions is not None else C.c_void_p(0), int(t), _native.RNG_PHILOX, _ptr(obs), _ptr(reward),
            _ptr(cum_reward), self._stream() if stream is None else stream), "pns_env_step")
        self._route_all(False)

    def _native_env_rollout(self, actions, obs, reward, cum_reward, t0, n_steps, host=None):
        h = [_ptr(x) if x is not None else C.c_void_p(0) for x in (host or (None, None, None))]
        _native.check(self.lib, self.lib.pns_env_rollout(
            C.byref(self.net), C.byref(self.state), C.byref(self.io), C.byref(self._env_struct), int(t0), int(n_steps),
            _native.RNG_PHILOX, _ptr(actions) if actions is not None else C.c_void_p(0), _ptr(obs), _ptr(reward),
            _ptr(cum_reward), h[0], h[1], h[2], self._stream()), "pns_env_rollout")
        self._route_all(False)

    def _native_kpi(self, role, scratch, out, t_last, any_od_path):
        _native.check(self.lib, self.lib.pns_kpi(C.byref(self.net), C.byref(self.state), C.byref(self.io),
                                                 int(t_last), _ptr(role), int(bool(any_od_path)), _ptr(scratch),
                                                 _ptr(out), self._stream()), "pns_kpi")

    # ------------------------------------------------------------------ setup
    def initialise(self, gate: np.ndarray, sep_np64: np.ndarray = None, tf_static: np.ndarray = None,
                   demand: np.ndarray = None, od_w: np.ndarray = None, tf_supplied: np.ndarray = None):
        """gate [L] (broadcast over replicas) or [L*R]: back gate width of plain links / lane width of
        separators; tf_static [n_edges] with tf_supplied[node] marking user-supplied fractions;
        demand [T, rows] or [T, rows*R]."""
        with self._guard():
            self.set_gate(gate, sep_np64)
            self.set_static_fractions(tf_static, tf_supplied)
            if demand is not None:
                self.set_demand(demand)
            if od_w is not None and od_w.size and not self._od_per_replica:
                self.od_w[: od_w.shape[0], : od_w.shape[1]].copy_(torch.from_numpy(np.ascontiguousarray(od_w)))
            ops.ltm_state_init(self.hist64, self.hist32, self.runsum, self.err, self.handle)
            self.nm_s.zero_()
            self.nm_r.zero_()
        self.t_done = 0
        self._last_row = 0
        self._initialised = True
        self._route_all(True)

    def _route_all(self, flag: bool):
        """Whether the next step call evaluates every route row (constant rows included) or only the dynamic ones."""
        self.io.route_all_rows = 1 if flag else 0
        if self._table_io is not None:
            self._table_io.route_all_rows = self.io.route_all_rows

    def _begin_steps(self, t0: int, n_steps: int):
        """Row contract of the node pass (pns_b200.h, pns_node_flows): inflow/outflow rows of a step must be
        zero before it runs.  Moving forward in time that holds by construction; when asked to repeat steps,
        clear the rows from t0 up to the last row ever written."""
        last = self._last_row
        if t0 > last:                         # the normal case: forward in time
            self._last_row = t0 + n_steps - 1
            return
        lo, hi = max(int(t0), 0), min(last, self.S)
        self.hist64[F64_INDEX["inflow"], lo:hi + 1].zero_()
        self.hist64[F64_INDEX["outflow"], lo:hi + 1].zero_()
        self.nm_s.zero_()        # the link pass skips a hand-over store when the slot already holds its 0
        self._last_row = max(last, int(t0) + int(n_steps) - 1)

    def set_replica_scenarios(self, classes: np.ndarray, lk_class: np.ndarray, od_w: np.ndarray = None):
        """Per-replica scenarios (domain randomisation): `classes` a table of pns_link_class records
        (plan.CLASS_DTYPE), `lk_class` [L, R] the class of every link in every replica, `od_w`
        [S+1, n_od, R] the OD weights of every replica (None: shared weights stay).  Call before
        `initialise` (initial travel times come from the classes)."""
        lk_class = np.ascontiguousarray(lk_class, dtype=np.int32)
        if lk_class.shape != (self.L, self.R):
            raise ValueError(f"lk_class must be [{self.L}, {self.R}]")
        if lk_class.min() < 0 or lk_class.max() >= len(classes):
            raise ValueError("lk_class refers to a class outside the table")
        dev = self.device
        cls_bytes = np.ascontiguousarray(classes).view(np.uint8).reshape(-1)
        self._net_t["classes"] = torch.from_numpy(cls_bytes.copy()).to(dev)
        self._net_t["lk_class"] = torch.from_numpy(lk_class.reshape(-1)).to(dev)
        net = self.net
        net.classes, net.lk_class = _ptr(self._net_t["classes"]), _ptr(self._net_t["lk_class"])
        net.n_classes = len(classes)
        C.memmove(C.byref(net.class0), np.ascontiguousarray(classes[:1]).ctypes.data, C.sizeof(net.class0))
        net.per_replica_scenario = 1
        n_od = int(self.plan["n_od"])
        if od_w is not None and n_od:
            od_w = np.ascontiguousarray(od_w, dtype=np.float64)
            if od_w.shape != (self.S + 1, n_od, self.R):
                raise ValueError(f"od_w must be [{self.S + 1}, {n_od}, {self.R}]")
            self.od_w = torch.from_numpy(od_w.reshape(self.S + 1, -1)).to(dev)
        elif n_od and not self._od_per_replica:
            self.od_w = self.od_w.repeat_interleave(self.R, dim=1).contiguous()   # shared weights, per-replica layout
        self._od_per_replica = True
        self.io.od_w = _ptr(self.od_w)
        if self._table_io is not None:
            self._table_io.od_w = _ptr(self.od_w)

    def set_replica_scenarios_device(self, classes: torch.Tensor, n_classes: int, lk_class: torch.Tensor,
                                     od_w: torch.Tensor = None):
        """Per-replica scenarios whose tables already live on the device (pns_env_randomize): `classes` a uint8
        tensor of n_classes pns_link_class records, `lk_class` int32 [L*R] (replica fastest), `od_w` float64
        [S+1, n_od*R].  Call before `initialise`."""
        self._net_t["classes"], self._net_t["lk_class"] = classes, lk_class
        net = self.net
        net.classes, net.lk_class = _ptr(classes), _ptr(lk_class)
        net.n_classes = int(n_classes)
        net.per_replica_scenario = 1
        if od_w is not None:
            self.od_w = od_w
            self.io.od_w = _ptr(od_w)
            if self._table_io is not None:
                self._table_io.od_w = _ptr(od_w)
        self._od_per_replica = True

    def set_gate(self, gate, sep_np64=None):
        g = torch.from_numpy(np.ascontiguousarray(gate, dtype=np.float64).reshape(-1))
        if g.shape[0] == self.L and self.R > 1:
            g = g.repeat_interleave(self.R)
        self.gate.copy_(g)
        if sep_np64 is not None:
            f = torch.from_numpy(np.ascontiguousarray(sep_np64, dtype=np.int32))
            if f.shape[0] == self.L and self.R > 1:
                f = f.repeat_interleave(self.R)
            self.sep_np64.copy_(f)

    def set_static_fractions(self, tf_static, tf_supplied=None):
        """Host-owned turning fractions.  Nodes flagged in tf_supplied read them (tf_mode 1); the
        others use uniform 1/(m-1) computed in the kernel (tf_mode 0); routed nodes keep tf_mode 2."""
        if tf_static is not None and len(tf_static):
            self.tf_static[: len(tf_static)].copy_(torch.from_numpy(np.ascontiguousarray(tf_static, dtype=np.float64)))
        if tf_supplied is not None:
            meta = self._meta_host
            mode = (meta[:, 1] >> 16) & 0xff
            new_mode = np.where(mode == 2, 2, np.where(np.asarray(tf_supplied, dtype=bool), 1, 0))
            meta[:, 1] = (meta[:, 1] & 0xffff) | (new_mode.astype(np.int32) << 16)
            self._net_t["nd_meta"].copy_(torch.from_numpy(meta))

    def set_demand(self, demand):
        d = torch.from_numpy(np.ascontiguousarray(demand, dtype=np.float64))
        if d.shape[1] * self.R == self.demand.shape[1] and self.R > 1:
            d = d.repeat_interleave(self.R, dim=1)
        self.demand[: d.shape[0]].copy_(d)

    # ------------------------------------------------------------------ single-network facade
    def bind_network(self, net):
        """Static visiting order of the reference's draws + host-side per-link constants."""
        self._net_ref = net
        links = list(net.links.values())
        ev_link, ev_recv = [], []
        for node in net.nodes.values():
            for l in node.incoming_links:
                if not l.is_virtual:
                    ev_link.append(l.index); ev_recv.append(0)
            for l in node.outgoing_links:
                if not l.is_virtual and not l.is_separator:
                    ev_link.append(l.index); ev_recv.append(1)
        self._ev_link = np.asarray(ev_link, dtype=np.int64)
        self._ev_recv = np.asarray(ev_recv, dtype=bool)
        self._act = np.asarray([l.activity_probability for l in links], dtype=np.float64)
        self._any_activity = bool((self._act > 0).any())
        self._sigma = np.asarray([l.speed_noise_std for l in links], dtype=np.float64)
        self._noisy = np.nonzero(self._sigma > 0)[0]
        self._demand_nodes = self.plan["demand_nodes"]
        self._od_arrays = ([net.od_manager.od_flows[k] for k in self.plan["od_keys"]]
                           if net.od_manager is not None else [])
        store = net._store
        od_w = (np.stack(self._od_arrays, axis=1) if self._od_arrays else None)
        tf, supplied = net._static_fractions()
        self.initialise(store.gate, store.sep_np64, tf, None, od_w, supplied)
        store.widths_dirty = False
        net._fractions_dirty = False

    def _push_host_edits(self, net, t):
        store = net._store
        if store.widths_dirty:
            self.set_gate(store.gate, store.sep_np64)
            store.widths_dirty = False
        if net._fractions_dirty:
            tf, supplied = net._static_fractions()
            self.set_static_fractions(tf, supplied)
            net._fractions_dirty = False
        if self._demand_nodes:
            row = np.array([float(n.demand[t - 1]) for n in self._demand_nodes], dtype=np.float64)
            self.demand[t - 1, : len(row)].copy_(torch.from_numpy(row))
        if self._od_arrays:
            row = np.array([float(a[t]) for a in self._od_arrays], dtype=np.float64)
            self.od_w[t, : len(row)].copy_(torch.from_numpy(row))

    def _push_all_demand(self, net):
        """Upload every origin's whole demand series and every OD weight series (multi-step `run` on a facade
        network: no host edits between the steps)."""
        if self._demand_nodes:
            table = np.zeros((self.S + 1, len(self._demand_nodes)))
            for k, node in enumerate(self._demand_nodes):
                d = np.asarray(node.demand, dtype=np.float64)
                table[: len(d), k] = d[: self.S + 1]
            self.set_demand(table)
        if self._od_arrays and not self._od_per_replica:
            od = np.stack([np.asarray(a, dtype=np.float64)[: self.S + 1] for a in self._od_arrays], axis=1)
            self.od_w[: od.shape[0], : od.shape[1]].copy_(torch.from_numpy(np.ascontiguousarray(od)))

    def step_network(self, net, t: int):
        if not (1 <= t <= self.S):
            raise IndexError(f"time step {t} outside [1, {self.S}]")
        with self._guard():
            self._begin_steps(t, 1)
            self._push_host_edits(net, t)
            if self.rng == "philox":
                ops.ltm_step(self.hist64, self.hist32, self.runsum, self.tf_routed, self.probs, self.err,
                             self.handle, t, 1, _native.RNG_PHILOX)
            else:
                self._numpy_draws(t)
                ops.ltm_step(self.hist64, self.hist32, self.runsum, self.tf_routed, self.probs, self.err,
                             self.handle, t, 1, _native.RNG_TABLE)
        if t != self.t_done + 1:
            net._store.invalidate()
        self.t_done = t

    def _numpy_draws(self, t):
        """Draw this step's R1..R4 outcomes from np.random in the reference's order."""
        n32 = self.L
        ops.ltm_draw_requests(self.hist64, self.hist32, self._req, self.err, self.handle, t)
        self._req_host.copy_(self._req, non_blocking=not self.emulation)
        if self._exp_arg is not None:
            self._exp_arg_host.copy_(self._exp_arg, non_blocking=not self.emulation)
        err = self.err.cpu() if self.emulation else self.err.to("cpu", non_blocking=False)
        self._raise_on_error(err)
        req = self._req_host.numpy()
        kind, n1, n3 = req[:n32], req[n32:2 * n32], req[2 * n32:3 * n32]
        rf = req[3 * n32:4 * n32].view(np.float32)
        sval = req[4 * n32:6 * n32].view(np.float64)
        out = self._draw_host.numpy()
        d1, d2, d3 = out[:n32], out[n32:2 * n32], out[2 * n32:3 * n32]
        noise = out[4 * n32:6 * n32].view(np.float64)
        ev, recv = self._ev_link, self._ev_recv
        if not self._any_activity:
            need = recv | (kind[ev] == 2)
            lk = ev[need]
            is_r = recv[need]
            n = np.where(is_r, n3[lk], n1[lk]).astype(np.int64)
            p = np.full(len(lk), 0.9)
            for i in np.nonzero(~is_r)[0]:
                # same scalar expression as the reference (link.py:317): float32 powf
                p[i] = 0.7 + (0.85 - 0.7) * rf[lk[i]] ** 0.8
            if len(lk):
                draws = np.random.binomial(n, p)
                d3[lk[is_r]] = draws[is_r]
                d1[lk[~is_r]] = draws[~is_r]
        else:
            act = self._act
            for l, is_r in zip(ev.tolist(), recv.tolist()):
                if is_r:
                    d3[l] = np.random.binomial(n=int(n3[l]), p=0.9)
                    continue
                k = kind[l]
                if k == 2:
                    flow = np.random.binomial(n=int(n1[l]), p=0.7 + (0.85 - 0.7) * rf[l] ** 0.8)
                    d1[l] = flow
                else:
                    flow = sval[l]
                if act[l] > 0 and flow > 1:
                    d2[l] = np.random.binomial(n=int(np.floor(flow)), p=act[l])
        if len(self._noisy):
            noise[self._noisy] = np.random.normal(0, self._sigma[self._noisy])
        self._draw.copy_(self._draw_host, non_blocking=not self.emulation)
        if self._exp_arg is not None:
            # the logit's exponentials with the host's numpy, as the reference evaluates them
            # (path_finder.py:585; elementwise, so one call over all groups gives the same values)
            np.exp(self._exp_arg_host.numpy(), out=self._exp_val_host.numpy())
            self._exp_val.copy_(self._exp_val_host, non_blocking=not self.emulation)

    # ------------------------------------------------------------------ batched / multi-step driving
    def run(self, t0: int, n_steps: int, rng_mode: int = _native.RNG_PHILOX):
        """Advance n_steps without host involvement (PHILOX, or TABLE after `set_draw_table`)."""
        with self._guard():
            self._begin_steps(t0, n_steps)
            ops.ltm_step(self.hist64, self.hist32, self.runsum, self.tf_routed, self.probs, self.err,
                         self.handle, t0, n_steps, rng_mode)
        self.t_done = t0 + n_steps - 1

    def run_profiled(self, t0: int, n_steps: int, rng_mode: int = _native.RNG_PHILOX):
        """`run` with per-kernel CUDA-event timing; returns (ms[4], launches[4]) for the kernels
        link_pair, route_probs, node_flows (slot 3 unused).  Synchronises."""
        ms = (C.c_double * 4)(0, 0, 0, 0)
        cnt = (C.c_int64 * 4)(0, 0, 0, 0)
        io = self._table_io if (rng_mode == _native.RNG_TABLE and self._table_io is not None) else self.io
        with self._guard():
            self._begin_steps(t0, n_steps)
            _native.check(self.lib, self.lib.pns_step_profiled(
                C.byref(self.net), C.byref(self.state), C.byref(io), t0, n_steps, rng_mode, self._stream(),
                ms, cnt), "pns_step_profiled")
            self._route_all(False)
        self.t_done = t0 + n_steps - 1
        return list(ms), list(cnt)

    def run_streamed(self, t0: int, n_steps: int, host_demand: torch.Tensor, host_metric: torch.Tensor,
                     rng_mode: int = _native.RNG_PHILOX):
        """Advance n_steps with the per-step host traffic enqueued natively: every step copies its demand
        row from `host_demand` (pinned float64, same layout as the device demand table) and copies the
        partial sums of the step's network-wide pedestrian count to `host_metric[k]` (pinned float64
        [n_steps, METRIC_ROW]; `streamed_metric(host_metric)` adds them up).
        Stream-ordered; synchronise before reading `host_metric`."""
        if not (host_demand.is_pinned() and host_metric.is_pinned()):
            raise ValueError("run_streamed needs pinned host tensors")
        if (host_demand.dtype != torch.float64 or host_metric.dtype != torch.float64
                or host_metric.numel() < n_steps * _native.METRIC_ROW or not host_metric.is_contiguous()):
            raise ValueError("host_demand / host_metric must be float64 and host_metric [n_steps, METRIC_ROW]")
        if host_demand.shape[0] < t0 + n_steps - 1 or host_demand.shape[1] != self.demand.shape[1]:
            raise ValueError("host_demand must cover rows [0, t0+n_steps-1) with the device table's width")
        if not hasattr(self, "_dev_metric") or self._dev_metric.numel() < n_steps * _native.METRIC_ROW:
            self._dev_metric = torch.zeros(n_steps * _native.METRIC_ROW, dtype=torch.float64, device=self.device)
        self._streamed_host = (host_demand, host_metric)
        with self._guard():
            self._begin_steps(t0, n_steps)
            ops.ltm_step_streamed(self.hist64, self.hist32, self.runsum, self.tf_routed, self.probs, self.err,
                                  self._dev_metric, self.handle, t0, n_steps, rng_mode)
        self.t_done = t0 + n_steps - 1

    @staticmethod
    def streamed_metric(host_metric: torch.Tensor, n_steps: int) -> np.ndarray:
        """Pedestrian count of every streamed step: the sum of its partial sums (exact, integer-valued)."""
        m = host_metric.reshape(-1)[: n_steps * _native.METRIC_ROW].numpy()
        return m.reshape(n_steps, _native.METRIC_SLOTS, _native.METRIC_STRIDE)[:, :, 0].sum(axis=1)

    def kpis(self, t_last: int, lk_role: np.ndarray, any_od_path: bool) -> torch.Tensor:
        """Episode KPIs of every replica from history rows 0..t_last (C-ABI pns_kpi; reference
        rl/rl_utils.py:770-1512): tensor [R, len(_native.KPI_NAMES)] on the engine's device.
        lk_role[l]: bit0 starts at an origin, bit1 ends at a destination, bit2 on an OD path."""
        role = torch.from_numpy(np.ascontiguousarray(lk_role, dtype=np.int32)).to(self.device)
        scratch = torch.empty((max(1, self.L * self.R) * 8,), dtype=torch.float64, device=self.device)
        out = torch.zeros((self.R, len(_native.KPI_NAMES)), dtype=torch.float64, device=self.device)
        with self._guard():
            ops.episode_kpis(self.hist64, self.hist32, self.demand, role, scratch, out, self.handle, int(t_last),
                             bool(any_od_path))
        return out

    def set_draw_table(self, draw_b: torch.Tensor, draw_n: torch.Tensor, draw_exp: torch.Tensor = None):
        """draw_b [rows, 3, L*R] int32, draw_n [rows, L*R] float64; row k serves step t0+k of `run`.
        draw_exp [rows, n_opts*R] float64 (optional): the route-choice exponentials of every step as the host
        evaluated them (without it the device evaluates them)."""
        self._table = (draw_b, draw_n, draw_exp)
        io = _native.PnsStepIO()
        C.memmove(C.byref(io), C.byref(self.io), C.sizeof(io))
        io.draw_b, io.draw_n = draw_b.data_ptr(), draw_n.data_ptr()
        io.draw_exp = draw_exp.data_ptr() if draw_exp is not None else 0
        io.draw_row_stride = 1
        self._table_io = io

    def keep_lp_solutions(self, keep=True):
        """'optimal' node model: keep the turn flows x of every step's linear programs (before the floor) in
        `self.lp_x` [n_edges * R], indexed like the turning fractions -- a test hook."""
        if keep and self.lp_x is None:
            self.lp_x = torch.zeros((max(1, self.net.n_edges) * self.R,), dtype=torch.float64, device=self.device)
        self.io.lp_x = _ptr(self.lp_x) if keep else None
        if self._table_io is not None:
            self._table_io.lp_x = self.io.lp_x

    def release(self):
        """Drop the device tensors and the reference to the network (the engine <-> network cycle would otherwise
        keep an episode's history alive until the garbage collector finds it; environments that rebuild their
        network at every reset call this on the old one, and torch's caching allocator hands the blocks to the
        next engine)."""
        net = self._net_ref
        if net is not None:
            if getattr(net, "_engine", None) is self:
                net._engine = None
            if getattr(net._store, "engine", None) is self:
                net._store.engine = None
        self._net_ref = None
        for name in ("hist64", "hist32", "gate", "sep_np64", "runsum", "tf_static", "tf_routed", "probs", "nm_s",
                     "nm_r", "demand", "od_w", "_req", "_draw", "_exp_arg", "_exp_val", "_dev_metric"):
            if hasattr(self, name):
                setattr(self, name, None)
        self._net_t = {}

    # ------------------------------------------------------------------ reading back
    def _raise_on_error(self, err_host):
        bits = int(np.bitwise_or.reduce(err_host.numpy())) if err_host.numel() else 0
        if bits:
            msgs = [m for b, m in _native.ERR_BITS.items() if bits & b]
            raise ValueError("LTM step fault on device: " + "; ".join(msgs))

    def check_errors(self):
        self._raise_on_error(self.err.cpu())

    def read_rows(self, field: str, lo: int, hi: int, out: np.ndarray):
        """Copy rows lo..hi of one history field (replica 0 layout for R=1) into `out[lo:hi+1]`."""
        if field in F64_INDEX:
            f = F64_INDEX[field]
            if f >= self.n_f64:
                return
            src = self.hist64[f, lo:hi + 1]
        else:
            src = self.hist32[F32_INDEX[field], lo:hi + 1]
        out[lo:hi + 1] = src.cpu().numpy().reshape(hi + 1 - lo, -1)[:, : out.shape[1]]
        self.check_errors()

    def routed_fractions(self, node_index: int):
        p = self.plan
        if p["nd_routed"][node_index] < 0:
            return None
        meta = p["nd_meta"][node_index]
        m = int(meta[1]) & 0xff
        a = int(meta[3])
        return self.tf_routed.view(-1, self.R)[a:a + m * (m - 1), 0].cpu().numpy()

    def history(self, field: str) -> torch.Tensor:
        """Device view [S+1, columns, R] of one field."""
        if field in F64_INDEX:
            return self.hist64[F64_INDEX[field]].view(self.S + 1, self.C64, self.R)
        return self.hist32[F32_INDEX[field]].view(self.S + 1, self.L, self.R)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()
