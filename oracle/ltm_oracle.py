"""TEST INFRASTRUCTURE ONLY -- CPU restatement of PedNStream's LTM timestep.

This file is the *oracle* the CUDA path is checked against.  It is imported only by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`; the
product package never imports it (and fails loudly without its CUDA library).

It restates, in scalar numpy/Python (so numpy-2 NEP-50 promotion rules apply exactly as they do
in the reference), the algorithm of `Network.network_loading(t)`:

    step driver            reference src/LTM/network.py:266-287  -> LtmOracle.network_loading
    sending flow           src/LTM/link.py:216-370, 199-214      -> _sending_flow, _diffusion_outflow
    receiving flow         src/LTM/link.py:372-416, 480-512      -> _receiving_flow
    node models            src/LTM/node.py:164-221, 230-242, 272-300 -> _assign_flows
    'optimal' node model   src/LTM/node.py:73-137, 249-271         -> lp_matrices, scipy_lp, _assign_flows
    logit route choice     src/LTM/path_finder.py:561-737        -> _turning_fractions
    density / speed        src/LTM/link.py:133-188, 430-452; src/utils/functions.py:112-134
                                                                   -> _update_link_states

Pinned against the live reference (same seed, bit-equal arrays) by tests/test_oracle_vs_reference.py
when /root/reference is present, and against the committed fixtures in tests/golden/ otherwise.

Random draws go through a small provider object so that the same restatement serves the three
modes of the product: `NumpyDraws` (global legacy MT19937 stream, reference order), `TableDraws`
(replay of recorded outcomes) and `PhiloxDraws` (counter-based, oracle/philox.py).
"""
from __future__ import annotations

import numpy as np

F64_FIELDS = ("inflow", "outflow", "cumulative_inflow", "cumulative_outflow",
              "sending_flow", "receiving_flow", "back_gate_width_data", "separator_width_data")
F32_FIELDS = ("num_pedestrians", "density", "speed", "travel_time", "avg_travel_time", "link_flow")


# ----------------------------------------------------------------------------- 'optimal' node model
def lp_matrices(m, turning_fractions, w=1e-2):
    """The linear program of RegularNode.solve(type='optimal') for a node with m incoming = m outgoing slots:
    cost vector, A_ub (Node.get_matrix_A, node.py:73-104) and A_eq (Node.update_matrix_A_eq, node.py:110-137).
    Variables: E = m(m-1) turn flows, then (p_e, n_e) penalty pairs interleaved."""
    E = m * m - m
    A_ub = np.zeros((2 * m, E + m + 2 * E))
    for i in range(m):                                                                 # node.py:86-89
        e = np.ones(m)
        e[i] = 0
        A_ub[i, i * m:(i + 1) * m] = e
    for j in range(m):                                                                 # node.py:92-97
        for k in range(m):
            A_ub[m + j, j + k * m] = 1 if k != j else 0
    A_ub = np.delete(A_ub, [i * m + i for i in range(m)], axis=1)                      # node.py:100-101
    A_eq = np.zeros((E, 3 * E))
    for i in range(E):                                                                 # node.py:127-135
        start = (i // (m - 1)) * (m - 1)
        A_eq[i, start:start + m - 1] = turning_fractions[i]
        A_eq[i, i] = turning_fractions[i] - 1
        A_eq[i, E + 2 * i:E + 2 * i + 2] = np.array([1, -1])
    c = np.concatenate((-1 * np.ones(E), w * np.ones(2 * E)))                          # node.py:252-254
    return c, A_ub, A_eq


def scipy_lp(m, s, r, turning_fractions, w=1e-2):
    """node.py:262: the reference's own solver call (scipy.optimize.linprog, default method and bounds).
    Returns (x, objective) or None when linprog reports failure."""
    from scipy.optimize import linprog
    c, A_ub, A_eq = lp_matrices(m, turning_fractions, w)
    res = linprog(c, A_ub=A_ub, A_eq=A_eq, b_ub=np.concatenate((s, r)), b_eq=np.zeros(m * m - m))
    return (res.x, res.fun) if res.success else None


# ----------------------------------------------------------------------------- draw providers
def numpy_release_prob(rel):
    """link.py:317 evaluated with numpy scalar semantics (float32 powf with a demoted exponent)."""
    return 0.7 + (0.85 - 0.7) * rel ** 0.8


class NumpyDraws:
    """The reference's own sampler: numpy's global legacy RandomState, consumed in visiting order."""

    release_prob = staticmethod(numpy_release_prob)
    exp = staticmethod(np.exp)

    def binomial(self, site, link, t, n, p):
        return np.random.binomial(n=n, p=p)

    def normal(self, link, t, sigma):
        return np.random.normal(0, sigma)


class TableDraws:
    """Replay of recorded outcomes: {(site, link_id, t): value}."""

    release_prob = staticmethod(numpy_release_prob)
    exp = staticmethod(np.exp)

    def __init__(self, table):
        self.table = table
        self.used = 0

    def binomial(self, site, link, t, n, p):
        self.used += 1
        return int(self.table[(site, link.link_id, t)])

    def normal(self, link, t, sigma):
        self.used += 1
        return float(self.table[("R4", link.link_id, t)])


# ----------------------------------------------------------------------------- model objects
class _OLink:
    """Parameter bag of one physical link; Python-typed values exactly as the scenario gives them."""

    def __init__(self, src, col, S):
        self.link_id = src.link_id
        self.col = col
        self.is_separator = bool(src.is_separator)
        self.length = src.length
        self._width = src._width
        self.free_flow_speed = src.free_flow_speed
        self.k_critical = src.k_critical
        self.k_jam = src.k_jam
        self.gamma = src.gamma
        self.activity_probability = src.activity_probability
        self.bi_factor = src.bi_factor
        self.fd_type = src.fd_type
        self.noise_std = src.speed_noise_std
        self.unit_time = src.unit_time
        self.capacity = self.free_flow_speed * self.k_critical
        self.shockwave_speed = self.capacity / (self.k_jam - self.k_critical)
        self.max_travel_time = self.length / 0.05
        self.front_gate_width = src.front_gate_width
        self.back_gate_width = src.back_gate_width
        self.separator_width = src.separator_width if self.is_separator else None
        self.reverse = None
        self.running_sum = None      # np.float32, link.py:84
        self.free_flow_tau = None
        self.window = None

    @property
    def area(self):
        if self.is_separator:
            return self.length * self.separator_width
        return self.length * self._width


class LtmOracle:
    """Scalar restatement of the reference step on [S+1, columns] history arrays."""

    def __init__(self, net, draws=None, lp_solver=None):
        """net: a host-side `pednstream_b200.Network` (topology, parameters, demand, route
        structures -- all built on the host and verified against the reference's own setup).
        lp_solver(node, t, m, s, r, tf) -> x or None replaces the scipy call of the 'optimal' node model
        (tests replay the device's optimal vertices through it: where a program's optimum is a face, two
        solvers legitimately return different x)."""
        self.net = net
        self.draws = draws or NumpyDraws()
        self.node_model = getattr(net, "assign_flows_type", "classic")
        self.lp_solver = lp_solver
        self.last_q = {}
        self.S = S = net.simulation_steps
        self.unit_time = net.unit_time
        links = list(net.links.values())
        self.L = L = len(links)
        self.V = len(net._virtual_cols)
        C = L + self.V
        self.h = {f: np.zeros((S + 1, C)) for f in F64_FIELDS}
        self.h["sending_flow"][:] = -1.0
        self.h["receiving_flow"][:] = -1.0
        for f in F32_FIELDS:
            self.h[f] = np.zeros((S + 1, L), dtype=np.float32)
        self.links = []
        for l in links:
            o = _OLink(l, l.index, S)
            self.links.append(o)
        for o in self.links:
            o.reverse = self.links[o.col ^ 1]
            tt = self.h["travel_time"]
            tt[0, o.col] = min(o.length / o.free_flow_speed, o.max_travel_time)     # link.py:83
            o.running_sum = tt[0, o.col]                                              # np.float32
            o.free_flow_tau = round(tt[0, o.col] / o.unit_time)                       # link.py:86
            o.window = round(100 / o.unit_time)                                       # link.py:89
            self.h["avg_travel_time"][: o.window, o.col] = tt[0, o.col]               # link.py:91
            self.h["back_gate_width_data"][:, o.col] = o._width                       # link.py:56
            if o.is_separator:
                self.h["separator_width_data"][:, o.col] = o._width / 2               # link.py:425
        self.nodes = list(net.nodes.values())
        self.tf = {}           # node_id -> current turning fractions
        self.turn_probs = {}   # node_id -> {od: {turn: p}}
        self.errors = []
        self._probs_stamp = {}

    # convenience -----------------------------------------------------------------------------
    def link_by_key(self, key):
        return self.links[self.net.links[key].index]

    def set_back_gate_width(self, key, value):
        """link.back_gate_width setter + coupling to the reverse front gate (link.py:121-126)."""
        o = self.link_by_key(key)
        o.back_gate_width = value
        o.reverse.front_gate_width = value

    def set_separator_width(self, key, value):
        """Separator.separator_width setter (link.py:462-478)."""
        o = self.link_by_key(key)
        o.separator_width = value
        o.front_gate_width = value
        o.back_gate_width = value
        r = o.reverse
        r.separator_width = o._width - value
        r.front_gate_width = o._width - value
        r.back_gate_width = o._width - value

    # ------------------------------------------------------------------------------ densities
    def _shared_density(self, o, t):
        """Link.get_density (link.py:190-197) / Separator.get_density (:427-428)."""
        num = self.h["num_pedestrians"]
        if o.is_separator:
            return self.h["density"][t, o.col]
        return (num[t, o.col] + num[t, o.reverse.col]) / o.area

    # ------------------------------------------------------------------------------ sending flow
    def _diffusion_outflow(self, o, t, tau):
        """Link.get_outflow (link.py:199-214): 4-tap geometric smoothing of lagged inflow.
        Negative indices wrap Python-style, exactly as numpy indexing does in the reference."""
        inflow = self.h["inflow"][:, o.col]
        travel_time = self.h["avg_travel_time"][t, o.col]
        F = 1 / (1 + o.gamma * travel_time)
        total = (F * inflow[t - tau] + F * (1 - F) * inflow[t - tau - 1] +
                 F * (1 - F) ** 2 * inflow[t - tau - 2] +
                 F * (1 - F) ** 3 * inflow[t - tau - 3])
        return max(np.ceil(total), 0)

    def _sending_flow(self, o, t):
        """Link.cal_sending_flow(t) with t = time_step-1 (link.py:216-370)."""
        h = self.h
        c = o.col
        density = self._shared_density(o, t)
        tau = round(h["avg_travel_time"][t, c] / o.unit_time)                         # :260
        if t < o.free_flow_tau:                                                        # :267-269
            h["sending_flow"][t, c] = 0
            return h["sending_flow"][t, c]
        idx = max(0, t + 1 - tau)                                                      # :274
        cong = np.clip((h["density"][t, c] - o.k_critical) / (o.k_jam - o.k_critical), 0, 1)   # :282
        stock = h["num_pedestrians"][t, c]
        arrived = max(0, h["cumulative_inflow"][idx, c] - h["cumulative_outflow"][t, c])
        boundary = cong * stock + (1 - cong) * arrived                                 # :287-288
        gate_cap = o.front_gate_width * o.k_critical * o.free_flow_speed * o.unit_time # :296
        flow = min(boundary, gate_cap)
        original = flow
        if flow > 0:
            rel = np.clip(density / o.k_jam, 0, 1)                                     # :315
            p_release = self.draws.release_prob(rel)                                   # :317
            if density <= o.k_critical:
                spread = self._diffusion_outflow(o, t, tau)
                if spread > 0:
                    w = 0.8
                    flow = int(np.floor(min(w * spread + (1 - w) * flow, flow)))        # :330
                else:
                    flow = self.draws.binomial("R1", o, t, int(np.floor(flow)), p_release)   # :336-338
            else:
                flow = self.draws.binomial("R1", o, t, int(np.floor(flow)), p_release)       # :342-344
            if flow < 0:
                raise ValueError(f"Negative sending flow detected, {flow} at {t}")
        if o.activity_probability > 0 and flow > 1:                                    # :351-358
            flow -= self.draws.binomial("R2", o, t, int(np.floor(flow)), o.activity_probability)
        flow = max(0, flow)
        flow = min(np.floor(0.8 * flow + 0.2 * h["sending_flow"][t - 1, c]), original) # :363-364
        if flow < 0:
            raise ValueError("Negative sending flow detected, sending flow more than original flow")
        h["sending_flow"][t, c] = flow
        return h["sending_flow"][t, c]

    # ------------------------------------------------------------------------------ receiving flow
    def _receiving_flow(self, o, t, reverse_sending):
        """cal_receiving_flow_with_reverse (link.py:372-416); Separator variant (:480-512)."""
        h = self.h
        c = o.col
        tau_sw = round(o.length / (o.shockwave_speed * o.unit_time))                   # :380
        space = o.k_jam * o.area
        if o.is_separator:
            if t + 1 - tau_sw < 0:
                bound = space
            else:
                bound = h["cumulative_outflow"][t + 1 - tau_sw, c] + space - h["cumulative_inflow"][t, c]
        else:
            opposing = h["num_pedestrians"][t, o.reverse.col]
            blockers = self.draws.binomial("R3", o, t, opposing, 0.9)                  # :381-382
            if t + 1 - tau_sw < 0:
                bound = space - blockers
            else:
                bound = max(0, h["cumulative_outflow"][t + 1 - tau_sw, c] + space
                            - blockers - h["cumulative_inflow"][t, c])                 # :389-390
        gate_cap = o.back_gate_width * o.k_critical * o.free_flow_speed * o.unit_time  # :393
        flow = min(bound, gate_cap)
        flow = max(flow, 0)
        prev = h["receiving_flow"][t - 1, c]
        if prev >= 0:                                                                  # :400-401
            flow = min(np.floor(flow * 0.8 + prev * 0.2), flow)
        if o.is_separator:
            return max(flow, 0)
        return max(flow - reverse_sending, 0)                                          # :415-416

    # ------------------------------------------------------------------------------ route choice
    def _group_probs(self, node, od, t, pf):
        """PathFinder.update_node_turn_probs (path_finder.py:561-589) for one OD at one node."""
        store = self.turn_probs.setdefault(node.node_id, {}).setdefault(od, {})
        if self._probs_stamp.get((node.node_id, od)) == t:      # idempotent within a step
            return store
        self._probs_stamp[(node.node_id, od)] = t
        for up, downs in node.turns_distances[od].items():
            if not downs:
                continue
            turns = [(up, d) for d in downs]
            dist = list(downs.values())
            dens, caps = [], []
            for d in downs:
                key = (node.node_id, d)
                if key in self.net.links:
                    o = self.link_by_key(key)
                    dens.append(self._shared_density(o, t - 1))
                    cap = self.h["receiving_flow"][t - 2, o.col]
                    caps.append(cap if cap >= 0 else
                                o.back_gate_width * o.free_flow_speed * o.k_critical * o.unit_time)
                else:                       # virtual destination link
                    dens.append(0)
                    caps.append(100)
            crowd = np.maximum(np.array(dens) - 2, 0) / (10 - 2)
            util = (pf.alpha * np.array(dist) / (np.sum(dist) + 1e-6)
                    + pf.beta * crowd
                    - pf.omega * np.array(caps) / (np.sum(caps) + 1e-6)) + pf.epsilon
            e = self.draws.exp(-pf.temp * util)
            store.update(dict(zip(turns, e / np.sum(e))))
        return store

    def _turning_fractions(self, node, t):
        """update_turning_fractions + check_fractions (path_finder.py:591-715)."""
        pf, odm = self.net.path_finder, self.net.od_manager
        tf = np.zeros(node.edge_num)
        up_od = {}
        for up, ods in node.up_od_probs.items():
            w = {od: odm.get_od_flow(od[0], od[1], t) for od in ods}
            total = 0
            for od in ods:
                total += w[od]
            if total > 0:
                w = {od: w[od] / total for od in ods}
            else:
                w = {od: (1.0 / len(ods) if len(ods) > 0 else 0) for od in ods}
            up_od[up] = w
        ups = [l.start_node.node_id if l.start_node is not None else -1 for l in node.incoming_links]
        downs = [l.end_node.node_id if l.end_node is not None else -1 for l in node.outgoing_links]
        k = 0
        for up in ups:
            for down in downs:
                if up == down:
                    continue
                acc = 0
                for od in node.ods_in_turns.get((up, down), set()):
                    probs = self._group_probs(node, od, t, pf)
                    acc += probs.get((up, down), 0) * up_od.get(up, {}).get(od, 0)
                tf[k] = acc
                k += 1
        rows = tf.reshape(node.dest_num, node.source_num - 1)
        for i in range(node.dest_num):
            s = np.sum(rows[i])
            if np.abs(s - 1) > 1e-3:
                if s > 1e-6:
                    rows[i] = rows[i] / s
                else:
                    rows[i] = np.ones(node.source_num - 1) / (node.source_num - 1)
        return rows.flatten()

    # ------------------------------------------------------------------------------ node phase
    def _assign_flows(self, node, t):
        """Node.assign_flows + solve + update_links (node.py:146-300)."""
        h = self.h
        m = node.source_num
        s = np.zeros(m)
        r = np.zeros(m)
        for i, l in enumerate(node.incoming_links):
            if l.is_virtual:
                s[i] = node.demand[t - 1]
            else:
                s[i] = self._sending_flow(self.links[l.index], t - 1)
        for j, l in enumerate(node.outgoing_links):
            if l.is_virtual:
                r[j] = node.M
            else:
                o = self.links[l.index]
                rev_send = h["sending_flow"][t - 1, o.reverse.col].copy()
                if rev_send < 0:
                    raise Warning(f"Negative reverse sending flow detected at time step {t}: {rev_send}")
                r[j] = self._receiving_flow(o, t - 1, rev_send)
                h["receiving_flow"][t - 1, o.col] = r[j]
        if np.any(s < 0) or np.any(r < 0):
            raise Warning(f"Negative flows detected at time step {t}: s={s}, r={r}")

        if node.kind == 0:                                                             # node.py:230-242
            a, b = np.min([s[0], r[1]]), np.min([s[1], r[0]])
            q = np.array([a, b, b, a])
            if np.any(q < 0):
                raise Warning(f"Negative flows detected: {q}")
        elif self.node_model == "optimal":                                             # node.py:249-271
            tf = self.tf[node.node_id]
            if self.lp_solver is not None:
                x = self.lp_solver(node, t, m, s, r, tf)
            else:
                sol = scipy_lp(m, s, r, tf, getattr(node, "w", 1e-2))
                x = None if sol is None else sol[0]
            if x is None:                      # node.py:265: `if res.success` -- q keeps its previous value
                q = self.last_q[node.node_id]
            else:
                _, A_ub, _ = lp_matrices(m, tf)
                xf = np.zeros(A_ub.shape[1])
                xf[:len(x)] = np.floor(x)
                q = np.maximum(0, A_ub @ xf)                                           # node.py:266-268
            self.last_q[node.node_id] = q
        else:                                                                          # node.py:272-300
            tf = self.tf[node.node_id]
            P = np.zeros((m, m))
            P[~np.eye(m, dtype=bool)] = tf
            W = P * np.tile(s, (m, 1)).T
            col = np.sum(W, axis=0, keepdims=True)
            share = W / np.where(col != 0, col, 1e-5)
            supply = r * share
            g = np.zeros((m, m))
            e = 0
            for i in range(m):
                for j in range(m):
                    if i == j:
                        continue
                    g[i, j] = min(tf[e] * s[i], supply[i][j])
                    e += 1
            f = np.floor(g)
            q = np.maximum(0, np.concatenate([f.sum(axis=1), f.sum(axis=0)]))
        for i, l in enumerate(node.incoming_links):                                    # node.py:154-161
            h["outflow"][t, l._col] = q[i]
            h["cumulative_outflow"][t, l._col] = h["cumulative_outflow"][t - 1, l._col] + q[i]
        for j, l in enumerate(node.outgoing_links):
            h["inflow"][t, l._col] = q[m + j]
            h["cumulative_inflow"][t, l._col] = h["cumulative_inflow"][t - 1, l._col] + q[m + j]

    # ------------------------------------------------------------------------------ link phase
    def _speed(self, o, k_self, k_opp, t):
        """BiDirectionalFd.__call__ (functions.py:112-134)."""
        k_eff = k_self + o.bi_factor * k_opp
        vf, kc, kj = o.free_flow_speed, o.k_critical, o.k_jam
        if o.fd_type == "greenshields":
            v = vf if k_eff <= kc else max(0, -vf * (k_eff - kj) / (kj - kc))
        elif o.fd_type == "yperman":
            v = vf if k_eff <= kc else max(0, (kc * vf) / (kj - kc) * (kj / k_eff - 1))
        elif o.fd_type == "smulders":
            v = vf * (1 - k_eff / kj) if k_eff <= kc else max(0, (vf * kc) * (1 / k_eff - 1 / kj))
        else:
            raise ValueError(f"Unknown model type: {o.fd_type}")
        if o.noise_std > 0:
            v += self.draws.normal(o, t, o.noise_std)
        return max(0, v)

    def _update_link_states(self, t):
        """Network.update_link_states (network.py:257-264): all densities, then all speeds."""
        h = self.h
        for o in self.links:                                                           # link.py:133-136
            c = o.col
            delta = h["inflow"][t, c] - h["outflow"][t, c]
            h["num_pedestrians"][t, c] = h["num_pedestrians"][t - 1, c] + delta
            h["density"][t, c] = h["num_pedestrians"][t, c] / o.area
        for o in self.links:                                                           # link.py:141-188
            c = o.col
            k_self = h["density"][t, c]
            k_opp = 0 if o.is_separator else h["density"][t, o.reverse.col]
            v = self._speed(o, k_self, k_opp, t)
            h["speed"][t, c] = v
            h["travel_time"][t, c] = o.length / v if v > 0 else o.max_travel_time
            h["link_flow"][t, c] = h["speed"][t, c] * h["density"][t, c]              # functions.py:97-101
            o.running_sum += h["travel_time"][t, c]
            if t >= o.window:
                o.running_sum -= h["travel_time"][t - o.window, c]
                h["avg_travel_time"][t, c] = o.running_sum / o.window
            if o.is_separator:                                                         # link.py:451-452
                h["separator_width_data"][t, c] = o.separator_width
                h["back_gate_width_data"][t, c] = o.separator_width
            else:
                h["back_gate_width_data"][t, c] = o.back_gate_width

    # ------------------------------------------------------------------------------ step driver
    def network_loading(self, t):
        """Network.network_loading(t) (network.py:266-287): nodes in dict order, then links."""
        pf = self.net.path_finder
        for node in self.nodes:
            if node.node_id not in self.tf:
                static = node._tf_static
                if static is None:
                    static = np.ones(node.edge_num) * (1 / (node.dest_num - 1))
                self.tf[node.node_id] = np.array(static, dtype=np.float64)
            if self.net.destination_nodes and node.node_id in pf.nodes_in_paths and node.source_num > 2:
                self.tf[node.node_id] = self._turning_fractions(node, t)
            self._assign_flows(node, t)
        self._update_link_states(t)

    def run(self, t_end, t_start=1):
        for t in range(t_start, t_end + 1):
            self.network_loading(t)
        return self.h
