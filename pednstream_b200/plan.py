"""Compile a host-side network description into the flat arrays the CUDA kernels read.

The *plan* is immutable topology + parameters:

node table   (network.nodes order)   kind, CSR slot range, per-slot in/out history columns,
                                      demand row, offset of its turning fractions
link table   (network.links order)   fp64 params + integer lags; the reverse of link l is l^1
                                      (links are created in (i,j),(j,i) pairs)
route plan   (routed nodes only)     see PathFinder.export_route_plan

Columns: physical link l owns column l of every history field; the virtual in/out links of the
k-th node that has them own fp64 columns L+2k and L+2k+1.
"""
from __future__ import annotations

import numpy as np

from .link import FD_TYPES

MAX_DEGREE = 8      # slots per node handled by the node kernel (shipped data: <= 6)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


CLASS_DTYPE = np.dtype([
    ("length", "f8"), ("area", "f8"), ("space", "f8"), ("kc", "f8"), ("vf", "f8"), ("kj", "f8"),
    ("act", "f8"), ("sigma", "f8"),
    ("kc32", "f4"), ("kj32", "f4"), ("kj_minus_kc32", "f4"), ("gamma32", "f4"), ("bi32", "f4"),
    ("area32", "f4"), ("length32", "f4"), ("yp_coef32", "f4"), ("neg_vf32", "f4"), ("vf32", "f4"),
    ("sm_gamma32", "f4"), ("inv_kj32", "f4"), ("max_tt32", "f4"), ("tt0", "f4"),
    ("fftau", "i4"), ("swtau", "i4"), ("flags", "i4"), ("pad_", "i4")], align=True)   # = pns_link_class


def class_record(length, width, vf, kc, kj, gamma, act, bi, sigma, fd_type, is_sep, unit_time):
    """One pns_link_class row.  Every derived constant is evaluated here with the reference's own
    Python expression (operand types and order as in src/LTM/link.py / src/utils/functions.py) and
    stored in the precision numpy would use it in next to a float32 history value."""
    f32 = np.float32
    rec = np.zeros((), dtype=CLASS_DTYPE)
    area = length * width                                   # link.py:131
    tt0 = f32(min(length / vf, length / 0.05))              # link.py:63,83
    shock = (vf * kc) / (kj - kc)                           # link.py:58,61
    rec["length"], rec["area"], rec["space"] = length, area, kj * area          # link.py:386
    rec["kc"], rec["vf"], rec["kj"], rec["act"], rec["sigma"] = kc, vf, kj, act, sigma
    rec["kc32"], rec["kj32"], rec["kj_minus_kc32"] = f32(kc), f32(kj), f32(kj - kc)
    rec["gamma32"], rec["bi32"], rec["area32"], rec["length32"] = f32(gamma), f32(bi), f32(area), f32(length)
    rec["yp_coef32"] = f32((kc * vf) / (kj - kc))           # functions.py:123
    rec["neg_vf32"] = f32(-vf)                              # functions.py:118
    rec["vf32"] = f32(vf)                                   # functions.py:126
    rec["sm_gamma32"] = f32(vf * kc)                        # functions.py:108,128
    rec["inv_kj32"] = f32(1 / kj)                           # functions.py:128
    rec["max_tt32"] = f32(length / 0.05)                    # link.py:63,177
    rec["tt0"] = tt0
    rec["fftau"] = round(tt0 / unit_time)                   # link.py:86
    rec["swtau"] = round(length / (shock * unit_time))      # link.py:380
    if int(rec["swtau"]) == 0:
        # a zero shock-wave lag makes cal_receiving_flow read cumulative_outflow[t], which the reference fills in
        # node-visiting order during the same step: the result depends on that order and has no parallel
        # counterpart (the travel-time lag has the same hazard; the kernels flag it as PNS_ERR_ZERO_LAG)
        raise ValueError(f"link of length {length} m: shock-wave lag round(length / (w * unit_time)) is 0 with "
                         f"w = {shock:.4g} m/s and unit_time = {unit_time}; use a smaller unit_time")
    rec["flags"] = (1 if is_sep else 0) | (FD_TYPES[fd_type] << 1)
    return rec


def link_classes(links, unit_time):
    """Deduplicate link parameter sets: returns (classes[CLASS_DTYPE], lk_class[int32])."""
    seen, rows, idx = {}, [], []
    for l in links:
        key = (l.length, l._width, l.free_flow_speed, l.k_critical, l.k_jam, l.gamma,
               l.activity_probability, l.bi_factor, l.speed_noise_std, l.fd_type, bool(l.is_separator))
        k = seen.get(key)
        if k is None:
            k = seen[key] = len(rows)
            rec = class_record(*key, unit_time)
            assert int(rec["fftau"]) == l.free_flow_tau and int(rec["swtau"]) == l.shockwave_tau
            rows.append(rec)
        idx.append(k)
    return np.array(rows, dtype=CLASS_DTYPE).reshape(len(rows)), _i32(idx)


def node_stride(max_degree: int) -> int:
    """Slots reserved per node in the node-major exchange arrays (4 or 8)."""
    return 4 if max_degree <= 4 else 8


def link_slots(nd_meta, nd_in_col, n_links, stride):
    """[n_links, 2] int32 exchange slots of every physical link (see `pns_net.lk_slots`):
    column 0: node-major slot that receives the link's sending flow / returns its outflow
              (END node * stride + position of the link in that node's incoming list);
    column 1: slot that receives its receiving flow / returns its inflow
              (START node * stride + position in the outgoing list = slot of its reverse)."""
    nd_meta = np.asarray(nd_meta)
    m = nd_meta[:, 1] & 0xff
    node_of_slot = np.repeat(np.arange(len(nd_meta)), m)
    slot = np.arange(len(nd_in_col)) - np.repeat(nd_meta[:, 0], m)
    cols = np.asarray(nd_in_col)
    phys = cols < n_links
    out = np.full((n_links, 2), -1, dtype=np.int32)
    where = (node_of_slot * stride + slot).astype(np.int32)
    out[cols[phys], 0] = where[phys]
    out[cols[phys] ^ 1, 1] = where[phys]
    assert (out >= 0).all(), "every link must end and start at a node slot"
    return out


def slot_links(nd_meta, nd_in_col, stride):
    """[n_nodes * stride] int32: history column of the incoming link of every node-major slot
    (`pns_net.nd_in_link`), -1 for unused slots; the slot's outgoing link is that column ^ 1."""
    nd_meta = np.asarray(nd_meta)
    m = nd_meta[:, 1] & 0xff
    node_of_slot = np.repeat(np.arange(len(nd_meta)), m)
    slot = np.arange(len(nd_in_col)) - np.repeat(nd_meta[:, 0], m)
    out = np.full((len(nd_meta) * stride,), -1, dtype=np.int32)
    out[node_of_slot * stride + slot] = np.asarray(nd_in_col)
    return out


def lane_block_order(nd_meta, nd_in_link, n_links, stride, block, hops=4):
    """Launch order of the single-replica link kernel (`pns_net.lane_order`): a permutation of the
    blocks of `block` consecutive links that puts first the blocks holding a link within `hops` nodes
    of a demand origin (queues form there, and the blockers draw of a jammed link is the longest
    serial piece of work in the step), index order otherwise.  Returns None when nothing moves."""
    nd_meta = np.asarray(nd_meta)
    n_nodes = len(nd_meta)
    slots = np.asarray(nd_in_link).reshape(n_nodes, stride)
    head = np.full(n_links, -1, dtype=np.int64)                 # end node of every physical link
    nn, kk = np.nonzero((slots >= 0) & (slots < n_links))
    head[slots[nn, kk]] = nn
    tail = head[np.arange(n_links) ^ 1]                         # start node = end node of the reverse link
    dist = np.full(n_nodes, hops + 1, dtype=np.int64)
    frontier = np.nonzero(nd_meta[:, 2] >= 0)[0]                # nodes with a demand row
    if frontier.size == 0:
        return None
    dist[frontier] = 0
    for d in range(1, hops + 1):
        reach = np.zeros(n_nodes, dtype=bool)
        reach[frontier] = True
        nxt = np.unique(head[reach[tail]])                      # nodes one link downstream of the frontier
        nxt = nxt[dist[nxt] > d]
        if nxt.size == 0:
            break
        dist[nxt] = d
        frontier = nxt
    near = (np.minimum(dist[head], dist[tail]) <= hops)
    n_blocks = (n_links + block - 1) // block
    pad = np.zeros(n_blocks * block, dtype=bool)
    pad[:n_links] = near
    first = pad.reshape(n_blocks, block).any(axis=1)
    if not first.any() or first.all():
        return None
    return np.concatenate([np.nonzero(first)[0], np.nonzero(~first)[0]]).astype(np.int32)


def route_row_tables(p):
    """Per-row views of the route plan for the route kernel (one thread per upstream slot of a routed node):
    rt_row_routed[row] the routed node of the row, rt_row_grp_ptr / rt_row_grp the (od, upstream) groups
    registered at it, rt_term_od the OD column of every accumulation term."""
    routed_nodes, row0 = np.asarray(p["rt_routed_nodes"]), np.asarray(p["rt_routed_row0"])
    n_rows = len(p["rt_row_ptr"]) - 1
    row_routed = np.zeros(n_rows, dtype=np.int32)
    bounds = list(row0) + [n_rows]
    for i in range(len(routed_nodes)):
        row_routed[bounds[i]:bounds[i + 1]] = i
    pos = {int(n): i for i, n in enumerate(routed_nodes)}
    grp_row = np.array([row0[pos[int(n)]] + int(u) for n, u in zip(p["rt_grp_node"], p["rt_grp_up"])], dtype=np.int64)
    order = np.argsort(grp_row, kind="stable").astype(np.int32)
    ptr = np.zeros(n_rows + 1, dtype=np.int32)
    if len(grp_row):
        ptr[1:] = np.cumsum(np.bincount(grp_row, minlength=n_rows))
    term_od = np.asarray(p["rt_row_od"], dtype=np.int32)[np.asarray(p["rt_term_row_entry"], dtype=np.int64)] \
        if len(p["rt_term_row_entry"]) else np.zeros(0, dtype=np.int32)
    # Rows whose fractions can change from step to step: a group with several options (the logit depends on
    # densities and receiving flows) or several registered ODs (P(od | up) depends on the step's OD weights).
    # Every other row -- nothing registered, or single-option groups of a single OD -- evaluates to the same
    # constants on every step (P(down | up, od) = exp(x)/exp(x) = 1, P(od | up) = w/w = 1): the route kernel
    # computes those once, on the first step after the state is initialised.
    n_opt = np.diff(np.asarray(p["rt_opt_ptr"], dtype=np.int64)) if len(p["rt_opt_ptr"]) > 1 else np.zeros(0, dtype=np.int64)
    row_ods = np.diff(np.asarray(p["rt_row_ptr"], dtype=np.int64))
    dynamic = row_ods > 1
    for g, row in enumerate(grp_row):
        if n_opt[g] > 1:
            dynamic[row] = True
    return dict(rt_row_routed=row_routed, rt_row_grp_ptr=ptr, rt_row_grp=order, rt_term_od=_i32(term_od),
                rt_dyn_rows=_i32(np.nonzero(dynamic)[0]))


def attach_empty_route_plan(p):
    for k in ("routed_nodes", "routed_edge0", "routed_row0", "grp_node", "grp_up", "grp_od",
              "grp_has_virtual", "opt_link", "opt_slot", "row_od", "term_opt", "term_row_entry"):
        p["rt_" + k] = _i32([])
    for k in ("opt_ptr", "row_ptr", "term_ptr"):
        p["rt_" + k] = _i32([0])
    p["rt_opt_dist"] = _f64([])
    p["rt_scalars"] = _f64([0.1, 1.0, 0.05, 0.05, 0.0])
    p["n_od"] = 0
    p["od_keys"] = []
    p.update(route_row_tables(p))


def attach_route_plan(p, path_finder, od_manager, nodes_in_order, link_index):
    """Adds the route-choice tables of `path_finder` to plan `p` (and marks the routed nodes in nd_meta).
    nodes_in_order: Node-like objects in plan node order with `.index` = plan node index (only nodes on OD paths
    need to be present); link_index: mapping (u, v) -> physical link index."""
    od_keys = list(od_manager.od_flows.keys())
    od_index = {k: i for i, k in enumerate(od_keys)}
    rp = path_finder.export_route_plan(nodes_in_order, od_index, link_index)
    p.update({"rt_" + k: v for k, v in rp.items()})
    p["rt_scalars"] = _f64([path_finder.temp, path_finder.alpha, path_finder.beta,
                            path_finder.omega, path_finder.epsilon])
    p["n_od"] = len(od_keys)
    p["od_keys"] = od_keys
    p.update(route_row_tables(p))
    meta = p["nd_meta"]
    routed = np.asarray(p["rt_routed_nodes"], dtype=np.int64)
    meta[routed, 1] = (meta[routed, 1] & 0xffff) | (2 << 16)


def compile_plan(nodes, links, unit_time, simulation_steps, path_finder=None, od_manager=None, node_model="classic"):
    """nodes: list of Node (network.nodes order, .index set); links: list of Link (network.links
    order, .index set).  node_model: Network.assign_flows_type -- with 'optimal' every regular node solves the
    linear program of reference node.py:249-271 (kind 2).  Returns dict name -> numpy array / python scalar."""
    L = len(links)
    for l in links:
        assert links[l.index ^ 1] is l.reverse_link, "links must come in forward/reverse pairs"

    p = {}
    p["n_links"] = L
    p["n_nodes"] = len(nodes)
    p["sim_steps"] = int(simulation_steps)
    p["unit_time"] = float(unit_time)
    p["window"] = int(round(100 / unit_time))

    # ---- link classes + per-link class index ------------------------------------------------
    classes, lk_class = link_classes(links, unit_time)
    p["classes"] = classes
    p["lk_class"] = lk_class
    p["lk_width"] = _f64([l._width for l in links])
    p["has_separators"] = any(l.is_separator for l in links)

    # ---- node table -------------------------------------------------------------------
    slot0, in_col, meta, tf_ptr = 0, [], [], 0
    n_virtual_nodes = 0
    demand_nodes = []
    routed_ids = set()
    if path_finder is not None:
        routed_ids = {n.node_id for n in nodes
                      if n.node_id in path_finder.nodes_in_paths and n.source_num > 2}
    for n in nodes:
        m = n.source_num
        if m != n.dest_num:
            raise ValueError(f"node {n.node_id}: {m} incoming vs {n.dest_num} outgoing links")
        if m > MAX_DEGREE:
            raise ValueError(f"node {n.node_id} has {m} link slots; the node kernel handles <= {MAX_DEGREE}")
        if n.kind == 0 and m != 2:
            raise ValueError(f"one-to-one node {n.node_id} must have exactly 2 slots, has {m}")
        for lin, lout in zip(n.incoming_links, n.outgoing_links):
            assert lout._col == lin._col ^ 1, "outgoing link of a slot must be the reverse of its incoming link"
            in_col.append(lin._col)
        dem_row = -1
        if n.virtual_incoming_link is not None:
            assert n.incoming_links[0] is n.virtual_incoming_link
            assert n.outgoing_links[0] is n.virtual_outgoing_link
            dem_row = len(demand_nodes)
            demand_nodes.append(n)
            n_virtual_nodes += 1
        tf_mode = 2 if n.node_id in routed_ids else 0
        if dem_row >= 0:
            assert n.incoming_links[0]._col == L + 2 * dem_row, "virtual columns follow the demand rows"
        kind = 2 if (node_model == "optimal" and n.kind == 1 and m >= 2) else n.kind
        meta.append((slot0, m | (kind << 8) | (tf_mode << 16), dem_row, tf_ptr))
        slot0 += m
        tf_ptr += m * (m - 1)
    p["nd_meta"] = _i32(meta).reshape(-1, 4)
    p["nd_in_col"] = _i32(in_col)                       # host-side only (used to derive lk_slots)
    p["max_degree"] = int(max((n.source_num for n in nodes), default=0))
    p["nd_stride"] = node_stride(p["max_degree"])
    p["lk_slots"] = link_slots(p["nd_meta"], p["nd_in_col"], L, p["nd_stride"])
    p["nd_in_link"] = slot_links(p["nd_meta"], p["nd_in_col"], p["nd_stride"])
    p["n_virtual"] = 2 * n_virtual_nodes
    p["n_demand_rows"] = len(demand_nodes)
    p["demand_nodes"] = demand_nodes
    p["n_edges"] = tf_ptr
    p["lp_w"] = float(next((n.w for n in nodes if hasattr(n, "w")), 0.01))     # Node.w (node.py:14)

    # ---- route plan -------------------------------------------------------------------
    if path_finder is not None:
        link_index = {(l.start_node.node_id, l.end_node.node_id): l.index for l in links}
        attach_route_plan(p, path_finder, od_manager, nodes, link_index)
    else:
        attach_empty_route_plan(p)
    # per-node routed index (-1 = static fractions)
    routed_of = np.full(len(nodes), -1, dtype=np.int32)
    routed_of[p["rt_routed_nodes"]] = np.arange(len(p["rt_routed_nodes"]), dtype=np.int32)
    p["nd_routed"] = routed_of
    assert set(np.nonzero(routed_of >= 0)[0].tolist()) == {n.index for n in nodes if n.node_id in routed_ids}
    return p
