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


def compile_plan(nodes, links, unit_time, simulation_steps, path_finder=None, od_manager=None):
    """nodes: list of Node (network.nodes order, .index set); links: list of Link (network.links
    order, .index set).  Returns dict name -> numpy array / python scalar."""
    L = len(links)
    for l in links:
        assert links[l.index ^ 1] is l.reverse_link, "links must come in forward/reverse pairs"

    p = {}
    p["n_links"] = L
    p["n_nodes"] = len(nodes)
    p["sim_steps"] = int(simulation_steps)
    p["unit_time"] = float(unit_time)
    p["window"] = int(round(100 / unit_time))

    # ---- link table -------------------------------------------------------------------
    p["lk_length"] = _f64([l.length for l in links])
    p["lk_width"] = _f64([l._width for l in links])
    p["lk_vf"] = _f64([l.free_flow_speed for l in links])
    p["lk_kc"] = _f64([l.k_critical for l in links])
    p["lk_kj"] = _f64([l.k_jam for l in links])
    p["lk_gamma"] = _f64([l.gamma for l in links])
    p["lk_act"] = _f64([l.activity_probability for l in links])
    p["lk_bi"] = _f64([l.bi_factor for l in links])
    p["lk_sigma"] = _f64([l.speed_noise_std for l in links])
    p["lk_tt0"] = np.asarray([l.travel_time0 for l in links], dtype=np.float32)
    p["lk_fftau"] = _i32([l.free_flow_tau for l in links])
    p["lk_swtau"] = _i32([l.shockwave_tau for l in links])
    # bit0 separator, bits 1-2 fundamental diagram
    p["lk_flags"] = _i32([(1 if l.is_separator else 0) | (FD_TYPES[l.fd_type] << 1) for l in links])
    p["has_separators"] = any(l.is_separator for l in links)

    # ---- node table -------------------------------------------------------------------
    node_ptr, in_col, out_col, kind, dem_row, tf_ptr = [0], [], [], [], [], [0]
    n_virtual_nodes = 0
    demand_nodes = []
    for n in nodes:
        m = n.source_num
        if m != n.dest_num:
            raise ValueError(f"node {n.node_id}: {m} incoming vs {n.dest_num} outgoing links")
        if m > MAX_DEGREE:
            raise ValueError(f"node {n.node_id} has {m} link slots; the node kernel handles <= {MAX_DEGREE}")
        if n.kind == 0 and m != 2:
            raise ValueError(f"one-to-one node {n.node_id} must have exactly 2 slots, has {m}")
        has_virtual = n.virtual_incoming_link is not None
        for lin, lout in zip(n.incoming_links, n.outgoing_links):
            in_col.append(lin._col)
            out_col.append(lout._col)
        if has_virtual:
            assert n.incoming_links[0] is n.virtual_incoming_link
            assert n.outgoing_links[0] is n.virtual_outgoing_link
            dem_row.append(len(demand_nodes))
            demand_nodes.append(n)
            n_virtual_nodes += 1
        else:
            dem_row.append(-1)
        kind.append(n.kind)
        node_ptr.append(len(in_col))
        tf_ptr.append(tf_ptr[-1] + m * (m - 1))
    p["nd_ptr"] = _i32(node_ptr)
    p["nd_in_col"] = _i32(in_col)
    p["nd_out_col"] = _i32(out_col)
    p["nd_kind"] = _i32(kind)
    p["nd_dem_row"] = _i32(dem_row)
    p["nd_tf_ptr"] = _i32(tf_ptr)
    p["n_virtual"] = 2 * n_virtual_nodes
    p["n_demand_rows"] = len(demand_nodes)
    p["demand_nodes"] = demand_nodes
    p["n_edges"] = tf_ptr[-1]

    # ---- route plan -------------------------------------------------------------------
    if path_finder is not None:
        od_keys = list(od_manager.od_flows.keys())
        od_index = {k: i for i, k in enumerate(od_keys)}
        link_index = {(l.start_node.node_id, l.end_node.node_id): l.index for l in links}
        rp = path_finder.export_route_plan(nodes, od_index, link_index)
        p.update({"rt_" + k: v for k, v in rp.items()})
        p["rt_scalars"] = _f64([path_finder.temp, path_finder.alpha, path_finder.beta,
                                path_finder.omega, path_finder.epsilon])
        p["n_od"] = len(od_keys)
        p["od_keys"] = od_keys
    else:
        from .path_finder import PathFinder  # noqa: F401  (empty plan, same keys)
        for k in ("routed_nodes", "routed_edge0", "routed_row0", "grp_node", "grp_up", "grp_od",
                  "grp_has_virtual", "opt_link", "opt_slot", "row_od", "term_opt", "term_row_entry"):
            p["rt_" + k] = _i32([])
        for k in ("opt_ptr", "row_ptr", "term_ptr"):
            p["rt_" + k] = _i32([0])
        p["rt_opt_dist"] = _f64([])
        p["rt_scalars"] = _f64([0.1, 1.0, 0.05, 0.05, 0.0])
        p["n_od"] = 0
        p["od_keys"] = []
    # per-node routed index (-1 = static fractions)
    routed_of = np.full(len(nodes), -1, dtype=np.int32)
    routed_of[p["rt_routed_nodes"]] = np.arange(len(p["rt_routed_nodes"]), dtype=np.int32)
    p["nd_routed"] = routed_of
    return p
