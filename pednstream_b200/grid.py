"""Sparse constructor for n x n lattice networks (BASELINE config 4: 512 x 512 nodes).

The reference can only build a network from a dense adjacency matrix (src/LTM/network.py:151-152,
194-248), which for 262 144 nodes would need ~550 GB.  This builder emits the same *plan* the
generic path compiles (pednstream_b200/plan.py) -- node creation order, (i,j),(j,i) link pairs by
ascending i then j, slot order "virtual first, then neighbours by ascending id", node classes of
network.py:141-167 -- straight from the lattice rule of data/create_grid.py:3-21 (node id =
row*size + col, edges to the right and down neighbour), without per-link Python objects.
`tests/test_grid.py` checks it against the generic constructor on small lattices.
"""
from __future__ import annotations

import numpy as np

from .link import FD_TYPES

DEFAULT_LINK = dict(length=50, width=4, free_flow_speed=1.1, k_critical=2, k_jam=6, gamma=0.01,
                    speed_noise_std=0.05, fd_type="yperman", activity_probability=0, bi_factor=1)


def grid_adjacency(size: int) -> np.ndarray:
    """Dense adjacency of the lattice (small sizes only; used by tests and examples)."""
    n = size * size
    adj = np.zeros((n, n), dtype=int)
    for r in range(size):
        for c in range(size):
            i = r * size + c
            if c < size - 1:
                adj[i, i + 1] = adj[i + 1, i] = 1
            if r < size - 1:
                adj[i, i + size] = adj[i + size, i] = 1
    return adj


def default_origins(size: int, stride: int = 64):
    """The 4 corners plus every `stride`-th node of the boundary walk."""
    n = size
    walk = ([(0, c) for c in range(n)] + [(r, n - 1) for r in range(1, n)] +
            [(n - 1, c) for c in range(n - 2, -1, -1)] + [(r, 0) for r in range(n - 2, 0, -1)])
    picked = {0, n - 1, (n - 1) * n, n * n - 1}
    picked.update(r * n + c for (r, c) in walk[::stride])
    return sorted(picked)


def build_grid_plan(size: int, sim_steps: int, unit_time=10, link=None, origins=None,
                    peak_lambda=50, base_lambda=30, demand_seed=0, locality_order=False, destinations=(),
                    node_model="classic"):
    """Returns (plan, gate[L], tf_static[n_edges], demand[S, rows]) for `Engine`.
    locality_order=True lists nodes by id instead of the reference's creation order.
    destinations: nodes that get virtual O/D links like origins but draw no demand (network.py:139-167); the
    route plan of a routed lattice is attached by `build_routed_grid_plan`.  demand_seed=None: the caller has
    positioned numpy's global stream (the reference draws demand from it while it creates the nodes).
    node_model: Network.assign_flows_type -- 'optimal' makes every regular node solve the linear program of
    node.py:249-271 (kind 2)."""
    lk = dict(DEFAULT_LINK)
    lk.update(link or {})
    n = size
    N = n * n
    origins = default_origins(n) if origins is None else sorted(origins)
    is_origin = np.zeros(N, dtype=bool)
    is_origin[origins] = True
    is_od = is_origin.copy()
    is_od[list(destinations)] = True

    ids = np.arange(N)
    r, c = ids // n, ids % n
    has_right, has_down = c < n - 1, r < n - 1
    # pair index of edge (i, i+1) / (i, i+n): pairs are enumerated by ascending i, right before down
    n_pairs_at = has_right.astype(np.int64) + has_down.astype(np.int64)
    first_pair = np.concatenate([[0], np.cumsum(n_pairs_at)[:-1]])
    pair_right = np.where(has_right, first_pair, -1)
    pair_down = np.where(has_down, first_pair + has_right, -1)
    P = int(n_pairs_at.sum())
    L = 2 * P

    # node creation order of init_nodes_and_links: i, then its not-yet-created neighbours j > i
    order = np.empty(N, dtype=np.int64)
    seen = np.zeros(N, dtype=bool)
    k = 0
    for i in range(N):
        if not seen[i]:
            seen[i] = True; order[k] = i; k += 1
        if has_right[i] and not seen[i + 1]:
            seen[i + 1] = True; order[k] = i + 1; k += 1
        if has_down[i] and not seen[i + n]:
            seen[i + n] = True; order[k] = i + n; k += 1
    assert k == N
    if locality_order:
        # Node order only fixes the visiting order of the reference's sequential RNG; with
        # counter-based draws any order gives the same result, and id order keeps the links of
        # neighbouring threads adjacent in memory.
        order = ids.copy()

    degree = (r > 0).astype(int) + (c > 0) + has_right + has_down
    # network.py:141-167: degree 2 -> one-to-one unless O/D; degree 1 -> one-to-one + virtual; else regular
    virtual = is_od | (degree == 1)
    kind = np.where((degree == 2) & ~is_od, 0, np.where(degree == 1, 0, 1)).astype(np.int32)
    if node_model == "optimal":
        kind[kind == 1] = 2
    vrank = np.full(N, -1, dtype=np.int64)           # rank among virtual-link owners, creation order
    vo = order[virtual[order]]
    vrank[vo] = np.arange(len(vo))

    m = degree + virtual
    m_o = m[order]
    nd_ptr = np.concatenate([[0], np.cumsum(m_o)]).astype(np.int32)
    in_col = np.empty(nd_ptr[-1], dtype=np.int32)
    out_col = np.empty(nd_ptr[-1], dtype=np.int32)
    # neighbour slots in ascending id: up (i-n), left (i-1), right (i+1), down (i+n)
    up_pair = np.where(r > 0, pair_down[np.maximum(ids - n, 0)], -1)
    left_pair = np.where(c > 0, pair_right[np.maximum(ids - 1, 0)], -1)
    cursor = nd_ptr[:-1].astype(np.int64).copy()
    sel = virtual[order]
    in_col[cursor[sel]] = L + 2 * vrank[order[sel]]
    out_col[cursor[sel]] = L + 2 * vrank[order[sel]] + 1
    cursor[sel] += 1
    # (pair, node is the larger id of the pair?)  forward link 2p runs small->large id
    for pair_of, node_is_large in ((up_pair, True), (left_pair, True), (pair_right, False), (pair_down, False)):
        pp = pair_of[order]
        sel = pp >= 0
        fwd = 2 * pp[sel]
        if node_is_large:      # incoming = forward (nbr -> i), outgoing = reverse (i -> nbr)
            in_col[cursor[sel]] = fwd
            out_col[cursor[sel]] = fwd + 1
        else:                  # incoming = reverse (nbr -> i), outgoing = forward
            in_col[cursor[sel]] = fwd + 1
            out_col[cursor[sel]] = fwd
        cursor[sel] += 1
    assert (cursor == nd_ptr[1:]).all()

    dem_row = np.where(virtual[order], vrank[order], -1).astype(np.int32)
    edges = (m_o * (m_o - 1)).astype(np.int64)
    tf_ptr = np.concatenate([[0], np.cumsum(edges)]).astype(np.int32)
    assert (out_col == (in_col ^ 1)).all()
    meta = np.zeros((N, 4), dtype=np.int32)
    meta[:, 0] = nd_ptr[:-1]
    meta[:, 1] = (m_o | (kind[order].astype(np.int64) << 8)).astype(np.int32)
    meta[:, 2] = dem_row
    meta[:, 3] = tf_ptr[:-1]
    from .plan import link_slots, node_stride, slot_links
    max_degree = int(m_o.max())
    stride = node_stride(max_degree)
    lk_slots = link_slots(meta, in_col, L, stride)

    from .plan import class_record, CLASS_DTYPE
    rec = class_record(lk["length"], lk["width"], lk["free_flow_speed"], lk["k_critical"], lk["k_jam"],
                       lk["gamma"], lk["activity_probability"], lk["bi_factor"], lk["speed_noise_std"],
                       lk["fd_type"], False, unit_time)
    plan = dict(
        n_links=L, n_nodes=N, sim_steps=int(sim_steps), unit_time=float(unit_time),
        window=int(round(100 / unit_time)),
        classes=np.array([rec], dtype=CLASS_DTYPE).reshape(1), lk_class=np.zeros(L, dtype=np.int32),
        lk_width=np.full(L, float(lk["width"])), has_separators=False,
        nd_meta=meta, nd_in_col=in_col, nd_routed=np.full(N, -1, dtype=np.int32), lk_slots=lk_slots,
        nd_in_link=slot_links(meta, in_col, stride),
        max_degree=max_degree, nd_stride=stride,
        n_virtual=2 * int(virtual.sum()), n_demand_rows=int(virtual.sum()), n_edges=int(tf_ptr[-1]),
        n_od=0, od_keys=[], demand_nodes=[], node_order=order,
    )
    from .plan import attach_empty_route_plan
    attach_empty_route_plan(plan)
    plan["lattice"] = dict(size=n, pair_right=pair_right, pair_down=pair_down, virtual=virtual, vrank=vrank,
                           length=float(lk["length"]))

    gate = np.full(L, float(lk["width"]))
    tf_static = np.repeat(1.0 / np.maximum(m_o - 1, 1), edges)      # uniform 1/(m-1), network.py:269-271

    # demand rows in virtual-owner creation order; Poisson around two gaussian peaks
    # (od_manager.py:145-155), drawn from the global numpy RNG in node creation order like the reference
    S = int(sim_steps)
    demand = np.zeros((S, int(virtual.sum())), dtype=np.float64)
    tgrid = np.arange(S)
    width2 = 2 * (S / 20) ** 2
    lam = (base_lambda + peak_lambda * np.exp(-(tgrid - S / 4) ** 2 / width2)
           + peak_lambda * np.exp(-(tgrid - 3 * S / 4) ** 2 / width2))
    if demand_seed is not None:
        np.random.seed(demand_seed)
    for node in vo:
        if is_origin[node]:
            demand[:, vrank[node]] = np.random.poisson(lam=lam)
    return plan, gate, tf_static, demand


# ---- routed lattices (BASELINE config 4b) ------------------------------------------------------------------
class _Stub:
    """Attribute bag standing in for a Node / Link of the generic constructor (PathFinder reads a handful of
    attributes and hangs its per-node route structures on the node)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class _LatticeLinks:
    """`network.links` of a lattice without per-link objects: iteration in the reference's order ((i,j),(j,i) by
    ascending i, then j), membership and index by arithmetic."""

    def __init__(self, lat):
        self.n, self.pair_right, self.pair_down, self.length = lat["size"], lat["pair_right"], lat["pair_down"], lat["length"]
        self._link = _Stub(length=self.length)

    def index(self, key):
        u, v = key
        a, b = (u, v) if u < v else (v, u)
        n = self.n
        if a < 0 or b >= n * n:
            return -1
        if b == a + 1 and a % n != n - 1:
            p = self.pair_right[a]
        elif b == a + n:
            p = self.pair_down[a]
        else:
            return -1
        return int(2 * p + (0 if u < v else 1))

    def __contains__(self, key):
        return self.index(key) >= 0

    def __getitem__(self, key):
        k = self.index(key)
        if k < 0:
            raise KeyError(key)
        return k if self._as_index else self._link

    _as_index = False

    def as_index(self):
        view = _LatticeLinks.__new__(_LatticeLinks)
        view.__dict__.update(self.__dict__)
        view._as_index = True
        return view

    def items(self):
        n = self.n
        for i in range(n * n):
            if i % n != n - 1:
                yield (i, i + 1), self._link
                yield (i + 1, i), self._link
            if i + n < n * n:
                yield (i, i + n), self._link
                yield (i + n, i), self._link


class _LatticeNodes:
    """`network.nodes` of a lattice, materialised on demand (only nodes on OD paths are ever touched): slot order
    is the reference's -- the virtual O/D link first, then neighbours by ascending id."""

    def __init__(self, lat, index_of):
        self.lat, self.index_of, self._made = lat, index_of, {}

    def __getitem__(self, i):
        node = self._made.get(i)
        if node is not None:
            return node
        n = self.lat["size"]
        r, c = divmod(int(i), n)
        nbrs = [j for j, ok in ((i - n, r > 0), (i - 1, c > 0), (i + 1, c < n - 1), (i + n, r < n - 1)) if ok]
        me = _Stub(node_id=int(i), index=int(self.index_of[i]), ods_in_turns={})
        inc, out = [], []
        if self.lat["virtual"][i]:
            inc.append(_Stub(start_node=None, end_node=me))
            out.append(_Stub(start_node=me, end_node=None))
        for j in nbrs:
            other = _Stub(node_id=int(j))
            inc.append(_Stub(start_node=other, end_node=me))
            out.append(_Stub(start_node=me, end_node=other))
        me.incoming_links, me.outgoing_links = inc, out
        me.source_num = me.dest_num = len(inc)
        self._made[i] = me
        return me


def build_routed_grid_plan(size: int, sim_steps: int, origins, destinations, params=None, unit_time=10, link=None,
                           peak_lambda=50, base_lambda=30, demand_seed=0, locality_order=False):
    """A routed lattice (BASELINE config 4b) without the dense adjacency matrix and per-link objects of the generic
    constructor: `build_grid_plan` for topology and demand, then the reference's own route setup -- networkx
    k-shortest simple paths on a graph whose edges are inserted in `network.links` order, turn structures per path
    node (path_finder.py:114-142, 199-234, 460-546; `PathFinder` is shared with the generic constructor, so ties
    are broken identically) -- over lazily materialised nodes.  Returns (plan, gate, tf_static, demand, od_w);
    equal to the plan the generic constructor compiles (tests/test_grid.py)."""
    from .od_manager import ODManager
    from .path_finder import PathFinder
    from .plan import attach_route_plan
    origins, destinations = list(origins), list(destinations)
    plan, gate, tf, demand = build_grid_plan(size, sim_steps, unit_time=unit_time, link=link, origins=origins,
                                             peak_lambda=peak_lambda, base_lambda=base_lambda,
                                             demand_seed=demand_seed, locality_order=locality_order,
                                             destinations=destinations)
    lat = plan["lattice"]
    order = plan["node_order"]
    index_of = np.empty(size * size, dtype=np.int64)
    index_of[order] = np.arange(size * size)
    links = _LatticeLinks(lat)
    nodes = _LatticeNodes(lat, index_of)
    odm = ODManager(sim_steps)
    odm.logger.disabled = True
    odm.init_od_flows(origins, destinations, None)
    pf = PathFinder(links, params=params or {}, controller_nodes=set(), controller_links=[], logger=None)
    pf.find_od_paths(od_pairs=odm.od_flows.keys(), nodes=nodes)
    on_path = sorted((nodes[i] for i in pf.nodes_in_paths), key=lambda nd: nd.index)
    attach_route_plan(plan, pf, odm, on_path, links.as_index())
    routed_of = np.full(size * size, -1, dtype=np.int32)
    routed_of[plan["rt_routed_nodes"]] = np.arange(len(plan["rt_routed_nodes"]), dtype=np.int32)
    plan["nd_routed"] = routed_of
    plan["od_paths"] = pf.od_paths
    od_w = np.stack([odm.od_flows[k] for k in plan["od_keys"]], axis=1)
    return plan, gate, tf, demand, od_w
