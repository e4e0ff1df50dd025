"""Host-side link views.

The simulation state lives in HBM as time-major structure-of-arrays history
(`state.StateStore`); these objects keep the reference's per-link read/write surface
(reference: src/LTM/link.py:4-131, 418-478) on top of it:

* the 13 per-link time series (`inflow`, ..., `link_flow`, `back_gate_width_data`,
  `separator_width_data`) are numpy views of a lazily synchronised host mirror, one column
  per link, length S+1, reference dtypes (fp64 / fp32);
* gate / separator widths are scalars in a small host table that is pushed to the device
  before the next step; the setters keep the reference's coupling rules
  (`back_gate_width` <-> reverse `front_gate_width`, link.py:110-126; separator widths sum to
  the corridor width and drive both gate widths, link.py:462-478).

No physics here: sending/receiving flow, density and speed are computed by the CUDA kernels in
`csrc/` (and restated for checking in `oracle/`).
"""
from __future__ import annotations

import numpy as np

FD_TYPES = {"yperman": 0, "greenshields": 1, "smulders": 2}

_F64_FIELDS = ("inflow", "outflow", "cumulative_inflow", "cumulative_outflow",
               "sending_flow", "receiving_flow")
_LINK_F64_FIELDS = ("back_gate_width_data",)
_F32_FIELDS = ("num_pedestrians", "density", "speed", "travel_time", "avg_travel_time", "link_flow")


def _series_property(field):
    def getter(self):
        return self._store.column(field, self._col)
    getter.__name__ = field
    return property(getter, doc=f"time series `{field}` (numpy view, length S+1)")


class BaseLink:
    """Virtual origin/destination link: only the six fp64 counters (link.py:4-28)."""

    is_virtual = True

    def __init__(self, store, col, link_id, start_node, end_node):
        self._store = store
        self._col = col
        self.link_id = link_id
        self.start_node = start_node
        self.end_node = end_node

    def update_speeds(self, time_step: int):
        return


for _f in _F64_FIELDS:
    setattr(BaseLink, _f, _series_property(_f))


class Link(BaseLink):
    """Physical directed link; `index` is its column in the device SoA (network.links order)."""

    is_virtual = False
    is_separator = False

    def __init__(self, store, index, link_id, start_node, end_node, simulation_steps, unit_time, **kw):
        super().__init__(store, index, link_id, start_node, end_node)
        self.index = index
        self.length = kw["length"]
        self._width = kw["width"]
        self.free_flow_speed = kw["free_flow_speed"]
        self.k_critical = kw["k_critical"]
        self.k_jam = kw["k_jam"]
        self.capacity = self.free_flow_speed * self.k_critical
        self.shockwave_speed = self.capacity / (self.k_jam - self.k_critical)
        self.current_speed = self.free_flow_speed
        self.max_travel_time = self.length / 0.05
        self.bi_factor = kw.get("bi_factor", 1)
        self.fd_type = kw.get("fd_type", "yperman")
        if self.fd_type not in FD_TYPES:
            raise ValueError(f"Unknown model type: {self.fd_type}")
        self.speed_noise_std = kw.get("speed_noise_std", 0)
        self.exponent = 0.8
        self.unit_time = unit_time
        self.gamma = kw.get("gamma", 2e-3)
        self.activity_probability = kw.get("activity_probability", 0.0)
        self.reverse_link = None
        # derived integers the kernels take as inputs: computed here with Python semantics
        # (fp32 free-flow travel time, round-half-even) so they are bit-identical to link.py:83-89
        self.travel_time0 = np.float32(min(self.length / self.free_flow_speed, self.max_travel_time))
        self.free_flow_tau = round(self.travel_time0 / unit_time)
        self.avg_travel_time_window = round(100 / unit_time)
        self.shockwave_tau = round(self.length / (self.shockwave_speed * unit_time))
        store.add_link_gate(index, self._initial_gate(), self.is_separator)

    def _initial_gate(self):
        return self._width

    # --- widths ------------------------------------------------------------------------
    # One value per directed link lives in the store (and on the device): the back gate width.
    # The reference's setters keep front_gate(l) == back_gate(reverse(l)) (link.py:110-126), so the
    # front gate is read from / written to the reverse link's entry.
    @property
    def width(self):
        return self._width

    @property
    def _back_gate_width(self):
        return self._store.get_gate(self.index)

    @_back_gate_width.setter
    def _back_gate_width(self, value):
        self._store.set_gate(self.index, value)

    @property
    def _front_gate_width(self):
        return self._store.get_gate(self.index ^ 1)

    @_front_gate_width.setter
    def _front_gate_width(self, value):
        self._store.set_gate(self.index ^ 1, value)

    @property
    def front_gate_width(self):
        return self._front_gate_width

    @front_gate_width.setter
    def front_gate_width(self, value):
        self._front_gate_width = value

    @property
    def back_gate_width(self):
        return self._back_gate_width

    @back_gate_width.setter
    def back_gate_width(self, value):
        self._back_gate_width = value

    @property
    def area(self):
        return self.length * self.width

    def get_density(self, time_step: int):
        """Shared-corridor density (link.py:190-197): both directions' pedestrians over one area."""
        other = 0
        if self.reverse_link is not None:
            other = self.reverse_link.num_pedestrians[time_step]
        return (self.num_pedestrians[time_step] + other) / self.area


for _f in _LINK_F64_FIELDS + _F32_FIELDS:
    setattr(Link, _f, _series_property(_f))


class Separator(Link):
    """Corridor whose two directions split the width with a movable separator (link.py:418-478)."""

    is_separator = True

    def _initial_gate(self):
        return self._width / 2

    separator_width_data = _series_property("separator_width_data")

    # A separator's lane width, front gate and back gate are one value (link.py:462-478); the
    # reverse direction holds the rest of the corridor.
    @property
    def _separator_width(self):
        return self._store.get_gate(self.index)

    @_separator_width.setter
    def _separator_width(self, value):
        self._store.set_gate(self.index, value)

    _front_gate_width = _separator_width
    _back_gate_width = _separator_width

    @property
    def front_gate_width(self):
        return self._separator_width

    @property
    def back_gate_width(self):
        return self._separator_width

    @property
    def area(self):
        return self.length * self._separator_width

    @property
    def separator_width(self):
        return self._separator_width

    @separator_width.setter
    def separator_width(self, value):
        self._separator_width = value
        if self.reverse_link:
            self.reverse_link._separator_width = self._width - value

    def get_density(self, time_step: int):
        return self.density[time_step]
