"""Host mirror of the device history + the mutable width table.

Device layout (see DESIGN.md "data layout in HBM"): two torch tensors owned by the engine,
`hist64[F64, S+1, C64*R]` (fp64) and `hist32[F32, S+1, L*R]` (fp32), time-major, replica index
fastest.  For the single-network facade (R=1) this class exposes each field as a lazily
synchronised `[S+1, columns]` numpy array whose columns are handed out to the Link objects
(`link.inflow` etc.), so consumers that call `.tolist()`, `len()` or index `[t]` keep working
(reference consumers: handlers/output_handler.py:34-88, src/utils/visualizer.py:124-137).

A field's host array is created on first access and refreshed from the device only for rows
that changed since the last access; before any step has run it holds the reference's initial
state (link.py:12-17, 56, 82-97).
"""
from __future__ import annotations

import numpy as np

F64_FIELDS = ("inflow", "outflow", "cumulative_inflow", "cumulative_outflow",
              "sending_flow", "receiving_flow", "back_gate_width_data", "separator_width_data")
F32_FIELDS = ("num_pedestrians", "density", "speed", "travel_time", "avg_travel_time", "link_flow")
F64_INDEX = {f: i for i, f in enumerate(F64_FIELDS)}
F32_INDEX = {f: i for i, f in enumerate(F32_FIELDS)}


class StateStore:
    def __init__(self, simulation_steps: int):
        self.S = simulation_steps
        self.n_links = 0           # physical links (columns 0..L-1)
        self.n_virtual = 0         # virtual links (columns L..L+V-1 of the fp64 fields)
        self._gate_rows = []       # back gate width / separator lane width per physical link
        self._sep_rows = []
        self.gate = None           # float64 [L] once frozen
        self.sep_np64 = None       # separator width was assigned a numpy float64 (dtype ledger)
        self.widths_dirty = True
        self._host = {}            # field -> ndarray
        self._synced_t = {}        # field -> last device step mirrored
        self.engine = None         # set by Network when the device runtime is attached
        self._init = None          # per-link initial values (set by freeze)

    # ---- construction -------------------------------------------------------------------
    def add_link_gate(self, index, value, is_separator):
        assert index == len(self._gate_rows)
        self._gate_rows.append(value)
        self._sep_rows.append(bool(is_separator))

    def freeze(self, n_links, n_virtual, tt0, window, bgw0):
        self.n_links, self.n_virtual = n_links, n_virtual
        self.gate = np.asarray(self._gate_rows, dtype=np.float64).reshape(n_links)
        self.is_separator = np.asarray(self._sep_rows, dtype=bool).reshape(n_links)
        self.has_separators = bool(self.is_separator.any())
        self.sep_np64 = np.zeros(n_links, dtype=np.int32)
        self._init = dict(tt0=np.asarray(tt0, dtype=np.float32), window=int(window),
                          bgw0=np.asarray(bgw0, dtype=np.float64))

    # ---- widths ---------------------------------------------------------------------------
    def get_gate(self, index):
        if self.gate is None:
            return self._gate_rows[index]
        return float(self.gate[index])

    def set_gate(self, index, value):
        if self.gate is None:
            self._gate_rows[index] = value
            return
        self.gate[index] = value
        if self.is_separator[index]:
            self.sep_np64[index] = 1 if isinstance(value, np.floating) else 0
        self.widths_dirty = True

    # ---- history views ----------------------------------------------------------------------
    def _initial(self, field):
        S, L = self.S, self.n_links
        if field in F64_INDEX:
            C = L + self.n_virtual
            if field in ("sending_flow", "receiving_flow"):
                return np.full((S + 1, C), -1.0)
            if field == "back_gate_width_data":
                a = np.zeros((S + 1, C)); a[:, :L] = self._init["bgw0"][None, :]; return a
            if field == "separator_width_data":
                a = np.zeros((S + 1, C)); a[:, :L] = np.where(self.is_separator, self._init["bgw0"] / 2, 0.0)[None, :]; return a
            return np.zeros((S + 1, C))
        a = np.zeros((S + 1, L), dtype=np.float32)
        if field == "travel_time":
            a[0] = self._init["tt0"]
        elif field == "avg_travel_time":
            a[: self._init["window"]] = self._init["tt0"][None, :]
        return a

    def field(self, name):
        """[S+1, columns] host array of one field, synchronised with the device."""
        arr = self._host.get(name)
        if arr is None:
            arr = self._initial(name)
            self._host[name] = arr
            self._synced_t[name] = 0
        eng = self.engine
        if eng is not None and eng.t_done > 0 and self._synced_t[name] < eng.t_done:
            lo = max(0, self._synced_t[name] - 1)       # row t-1 is rewritten by step t
            eng.read_rows(name, lo, eng.t_done, arr)
            self._synced_t[name] = eng.t_done
        return arr

    def column(self, name, col):
        return self.field(name)[:, col]

    def invalidate(self):
        """Forget mirrored rows (used when the device state is reset)."""
        self._host.clear()
        self._synced_t.clear()
