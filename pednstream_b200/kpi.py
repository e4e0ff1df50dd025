"""Episode KPIs (reference rl/rl_utils.py:770-1512).  The reference computes them from the JSON files
written by handlers/output_handler.py; here the sums come from the device (C-ABI `pns_kpi`) and this
module forms the same dictionaries of totals and ratios."""
from __future__ import annotations

import numpy as np

from . import _native


def kpi_dict(raw: np.ndarray) -> list:
    """raw [R, len(KPI_NAMES)] -> one dict per replica with the reference's result keys:
    compute_network_throughput (throughput, completed_demand, total_demand),
    compute_served_trips_rate (served_trips_rate, total_inflow, total_outflow),
    compute_average_travel_time_spent (avg_travel_time_spent, total_person_time, total_trips),
    compute_total_network_delay (total_delay, delay_intensity, total_person_time_moving),
    compute_network_congestion_metric (congestion_time, avg_congestion_density, congestion_fraction,
    total_area_time), compute_network_travel_time (avg_travel_time)."""
    raw = np.atleast_2d(np.asarray(raw, dtype=np.float64))
    out = []
    for row in raw:
        k = dict(zip(_native.KPI_NAMES, (float(x) for x in row)))
        ratio = lambda a, b: a / b if b > 0 else 0.0
        out.append({
            "throughput": ratio(k["total_outflow"], k["total_demand"]),
            "completed_demand": k["total_outflow"], "total_demand": k["total_demand"],
            "served_trips_rate": ratio(k["total_outflow"], k["total_inflow"]),
            "total_inflow": k["total_inflow"], "total_outflow": k["total_outflow"],
            "avg_travel_time_spent": ratio(k["person_time"], k["total_inflow"]),
            "total_person_time": k["person_time"], "total_trips": k["total_inflow"],
            "total_delay": k["total_delay"], "total_person_time_moving": k["person_time_moving"],
            "delay_intensity": ratio(k["total_delay"], k["person_time_moving"]),
            "congestion_time": k["congestion_time"], "total_area_time": k["area_time"],
            "avg_congestion_density": ratio(k["congestion_time"], k["area_time"]),
            "congestion_fraction": ratio(k["congested_steps"], k["steps"]) if k["area_time"] > 0 else 0.0,
            "avg_travel_time": k["avg_travel_time"],
        })
    return out
