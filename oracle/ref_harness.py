"""TEST / BASELINE INFRASTRUCTURE ONLY -- live-reference harness.

Imports the *unmodified* reference (WaimenMak/PedNStream) with runtime shims only (matplotlib /
pettingzoo / gymnasium stubs, a `verbose` kwarg for create_network -- SURVEY.md Appendix C), from
a writable temp copy of /root/reference where that tree exists (the build container), otherwise
from `baseline/_ref/` -- the git-ignored install of the reference that `__graft_entry__.build()`
makes in the build container and that travels to the GPU box with the repository snapshot.
Used by `oracle/gen_golden.py` to produce the committed fixtures under `tests/golden/`, by
CPU-side tests (skipped when no reference tree is present) to pin the oracle restatement, and by
`bench.py --impl reference` (the reference arm times the reference's own code on the host cores).

Nothing in the product package, the `-m gpu` tests or `smoke()` imports this file.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REFERENCE_SUBDIRS = ("src", "data", "handlers", "rl")
_IGNORE = shutil.ignore_patterns("*.pt", "outputs", "*_agents_*", "ppo_agent_best", "*.gif", "__pycache__", "*.md")


def install_reference(force: bool = False) -> str | None:
    """Copy the reference's importable trees into baseline/_ref (git-ignored; ships with the snapshot).
    Returns the path, or None when /root/reference is not present."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "LTM")):
        return INSTALLED_ROOT if os.path.isdir(os.path.join(INSTALLED_ROOT, "src", "LTM")) else None
    stamp = os.path.join(INSTALLED_ROOT, ".installed")
    if os.path.exists(stamp) and not force:
        return INSTALLED_ROOT
    if os.path.isdir(INSTALLED_ROOT):
        shutil.rmtree(INSTALLED_ROOT)
    os.makedirs(INSTALLED_ROOT)
    for sub in REFERENCE_SUBDIRS:
        shutil.copytree(os.path.join(REFERENCE_ROOT, sub), os.path.join(INSTALLED_ROOT, sub), ignore=_IGNORE)
    open(stamp, "w").write("copied from /root/reference by oracle/ref_harness.install_reference\n")
    return INSTALLED_ROOT

LINK_FIELDS_F64 = ("inflow", "outflow", "cumulative_inflow", "cumulative_outflow",
                   "sending_flow", "receiving_flow", "back_gate_width_data")
LINK_FIELDS_F32 = ("num_pedestrians", "density", "speed", "travel_time",
                   "avg_travel_time", "link_flow")
LINK_FIELDS = LINK_FIELDS_F64 + LINK_FIELDS_F32

_REF_COPY = None


def reference_available() -> bool:
    return (os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "LTM"))
            or os.path.isdir(os.path.join(INSTALLED_ROOT, "src", "LTM")))


class _Anything:
    """Stand-in for any plotting class/function the reference imports but the hot path never calls."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


def _install_stubs():
    if "matplotlib" not in sys.modules:
        mpl = _StubModule("matplotlib")
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
        for name in ("pyplot", "animation", "cm", "colors", "patches", "lines", "collections", "gridspec",
                     "ticker", "colorbar", "figure", "axes"):
            sub = _StubModule(f"matplotlib.{name}")
            setattr(mpl, name, sub)
            sys.modules[f"matplotlib.{name}"] = sub
    if "pettingzoo" not in sys.modules:
        pz = types.ModuleType("pettingzoo")
        pz.ParallelEnv = type("ParallelEnv", (), {"__init__": lambda self, *a, **k: None})
        sys.modules["pettingzoo"] = pz
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")

        class Space:  # minimal duck type
            pass

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.dtype = dtype
                self.shape = tuple(shape) if shape is not None else np.shape(low)
                self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

            def sample(self):
                return np.random.uniform(self.low, self.high).astype(self.dtype)

        spaces.Space, spaces.Box = Space, Box
        gym.spaces = spaces
        sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces})


def reference_copy() -> str:
    """Writable copy of the reference tree (the reference mkdirs outputs/logs and data/)."""
    global _REF_COPY
    if _REF_COPY is None:
        if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "LTM")):
            _REF_COPY = INSTALLED_ROOT               # the installed copy is ours to write into
            return _REF_COPY
        dst = tempfile.mkdtemp(prefix="pns_ref_")
        for sub in REFERENCE_SUBDIRS:
            shutil.copytree(os.path.join(REFERENCE_ROOT, sub), os.path.join(dst, sub), ignore=_IGNORE)
        _REF_COPY = dst
    return _REF_COPY


_PREFIXES = ("src", "rl", "handlers")
_REF_MODULES = {}          # module name -> reference module object, kept out of sys.modules between uses


def _is_ref_name(name):
    return any(name == p or name.startswith(p + ".") for p in _PREFIXES)


class reference_modules:
    """Context manager: while active, the names `src.*`, `rl.*`, `handlers.*` in sys.modules are the reference's own
    modules (this repository has import-path shims under the same `src` name, which must not be what the
    reference's `from src.LTM...` statements find); afterwards whatever was there before is put back.  The
    reference module objects stay alive in `_REF_MODULES`, so classes handed out keep working."""

    def __enter__(self):
        _install_stubs()
        self.root = reference_copy()
        self.saved = {k: v for k, v in sys.modules.items() if _is_ref_name(k)}
        for k in self.saved:
            del sys.modules[k]
        sys.modules.update(_REF_MODULES)
        sys.path.insert(0, self.root)
        return self

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if _is_ref_name(k)]:
            _REF_MODULES[k] = sys.modules.pop(k)
        sys.modules.update(self.saved)
        try:
            sys.path.remove(self.root)
        except ValueError:
            pass
        return False


def import_reference():
    """Returns (Network class, NetworkEnvGenerator class) of the live reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    import logging
    with reference_modules():
        from src.LTM.network import Network  # noqa
        from src.utils.env_loader import NetworkEnvGenerator  # noqa
    assert "pednstream_b200" not in Network.__module__ and Network.__module__ == "src.LTM.network"
    assert not hasattr(Network, "engine"), "resolved to this repository's facade instead of the reference"
    logging.getLogger("src.LTM.network").setLevel(logging.ERROR)
    return Network, NetworkEnvGenerator


def make_reference_env(dataset, **kw):
    """reference rl/pz_pednet_env.PedNetParallelEnv with the `verbose` kwarg shim."""
    import logging
    _, Gen = import_reference()
    log = logging.getLogger("src.LTM.network")
    if not log.handlers:
        log.addHandler(logging.NullHandler())     # keeps Network.setup_logger from resetting the level
    log.setLevel(logging.ERROR)
    if not getattr(Gen.create_network, "_accepts_verbose", False):
        orig = Gen.create_network

        def create_network(self, yaml_file_path, *a, verbose=True, **k):
            return orig(self, yaml_file_path, *a, **k)

        create_network._accepts_verbose = True
        Gen.create_network = create_network
    with reference_modules():
        from rl.pz_pednet_env import PedNetParallelEnv
        return PedNetParallelEnv(dataset, **kw)


def create_network(name: str, steps_override: int | None = None, verbose: bool = True, default_link: dict = None,
                   params: dict = None, **kw):
    """reference NetworkEnvGenerator.create_network(name) (env_loader.py:81), optional
    simulation_steps override and overrides of the scenario's default_link block."""
    import logging
    Network, Gen = import_reference()
    g = Gen()
    if steps_override is not None or default_link or params:
        g.network_data = g.load_network_data(name)
        if steps_override is not None:
            g.config["params"]["simulation_steps"] = steps_override
        if default_link:
            g.config["params"]["default_link"].update(default_link)
        if params:
            g.config["params"].update(params)
    net = g.create_network(name, **kw)
    if net.logger is not None:
        net.logger.setLevel(logging.ERROR)
    return net, g


class DrawRecorder:
    """Records every in-step RNG draw of the reference keyed (site, link_id, time index).

    Sites: 'R1' release binomial (link.py:337,343), 'R2' activity binomial (:356),
    'R3' reverse-pedestrian binomial (:382), 'R4' speed noise normal (functions.py:133).
    Pure observation: the global MT19937 stream is not perturbed.
    """

    def __init__(self):
        self.draws = {}      # (site, link_id, t) -> value
        self.requests = {}   # (site, link_id, t) -> (n, p) / (loc, scale)
        self._ctx = None
        self._k = 0

    def install(self):
        with reference_modules():
            import src.LTM.link as rl
        self._rl = rl
        self._orig = (np.random.binomial, np.random.normal,
                      rl.Link.cal_sending_flow, rl.Link.cal_receiving_flow, rl.Link.update_speeds,
                      rl.Separator.cal_receiving_flow, rl.Separator.update_speeds)
        rec = self
        ob, on = np.random.binomial, np.random.normal

        def binomial(n, p, size=None):
            v = ob(n, p, size)
            if rec._ctx is not None:
                kind, link, t = rec._ctx
                lid = link.link_id
                if kind == "recv":
                    site = "R3"
                else:  # first binomial in cal_sending_flow is R2 when the diffusion branch skipped R1
                    ap = link.activity_probability
                    site = "R2" if (ap > 0 and float(p) == float(ap)) else "R1"
                rec.draws[(site, lid, t)] = int(v)
                rec.requests[(site, lid, t)] = (int(n), float(p))
                rec._k += 1
            return v

        def normal(loc=0.0, scale=1.0, size=None):
            v = on(loc, scale, size)
            if rec._ctx is not None and rec._ctx[0] == "speed":
                _, link, t = rec._ctx
                rec.draws[("R4", link.link_id, t)] = float(v)
            return v

        def wrap(fn, kind):
            def inner(self_, time_step, *a, **k):
                prev, prevk = rec._ctx, rec._k
                rec._ctx, rec._k = (kind, self_, time_step), 0
                try:
                    return fn(self_, time_step, *a, **k)
                finally:
                    rec._ctx, rec._k = prev, prevk
            return inner

        np.random.binomial, np.random.normal = binomial, normal
        rl.Link.cal_sending_flow = wrap(self._orig[2], "send")
        rl.Link.cal_receiving_flow = wrap(self._orig[3], "recv")
        rl.Link.update_speeds = wrap(self._orig[4], "speed")
        rl.Separator.cal_receiving_flow = wrap(self._orig[5], "recv")
        rl.Separator.update_speeds = wrap(self._orig[6], "speed")
        return self

    def uninstall(self):
        rl = self._rl
        (np.random.binomial, np.random.normal, rl.Link.cal_sending_flow, rl.Link.cal_receiving_flow,
         rl.Link.update_speeds, rl.Separator.cal_receiving_flow, rl.Separator.update_speeds) = self._orig


def collect_link_arrays(net) -> dict:
    """Stack every per-link history array into [S+1, L] matrices in network.links order."""
    out = {}
    links = list(net.links.values())
    for f in LINK_FIELDS:
        out[f] = np.stack([np.asarray(getattr(l, f)) for l in links], axis=1)
    out["link_keys"] = np.array(list(net.links.keys()), dtype=np.int64)
    return out


def array_digest(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()
