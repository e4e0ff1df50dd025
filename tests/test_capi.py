"""The C-ABI shared library loads and exports every symbol include/pns_b200.h declares (CPU)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT
from pednstream_b200 import _native


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pns_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pns_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    assert os.path.exists(_native.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared_symbols()
    assert set(names) == set(_native.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.pns_abi_version() == _native.ABI_VERSION


def test_library_targets_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout: compile a probe against the header with gcc."""
    import tempfile
    src = '#include <stdio.h>\n#include "pns_b200.h"\nint main(){printf("%zu %zu %zu", sizeof(pns_net), sizeof(pns_state), sizeof(pns_step_io));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_native.PnsNet), ctypes.sizeof(_native.PnsState),
                     ctypes.sizeof(_native.PnsStepIO)]
