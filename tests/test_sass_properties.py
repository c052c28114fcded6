"""Properties of the compiled sm_100a code that the measured performance depends on, checked on the built objects with
cuobjdump (CPU only): the streaming kernels really use TMA / mbarriers / packed fp32 math, and the adjoint's consumer
loop stays on the uniform datapath (a CALL or a top-level spin loop on the consumers' path makes ptxas
re-materialise the memory descriptor of every access with R2UR and guard the shuffles with BRA.DIV, +9-11 % time)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "percnn_b200", "build")


def _sass_counts(obj, pattern):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    path = os.path.join(BUILD, obj)
    if not os.path.exists(path):
        from percnn_b200.build import build_library
        build_library()
    out = subprocess.run([exe, "-sass", path], capture_output=True, text=True, timeout=600).stdout
    counts, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1) if re.search(pattern, m.group(1)) else None
            if name:
                counts[name] = {}
            continue
        if name:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                counts[name][op] = counts[name].get(op, 0) + 1
                if "BRA.DIV" in line:
                    counts[name]["BRA.DIV"] = counts[name].get("BRA.DIV", 0) + 1
    return counts


def test_adjoint_kernels_stay_on_the_uniform_datapath():
    counts = _sass_counts("tu_tma_bwd.o", r"k_gs3d_bwd_tma")
    assert len(counts) == 18            # 6 parameter slots x {periodic, slab up, slab down}
    for name, c in counts.items():
        assert c.get("UTMALDG", 0) >= 10, name          # TMA loads into the ring
        assert c.get("SYNCS", 0) >= 10, name            # mbarrier traffic
        assert c.get("FFMA2", 0) >= 90, name            # packed fp32 math
        assert c.get("R2UR", 0) <= 60, (name, c.get("R2UR"))          # 15 (periodic) / 34-36 (slab); 127 with a CALL in the loop
        assert c.get("BRA.DIV", 0) <= 5, (name, c.get("BRA.DIV"))     # 1 / 3; 29 when the loop is treated as divergent
        assert c.get("CALL", 0) == 0, name


def test_forward_kernels_use_tma_and_packed_math():
    counts = _sass_counts("tu_tma_fwd.o", r"k_gs3d_fwd_(tma|slab)")
    assert len(counts) == 18
    for name, c in counts.items():
        assert c.get("UTMALDG", 0) >= 10 and c.get("FFMA2", 0) >= 500 and c.get("USETMAXREG", 0) >= 2, name
