"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every
symbol include/speedy_b200.h declares, and refuses to compute without a device
(no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import speedy_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "speedy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
        n = m.group(1)
        if n.startswith(("sonic", "speedyBatch", "speedySessionPool", "getSonic")):
            names.add(n)
    return sorted(names)


def test_library_is_built_in_tree():
    import importlib
    importlib.import_module("speedy_b200.build").build()
    assert os.path.exists(sb.LIB_PATH)
    assert os.path.dirname(sb.LIB_PATH) == os.path.join(ROOT, "speedy_b200")


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(sb.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 50, names
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    # and the python mirror binds all of them
    bound = set(sb.lib()._declared)
    assert set(names) <= bound, sorted(set(names) - bound)


def test_sm100a_code_is_embedded():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", sb.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_geometry_matches_reference_formulae():
    # speedy.c:213-214, 335-338
    assert sb.frame_geometry(16000) == (240, 480, 160)
    assert sb.frame_geometry(22050) == (330, 660, 220)
    assert sb.frame_geometry(24000) == (360, 720, 240)
    assert sb.frame_geometry(48000) == (720, 1440, 480)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        sb.Batch(4)
    L = sb.lib()
    assert not L.sonicCreateStream(16000, 1)  # NULL, as on allocation failure (soniclib.c:95-111)


def test_default_config_matches_reference_defaults():
    cfg = sb.BatchConfig()
    sb.lib().speedyBatchDefaultConfig(C.byref(cfg))
    # soniclib.c:114-122
    assert cfg.speed == 1.0 and cfg.nonlinear_factor == 0.0
    assert abs(cfg.feedback_strength - 0.1) < 1e-7
    assert cfg.match_matlab == 0


def test_product_does_not_link_the_oracle():
    """The shipped library must not reference anything under oracle/."""
    import subprocess
    out = subprocess.run(["nm", "-D", sb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "ref_run" not in out
    for f in os.listdir(os.path.join(ROOT, "speedy_b200", "csrc")):
        for line in open(os.path.join(ROOT, "speedy_b200", "csrc", f)):
            if line.lstrip().startswith("#include"):
                assert "oracle" not in line, (f, line)
    for f in ("__init__.py", "build.py"):
        for line in open(os.path.join(ROOT, "speedy_b200", f)):
            if "import" in line or "CDLL" in line:
                assert "oracle" not in line, (f, line)


def test_ctypes_mirrors_match_the_header_structs(tmp_path):
    """The Python mirrors of the C structs (tests and bench.py go through them) have the header's
    sizes and field offsets: compiled with gcc from include/speedy_b200.h, no GPU involved."""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include "speedy_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(speedyBatchConfig), offsetof(speedyBatchConfig, max_write_frames),
         offsetof(speedyBatchConfig, threads_per_stream), offsetof(speedyBatchConfig, analysis_frame_step));
  printf("%zu %zu %zu\n", sizeof(speedySessionPoolConfig), offsetof(speedySessionPoolConfig, min_speed),
         offsetof(speedySessionPoolConfig, auto_step_sessions));
  printf("%zu %zu %zu\n", sizeof(speedySessionPoolStats), offsetof(speedySessionPoolStats, open_sessions),
         offsetof(speedySessionPoolStats, last_step_ms));
  return 0;
}
''')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    rows = [[int(x) for x in line.split()] for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines()]
    B, P, S = sb.BatchConfig, sb.SessionPoolConfig, sb.SessionPoolStats
    assert rows[0] == [C.sizeof(B), B.max_write_frames.offset, B.threads_per_stream.offset, B.analysis_frame_step.offset]
    assert rows[1] == [C.sizeof(P), P.min_speed.offset, P.auto_step_sessions.offset]
    assert rows[2] == [C.sizeof(S), S.open_sessions.offset, S.last_step_ms.offset]
