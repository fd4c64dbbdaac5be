"""CPU tests of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports every symbol
include/fsim.h declares; the ctypes mirrors have the C struct sizes; creating a simulation without a GPU fails
loudly (no CPU fallback)."""
import ctypes
import importlib
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as ol

fs = importlib.import_module("fluid-sim_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(fs.LIB_PATH):
        fs.build()
    return fs.lib()


def declared_symbols():
    src = open(fs.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsim_[a-z_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    assert sorted(fs.EXPORTS) == syms
    for name in syms:
        assert hasattr(built_lib, name), name


def test_library_has_sm100a_code():
    if not os.path.exists(fs.LIB_PATH):
        fs.build()
    out = subprocess.run(["cuobjdump", "-lelf", fs.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_c(built_lib):
    prog = r'''
#include <stdio.h>
#include "fsim.h"
int main(void){printf("%zu %zu %zu %zu\n", sizeof(fsim_config), sizeof(fsim_options), sizeof(fsim_stats), sizeof(fsim_host_mirror));return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(fs.FsimConfig), ctypes.sizeof(fs.FsimOptions), ctypes.sizeof(fs.FsimStats),
                     ctypes.sizeof(fs.FsimHostMirror)]


def test_struct_fields_match_c_by_name_and_offset(built_lib):
    """every field of the ctypes mirrors exists in include/fsim.h under the same name at the same offset"""
    pairs = [("fsim_config", fs.FsimConfig), ("fsim_options", fs.FsimOptions), ("fsim_stats", fs.FsimStats),
             ("fsim_host_mirror", fs.FsimHostMirror)]
    lines, want = [], []
    for cname, klass in pairs:
        for fname, _ in klass._fields_:
            lines.append('printf("%%zu\\n", offsetof(%s, %s));' % (cname, fname))
            want.append(getattr(klass, fname).offset)
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "fsim.h"\nint main(void){%s return 0;}\n' % " ".join(lines)
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert got == want


def test_default_options_are_reference_constants(built_lib):
    opt = fs.FsimOptions()
    built_lib.fsim_default_options(ctypes.byref(opt))
    assert opt.pcgTol == 1e-12 and opt.pcgMaxIters == 200  # reference src/FluidSim2D.cpp:429,453
    assert opt.seedParticles == 1 and opt.slDoubleBuffer == 0


def test_create_rejects_bad_input_or_missing_gpu(built_lib):
    import torch
    cells = ol.dam_break_cells(32)
    bad = cells.copy()
    bad[0, 5] = ol.EMPTY  # border must be SOLID
    with pytest.raises(fs.FsimError):
        fs.FluidSim2D(bad, dt=0.005, dx=0.04)
    if not torch.cuda.is_available():
        with pytest.raises(fs.FsimError) as e:
            fs.FluidSim2D(cells, dt=0.005, dx=0.04)
        assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "fluid-sim_b200")
    for base, _, files in os.walk(pkg):
        if "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(base, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle/" not in txt, (base, f)
    out = subprocess.run(["ldd", fs.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "fsim_ref" not in out


def test_checkpoint_load_refuses_bad_files_without_touching_the_gpu(tmp_path):
    """fsim_checkpoint_load validates the file before it creates anything on a device"""
    import ctypes
    L = fs.lib()
    h = ctypes.c_void_p()
    assert L.fsim_checkpoint_load(str(tmp_path / "missing.ckp").encode(), None, ctypes.byref(h)) == -4  # FSIM_E_STATE
    bad = tmp_path / "bad.ckp"
    bad.write_bytes(b"NOTACKPT" + bytes(200))
    assert L.fsim_checkpoint_load(str(bad).encode(), None, ctypes.byref(h)) == -1  # FSIM_E_INVALID
    assert b"FSIMCKP1" in L.fsim_last_error()
    assert not h.value


def test_headless_driver_rejects_unknown_options():
    import subprocess
    exe = os.path.join(os.path.dirname(fs.LIB_PATH), "..", "bin", "fsim_run")
    if not os.path.exists(exe):
        pytest.skip("fsim_run not built")
    r = subprocess.run([exe, "--no-such-option"], capture_output=True, text=True)
    assert r.returncode == 2 and "unknown option" in r.stderr
