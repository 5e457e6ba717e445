"""The C-ABI library loads, exports every symbol include/actinon_b200.h declares, validates its
inputs, and refuses to render without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import actinon_b200 as acn
from actinon_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "actinon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = acn.load_library()
    declared = _declared_symbols()
    assert len(declared) > 60
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(api.EXPORTED_SYMBOLS) == declared


def test_version_and_defaults():
    lib = acn.load_library()
    assert b"actinon_b200" in lib.acn_version()
    o = api.Options(seed_mode=7, precision=3)
    lib.acn_options_default(C.byref(o))
    assert (o.seed_mode, o.precision, o.device) == (acn.SEED_POSITION_HASH, acn.PRECISION_F32, -1)


def test_bad_scene_is_rejected_with_error_code_not_abort():
    lib = acn.load_library()
    sc = acn.scenes.primitives(32, 24)
    flat = sc.flatten()
    fs = api.FlatSceneStruct()
    C.memmove(C.byref(fs), flat.ptr, C.sizeof(fs))
    nodes = (api.FlatNode * fs.n_nodes)()
    C.memmove(nodes, fs.nodes, C.sizeof(nodes))
    nodes[3].kind = 99                       # unknown type tag
    fs.nodes = C.cast(nodes, C.POINTER(api.FlatNode))
    p = C.c_void_p()
    rc = lib.acn_tracer_create(C.byref(fs), None, C.byref(p))
    assert rc == -2 and b"unknown kind" in lib.acn_last_error()
    nodes[3].kind = 1
    nodes[4].material = 1000                 # material index out of range
    rc = lib.acn_tracer_create(C.byref(fs), None, C.byref(p))
    assert rc == -2
    assert lib.acn_tracer_create(None, None, C.byref(p)) == -1


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="box has a GPU")
def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the tracer must fail loudly; nothing renders on the CPU."""
    sc = acn.scenes.primitives(32, 24)
    with pytest.raises(acn.AcnError) as e:
        acn.Tracer(sc.flatten())
    assert e.value.code == -3
    with pytest.raises(acn.AcnError):
        acn.device_count()


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under actinon_b200/ may import, link or name it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "actinon_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"oracle_lib|libacn_oracle|acn_oracle\.cpp|oracle/_ref|import\s+oracle", txt):
                    bad.append(f)
    assert not bad, bad
    sh = open(os.path.join(ROOT, "build.sh")).read()
    assert "oracle" not in sh


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree absent (GPU box)")
def test_reference_side_bridge_parses_against_the_reference_headers():
    """bridge/acn_bridge.c (the flattener + the replaced call site a maintainer adds) must parse against the
    reference's own objects.h / compound.h / distance.h / scene.h where they lie, with bridge/shim standing in for beth."""
    import subprocess
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "bridge"), "check"], capture_output=True, text=True)
    assert r.returncode == 0 and "syntax OK" in r.stdout, r.stdout + r.stderr
