"""not gpu: libvofod_cuda.so loads and exports every entry point include/vofod_cuda.h declares; the ctypes struct mirrors
have the C layout.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vofod_cuda.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vofod_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vofod_b200 import capi
    lib = capi.load_library()
    names = declared_functions()
    assert len(names) >= 48
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vofod_cuda.h but not exported"
        assert n in lib._vofod_sigs, f"{n} has no ctypes signature in capi.py"


def test_struct_layouts_match_c(tmp_path):
    from vofod_b200 import abi
    structs = {"vofod_pt": abi.Pt, "vofod_vox": abi.Vox, "vofod_xyzi": abi.Xyzi, "vofod_pose": abi.Pose, "vofod_params": abi.Params,
               "vofod_map_info": abi.MapInfo, "vofod_cluster_info": abi.ClusterInfo, "vofod_detection": abi.Detection,
               "vofod_schedule": abi.Schedule, "vofod_scan_result": abi.ScanResult}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, py in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in py._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, py in structs.items():
        assert int(got[cname]) == C.sizeof(py), cname
        for fname, _ in py._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(py, fname).offset, f"{cname}.{fname}"


def test_default_params_match_header_defaults():
    from vofod_b200 import abi, capi
    lib = capi.load_library()
    p = abi.Params()
    lib.vofod_default_params(C.byref(p))  # pure host function, no device needed
    q = abi.default_params()
    for f, _ in abi.Params._fields_:
        a, b = getattr(p, f), getattr(q, f)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), f
        else:
            assert a == b or f.startswith("_"), f


def test_no_cpu_fallback_without_device():
    """On a box without a GPU vofod_create must fail loudly (VOFOD_E_CUDA), never fall back to the CPU."""
    from vofod_b200 import abi, capi
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    with pytest.raises(capi.VofodError) as e:
        capi.Vofod(0)
    assert e.value.code == abi.VOFOD_E_CUDA


def test_product_does_not_import_oracle():
    """The product package must not reference oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "vofod_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "libvofod_oracle" not in text, os.path.join(dirpath, f)


def test_no_global_access_before_the_dependency_wait():
    """Kernels are chained by programmatic dependent launch: each must execute griddepcontrol.wait (SASS ACQBULK) before its
    first global memory access — checked on the SASS of the built library, because ptxas may hoist read-only loads."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_pdl_sass
    n, bad = check_pdl_sass.check()
    assert n >= 58 and not bad, bad


def test_option_constants_match_the_header():
    """abi.OPT_* mirror the VOFOD_OPT_* switches of include/vofod_cuda.h."""
    from vofod_b200 import abi
    text = open(os.path.join(ROOT, "include", "vofod_cuda.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+VOFOD_OPT_(\w+)\s+(\d+)", text)}
    assert len(set(defs.values())) == len(defs)                      # no two switches share a number
    mirrored = {k[4:]: v for k, v in vars(abi).items() if k.startswith("OPT_")}
    assert mirrored, "abi.py mirrors at least the switches the tests use"
    for name, value in mirrored.items():
        assert defs.get(name) == value, (name, value, defs.get(name))
