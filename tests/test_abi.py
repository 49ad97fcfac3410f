"""CPU: the C-ABI library loads, exports every symbol include/rst_align.h declares, and its POD
layouts match the ctypes mirrors. No compute call is made without a GPU."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest

from conftest import has_gpu, ROOT
from realsensetracker_b200 import _native as N


def declared_functions():
    hdr = (ROOT / "include" / "rst_align.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rst_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = N.align_lib()
    names = declared_functions()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(N.ALIGN_SYMBOLS) == names, "ALIGN_SYMBOLS is out of sync with the header"
    assert lib.rst_abi_version() == 2


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rst_align.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rst_params), sizeof(rst_stats),'
                   'sizeof(rst_frame), sizeof(rst_intrinsics), sizeof(rst_profile), offsetof(rst_stats, A),'
                   'offsetof(rst_params, robust_scale), offsetof(rst_frame, width));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(N.Params), C.sizeof(N.Stats), C.sizeof(N.Frame), C.sizeof(N.Intrinsics), C.sizeof(N.Profile),
            N.Stats.A.offset, N.Params.robust_scale.offset, N.Frame.width.offset]
    assert got == want


def test_default_params_match_the_oracle_defaults():
    from oracle import oracle as O
    p = N.Params()
    N.align_lib().rst_params_default(C.byref(p))
    q = O.default_params()
    for name, _ in N.Params._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b, name


@pytest.mark.skipif(has_gpu(), reason="checks the no-device error path")
def test_no_cpu_fallback_without_a_device():
    """The product path must fail loudly, not fall back, when there is no GPU."""
    from realsensetracker_b200 import Aligner, RstError
    with pytest.raises(RstError) as e:
        Aligner(640, 480, 2, 1)
    assert e.value.code in (N.RST_ERR_NO_DEVICE, N.RST_ERR_CUDA, N.RST_ERR_ARCH)


def test_null_and_bad_arguments_are_rejected_without_a_device():
    lib = N.align_lib()
    ctx = C.c_void_p()
    assert lib.rst_ctx_create(0, 8, 8, 2, 1, None, C.byref(ctx)) == N.RST_ERR_INVALID_ARG
    assert b"capacity" in lib.rst_last_create_error()
    assert lib.rst_sync(None) == N.RST_ERR_INVALID_ARG
    assert lib.rst_launch_count(None) == 0
    lib.rst_ctx_destroy(None)  # must be a no-op
    # the cloud entry points need a context as well: no context, no answer (and no CPU path behind them)
    cloud = N.Cloud(None, 0)
    buf = (C.c_float * 16)()
    n = C.c_int32(0)
    tree = C.c_void_p()
    assert lib.rst_cloud_centroid(None, C.byref(cloud), buf) == N.RST_ERR_INVALID_ARG
    assert lib.rst_cloud_extents(None, C.byref(cloud), buf, buf) == N.RST_ERR_INVALID_ARG
    assert lib.rst_orient_normals(None, C.byref(cloud), buf, buf) == N.RST_ERR_INVALID_ARG
    assert lib.rst_find_correspondences(None, C.byref(cloud), C.byref(cloud), 0.0, None, None) == N.RST_ERR_INVALID_ARG
    assert lib.rst_cloud_covariances(None, C.byref(cloud), 0, 0.0, buf) == N.RST_ERR_INVALID_ARG
    assert lib.rst_downsample_voxel(None, C.byref(cloud), 0.05, buf, C.byref(n)) == N.RST_ERR_INVALID_ARG
    assert lib.rst_remove_nans(None, C.byref(cloud), buf, C.byref(n)) == N.RST_ERR_INVALID_ARG
    assert lib.rst_tree_create(None, C.byref(cloud), 0.0, C.byref(tree)) == N.RST_ERR_INVALID_ARG and not tree.value
    assert lib.rst_tree_query(None, None, buf, 1, 1, None, None) == N.RST_ERR_INVALID_ARG
    assert lib.rst_tree_size(None) == 0
    lib.rst_tree_destroy(None)  # no-op
    assert lib.rst_gicp_minimize(None, C.byref(cloud), C.byref(cloud), buf, buf, None, 4, 0.5, buf, None) == N.RST_ERR_INVALID_ARG


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under realsensetracker_b200/ may reference it."""
    for p in (ROOT / "realsensetracker_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".c", ".cpp") and p.is_file():
            txt = p.read_text()
            assert "oracle" not in txt.lower(), f"{p} mentions the oracle"
