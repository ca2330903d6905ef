"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every function that
include/tahoe_b200.h declares; the harness binding lists the same set; compute calls fail loudly without a device."""
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(REPO, "include", "tahoe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb2_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared()
    assert len(names) >= 60
    for must in ("tb2_form_internal_force", "tb2_form_stiffness", "tb2_matrix_pcg", "tb2_explicit_run", "tb2_comm_sum_interface"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from tahoe_b200 import capi
    lib = capi.lib()
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == _declared()
    assert b"sm_100a" in lib.tb2_version()


def test_no_cpu_fallback_without_a_device():
    """on a box without a GPU every compute entry point must fail with TB2_ERR_CUDA, never compute on the host"""
    from tahoe_b200 import capi, mesh as tmesh
    try:
        n = capi.device_count()
    except capi.Tb2Error:
        n = 0
    if n > 0:
        pytest.skip("a CUDA device is present")
    X, conn, _ = tmesh.structured_cube(2)
    with pytest.raises(capi.Tb2Error) as e:
        capi.Mesh(X, conn)
    assert e.value.code == 3  # TB2_ERR_CUDA


def test_header_is_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: include/tahoe_b200.h must compile as C99 (no C++ types, no CUDA headers) and every
    declared function must link against the shared library"""
    import subprocess
    src = tmp_path / "abi.c"
    names = _declared()
    body = "\n".join("    p[%d] = (fn_t)%s;" % (i, n) for i, n in enumerate(names))
    src.write_text('#include "tahoe_b200.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n    fn_t p[%d];\n%s\n'
                   '    printf("%%s %%d\\n", tb2_version(), (int)(sizeof p / sizeof p[0]));\n    return p[0] == 0;\n}\n' % (len(names), body))
    exe = tmp_path / "abi"
    lib_dir = os.path.join(REPO, "tahoe_b200", "lib")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-ltahoe_b200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and "sm_100a" in out.stdout and str(len(names)) in out.stdout
