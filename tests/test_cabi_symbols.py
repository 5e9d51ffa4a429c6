"""CPU: libsrst.so loads without a GPU and exports every entry point include/srst.h declares."""
import ctypes
import os
import re

from tests.helpers import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "srst.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srst_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = _declared()
    for must in ("srst_st_forward", "srst_st_backward", "srst_bb_forward", "srst_bb_backward",
                 "srst_st_workspace_bytes", "srst_bb_workspace_bytes", "srst_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from srgan_st_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in srst.h but not exported: {missing}"


def test_binding_table_matches_header():
    from srgan_st_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == _declared()


def test_no_compute_entry_points_that_need_no_gpu():
    """Pure host queries work on the CPU box (no kernel is launched)."""
    from srgan_st_b200 import _cabi
    lib = _cabi.lib()
    assert lib.srst_version() // 100 == _cabi.ABI_MAJOR == 2
    assert lib.srst_st_supported(2, 8) == 1
    assert lib.srst_st_supported(4, 12) == 1 and lib.srst_st_supported(1, 3) == 1   # padded radius classes
    assert lib.srst_st_supported(5, 8) == 2 and lib.srst_st_supported(2, 13) == 2   # generic-radius path
    assert lib.srst_st_supported(65, 8) == 0 and lib.srst_st_supported(2, 65) == 0
    assert lib.srst_st_workspace_bytes_r(16, 96, 96, 2, 8) == lib.srst_st_workspace_bytes(16, 96, 96)
    assert lib.srst_st_workspace_bytes_r(16, 96, 96, 6, 16) > 11 * 16 * 96 * 96 * 4
    assert lib.srst_st_backward_workspace_bytes(16, 96, 96, 2, 8) == 0
    assert lib.srst_st_backward_workspace_bytes(16, 96, 96, 6, 16) >= 5 * 16 * 96 * 96 * 4
    assert lib.srst_st_workspace_bytes(16, 96, 96) >= 16 * 3 * 2 * 4
    assert lib.srst_st_workspace_bytes(0, 96, 96) == 0
    assert b"workspace" in lib.srst_error_string(-3)
    assert lib.srst_st_ixy_floats(2, 5, 8) == 2 * 2 * 3 * 8 * 2       # [B][2][ceil(H/2)][W][2]
    assert lib.srst_st_num_cfgs(0) >= 2 and lib.srst_st_num_cfgs(1) >= 2
    assert lib.srst_st_force_cfg(99, 0) == -1 and lib.srst_st_force_cfg(-1, -1) == 0
    # argument validation happens before any CUDA call
    assert lib.srst_st_forward(None, None, 1, 8, 8, None, None, 2, None, 8, 1, 1e-12, None, None, None,
                               None, None, None, 0, None) == -1
    assert lib.srst_st_backward(None, None, None, 1, 8, 8, None, None, 2, None, 8, None, None) == -1
    assert lib.srst_st_backward_ws(None, None, None, 1, 8, 8, None, None, 6, None, 16, None, None, 0, None) == -1


def test_integration_md_binding_snippet_matches_the_abi():
    """INTEGRATION.md section 2 shows a maintainer the ctypes stub; its argtypes lists must be the ones the
    package binds (round 1 shipped a stale 18-argument srst_st_forward there: copying it corrupted the call)."""
    from srgan_st_b200 import _cabi
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    env = {"ctypes": ctypes, "vp": ctypes.c_void_p, "fp": ctypes.POINTER(ctypes.c_float)}
    found = 0
    for name, body in re.findall(r"lib\.(srst_\w+)\.argtypes = (\[.*?\])\s*(?:#[^\n]*)?\n(?=\S)", md, flags=re.S):
        body = re.sub(r"#[^\n]*", "", body)
        argtypes = eval(body, env)  # noqa: S307 - a literal list of ctypes names from our own document
        assert argtypes == _cabi.SIGNATURES[name][1], f"INTEGRATION.md: {name} argtypes differ from _cabi.SIGNATURES"
        found += 1
    assert found >= 2
    m = re.search(r"srst_version\(\) // 100 == (\d+)", md)
    assert m and int(m.group(1)) == _cabi.ABI_MAJOR
