"""The C-ABI library loads on a CPU-only box and exports every symbol include/qpg.h declares
(no compute calls here)."""
import os
import re

from qpgesture_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "qpg.h")).read()
    declared = set(re.findall(r"\b(qpg_[a-z0-9_]+)\s*\(", header))
    declared -= {"qpg_pair_t", "qpg_conv_desc_t", "qpg_conv_tc_desc_t"}
    assert len(declared) >= 18
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/qpg.h but not exported"
    assert set(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ declared
    assert lib.qpg_version() >= 100
    assert lib.qpg_packed_bytes(13312, 6144) == 13312 * 6144 * 4
    assert lib.qpg_packed_bytes(9, 130) == 2 * 2 * 4096


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    assert lib.qpg_table_init(None, -1, None) < 0
    assert b"n_entries" in lib.qpg_last_error()
    assert lib.qpg_cand_cosine_minbycode(None, None, None, -5, 128, 0, None, 1, None, 0, None) < 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "qpgesture_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_header_is_plain_c_and_links(tmp_path):
    """include/qpg.h compiles as strict C99 and a plain-C program links against the library (examples/scan_from_c.c)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    _lib.load()
    exe = str(tmp_path / "scan_from_c")
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "scan_from_c.c"), "-L" + os.path.join(ROOT, "qpgesture_b200"), "-lqpg_sm100",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(ROOT, "qpgesture_b200"), "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert os.path.isfile(exe)
