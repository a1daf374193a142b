"""The C-ABI library loads and exports every symbol include/deepmod_b200.h declares (no GPU calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "deepmod_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    fns = declared_functions()
    for must in ("dm_create", "dm_destroy", "dm_detect_batch", "dm_forward_windows", "dm_set_genome",
                 "dm_hist_nonzero", "dm_write_bed", "dm_hist_device_ptr", "dm_last_error"):
        assert must in fns


def test_library_builds_and_exports_every_declared_symbol():
    from deepmod_b200 import build, capi
    build.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), "%s declared in the header but not exported" % name
    # and the ctypes table binds exactly the declared surface
    assert sorted(capi.SIGNATURES) == declared_functions()
    assert capi.load_library().dm_version() >= 100


def test_no_cpu_fallback_without_a_device():
    """Without a GPU the product path must fail loudly, not compute on the host."""
    from deepmod_b200 import capi, checkpoint
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(capi.DeepModError) as e:
        capi.Context(checkpoint.random_model())
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deepmod_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """include/deepmod_b200.h compiles as C99 with -Wall -Werror, and a C program linked against the library gets status
    codes through the boundary (examples/c_abi_probe.c).  Without a GPU its dm_create fails with DM_ERR_CUDA: exit 0."""
    import shutil
    import subprocess
    from deepmod_b200 import build, capi
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    build.build()
    exe = str(tmp_path / "probe")
    lib_dir = os.path.dirname(capi.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "c_abi_probe.c"), "-L", lib_dir, "-ldeepmod_b200", "-Wl,-rpath," + lib_dir,
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dm_version" in r.stdout
