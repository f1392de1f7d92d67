"""include/tfhe_b200.h consumed from plain C (gcc -std=c99 -pedantic) and linked against libtfhe_b200.so: the header is what a
host-language binding (the Rust extern block of INTEGRATION.md, cgo, JNI ...) would mirror, so it must stand on its own."""
import os
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_header_is_self_sufficient_c_and_library_links(tmp_path):
    import fhe_string_bounty_b200 as F
    so = F.build_native()
    exe = tmp_path / "link_test"
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "tests/c_link/link_test.c"),
           "-o", str(exe), "-L", str(so.parent), "-ltfhe_b200", f"-Wl,-rpath,{so.parent}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "")))
    assert run.returncode == 0, run.stdout + run.stderr
    assert "string_eq(8, 8): 64 inputs, 36 PBS in 3 levels" in run.stdout


def test_every_declared_symbol_is_exported_and_bound():
    """every function include/tfhe_b200.h declares is exported by the library, typed by the ctypes loader and named in the C link test"""
    import fhe_string_bounty_b200 as F
    from fhe_string_bounty_b200 import _native
    F.build_native()
    lib = F.load_native()
    header = (ROOT / "include/tfhe_b200.h").read_text()
    declared = set(re.findall(r"\b(tfhe_b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"tfhe_b200_ctx", "tfhe_b200_program", "tfhe_b200_exchange"}
    assert len(declared) >= 40
    link_src = (ROOT / "tests/c_link/link_test.c").read_text()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared but not exported"
        assert name in _native.EXPORTS, f"{name} is not typed in _native.EXPORTS"
        assert name in link_src, f"{name} is not referenced by tests/c_link/link_test.c"
