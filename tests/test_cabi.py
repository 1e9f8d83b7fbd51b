"""CPU: the C-ABI library loads and exports every symbol include/parry_b200.h declares; the ctypes table matches the
header; without a CUDA device the product path fails loudly (no CPU fallback, nothing routes through oracle/)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "parry_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(pb2_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    names = header_functions()
    for must in ("pb2_ctx_create", "pb2_bvh_build", "pb2_bvh_refit", "pb2_bvh_self_pairs", "pb2_bvh_intersect_aabbs",
                 "pb2_bvh_leaf_pairs", "pb2_trimesh_create", "pb2_trimesh_cast_rays", "pb2_bvh_cast_rays_shapes",
                 "pb2_contact_batch", "pb2_contact_batch_compact", "pb2_shapes_compute_aabbs"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from parry_b200 import _ffi, build
    if not os.path.exists(_ffi.LIB_PATH):
        build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (pb2_[a-z0-9_]+)", out))
    declared = header_functions()
    missing = [n for n in declared if n not in exported]
    assert not missing, missing
    # the ctypes table binds exactly the declared functions
    assert sorted(_ffi.SIGNATURES) == declared
    lib = _ffi.lib()
    assert lib.pb2_version() >= 100


def test_no_torch_or_oracle_in_the_boundary():
    from parry_b200 import _ffi
    out = subprocess.run(["ldd", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out
    # product sources never reference the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "parry_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "harness" not in txt and "oracle/" not in txt, f


def test_fails_loudly_without_a_gpu():
    import parry_b200
    lib = parry_b200.lib()
    if lib.pb2_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(parry_b200.Pb2Error):
        parry_b200.Context(0)


def test_rust_sys_crate_matches_the_header():
    """rust-shim/parry-b200-sys/src/lib.rs is generated from include/parry_b200.h (rust-shim/gen_sys.py): the committed file must be
    the generator's output, bind every prototype with the same number of arguments, and the shim crate may only call what exists."""
    import importlib.util
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_sys", os.path.join(root, "rust-shim", "gen_sys.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text, funcs = gen.generate()
    committed = open(os.path.join(root, "rust-shim", "parry-b200-sys", "src", "lib.rs")).read()
    assert committed == text, "run `python rust-shim/gen_sys.py` after changing include/parry_b200.h"
    declared = header_functions()
    bound = {name: params for name, params, _ in funcs}
    assert sorted(bound) == declared
    from parry_b200 import _ffi
    for name, params in bound.items():          # same arity as the ctypes table the tests call through
        assert len(params) == len(_ffi.SIGNATURES[name][1]), name
    for line in re.findall(r"pub fn (pb2_\w+)\(([^)]*)\)", committed):
        assert "*const" in line[1] or "*mut" in line[1] or line[1] == "" or ":" in line[1]
    shim = open(os.path.join(root, "rust-shim", "parry-b200", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(pb2_[a-z0-9_]+)\(", shim))
    assert len(used) > 25 and not [u for u in used if u not in bound], [u for u in used if u not in bound]
    assert "/* pb2_" not in shim and "todo!" not in shim and "unimplemented!" not in shim      # bodies, not comments
    # every trait method of QueryDispatcher is implemented (query_dispatcher.rs:408-506)
    for m in ("fn intersection_test", "fn distance", "fn contact", "fn closest_points", "fn cast_shapes", "fn cast_shapes_nonlinear"):
        assert m in shim, m


def test_cpp_mirror_header_compiles(tmp_path):
    """include/parry_b200.hpp (the header-only C++ host mirror) compiles against include/parry_b200.h with a plain host compiler."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hpp_check.cpp"
    src.write_text('#include "parry_b200.hpp"\nint main() { return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-I", os.path.join(root, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
