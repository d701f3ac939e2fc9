"""CPU: the C-ABI library loads and exports every symbol include/dockgpu.h declares.
No compute calls are made here (no GPU in this container)."""
import ctypes
import os
import re

from crypto_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'dockgpu.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dg_[a-z0-9_]+)\s*\(', src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(lib.EXPORTS)


def test_library_exports_every_symbol():
    assert os.path.exists(lib.LIB_PATH), 'libdockgpu.so missing: run python -m crypto_b200.build'
    so = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(so, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device dg_init must fail loudly (never a silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        return
    so = lib.load()
    rc = so.dg_init(ctypes.c_int32(0))
    assert rc < 0
    buf = ctypes.create_string_buffer(256)
    so.dg_last_error(buf, ctypes.c_size_t(256))
    assert buf.value
    # and a compute entry point refuses to run uninitialised
    out = ctypes.create_string_buffer(144)
    assert so.dg_msm_g1(ctypes.c_uint64(0), None, None, ctypes.c_size_t(0), out) == -4


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/."""
    pkg = os.path.join(ROOT, 'crypto_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'cpu_ref' not in text and 'libcpuref' not in text, f
                if f.endswith('.py'):
                    assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f


def test_rust_ffi_mirrors_the_header():
    """bindings/rust/dock_gpu/src/ffi.rs is generated from include/dockgpu.h (tools/gen_rust_ffi.py): every entry point
    is declared, the committed file is the generator's current output, and the safe wrappers in lib.rs only call
    functions that exist."""
    from tools import gen_rust_ffi
    text, names = gen_rust_ffi.render()
    assert sorted(names) == header_symbols()
    assert open(gen_rust_ffi.OUT).read() == text, 'run python tools/gen_rust_ffi.py'
    glue = open(os.path.join(ROOT, 'bindings', 'rust', 'dock_gpu', 'src', 'lib.rs')).read()
    used = set(re.findall(r'ffi::(dg_[a-z0-9_]+)', glue))
    assert used and used <= set(names), sorted(used - set(names))
