"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/milagro_bls_b200.h declares; the product fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build_cuda()
    return g.LIB


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "milagro_bls_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b3_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built)
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_python_binding_covers_header(built):
    from milagro_bls_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()
    _lib.lib()


def test_rust_shim_binds_exactly_the_header():
    """shim/src/b200/ffi.rs (the reference-side binding a maintainer adds; not compiled here: no rustc) declares exactly the
    functions of the header, each with the header's number of parameters."""
    rs = open(os.path.join(ROOT, "shim", "src", "b200", "ffi.rs")).read()
    rs_fns = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (b3_[a-z0-9_]+)\(([^)]*)\)", rs, flags=re.S)}
    assert sorted(rs_fns) == _declared_symbols()
    txt = open(os.path.join(ROOT, "include", "milagro_bls_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for name, args in re.findall(r"\b(b3_[a-z0-9_]+)\s*\(([^)]*)\)", txt, flags=re.S):
        n_c = 0 if args.strip() in ("", "void") else args.count(",") + 1
        n_rs = 0 if not rs_fns[name].strip() else rs_fns[name].count(",") + 1
        assert n_c == n_rs, f"{name}: {n_c} parameters in the header, {n_rs} in ffi.rs"
    # the method bodies only call functions that ffi.rs declares
    for f in ("keys_b200.rs", "signature_b200.rs", "aggregates_b200.rs", os.path.join("b200", "ctx.rs")):
        for used in set(re.findall(r"\b(b3_[a-z0-9_]+)\s*\(", open(os.path.join(ROOT, "shim", "src", f)).read())):
            assert used in rs_fns, f"{f} calls undeclared {used}"


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import milagro_bls_b200 as mb
    with pytest.raises(RuntimeError):
        mb.Engine(0)
    with pytest.raises(RuntimeError):
        mb.Signature.from_bytes(bytes([0xc0]) + bytes(95))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "milagro_bls_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/ is", ""), f


def test_rng_draw_rule():
    from milagro_bls_b200.rng import SeededRng, draw_scalar
    from oracle import bls_oracle as O
    a, b = SeededRng(b"k"), O.SeededRng(b"k")
    for _ in range(50):
        assert draw_scalar(a) == O.draw_scalar(b.fill)

    class Fixed:
        def __init__(self, chunks):
            self.c = list(chunks)

        def fill(self, n):
            return self.c.pop(0)

    assert draw_scalar(Fixed([bytes(8), (1 << 63).to_bytes(8, "big"), b"\xff" * 8])) == 1
