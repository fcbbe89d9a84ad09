"""The C-ABI library loads and exports every symbol include/clq.h declares; host-only entry points behave.
No compute is attempted here (no GPU in the build container)."""
import ctypes as C
import os
import re

import pytest

from clique_b200 import _lib as L
from clique_b200 import AffineScoring, ClqError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "clq.h")) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(clq_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = L.load_library()
    names = declared_symbols()
    assert len(names) >= 19
    assert sorted(L.SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.clq_version() == 100


def test_header_structs_match_ctypes():
    assert C.sizeof(L.AffineInt) == 44 and C.sizeof(L.Result) == 28
    assert C.sizeof(L.Limits) == 48 and C.sizeof(L.Stats) == 56


def test_affine_from_f64():
    s = AffineScoring.align_reads_default().to_int()
    assert (s.scale, s.match, s.mismatch, s.special, s.oe_in, s.e_in, s.oe_fin, s.e_fin, s.b0, s.b1, s.max_neg) == \
           (1, 10, -9, 9, -22, -2, -22, -2, -20, -2, -100000)
    s = AffineScoring.default_dna().to_int()
    assert (s.scale, s.match, s.oe_in, s.e_in, s.oe_fin, s.e_fin, s.b0, s.b1) == (4, 20, -42, -2, -41, -1, -20, -1)
    s = AffineScoring.merger_default().to_int()
    assert (s.scale, s.oe_fin, s.e_fin, s.b0, s.b1, s.max_neg) == (4, -61, -1, -15, -1, -400000)
    with pytest.raises(ClqError) as e:
        AffineScoring(1.0, -1.0, 1.0, -1.0 / 3.0, -1.0, 1.0).to_int()
    assert e.value.code == L.SCORING_NOT_REPRESENTABLE
    with pytest.raises(ClqError):  # distance_dna has gap_open = 0: no exact direction encoding
        AffineScoring(0.0, -1.0, -1.0, 0.0, -1.0, 1.0).to_int()


def test_oracle_and_library_agree_on_integer_scoring():
    import _oracle as O
    for sc in [(10.0, -9.0, 9.0, -20.0, -2.0, 1.0), (5.0, -4.0, 4.0, -10.0, -0.5, 0.5), (10.0, -5.0, 8.0, -15.0, -1.0, 0.25)]:
        a = AffineScoring(*sc).to_int()
        rc, b = O.affine_int(sc)
        assert rc == 0
        for f, _ in L.AffineInt._fields_:
            assert getattr(a, f) == getattr(b, f), f


def test_no_gpu_fails_loudly():
    lib = L.load_library()
    if lib.clq_device_count() > 0:
        pytest.skip("a GPU is present")
    lim = L.Limits(16, 1 << 20, 1 << 16, 16, 1 << 20, 1 << 10, 1)
    ctx = C.c_void_p()
    assert lib.clq_ctx_create(0, C.byref(lim), C.byref(ctx)) == L.E_CUDA
    from clique_b200 import Aligner
    with pytest.raises(ClqError):
        Aligner(device=0, max_reads=16)


def test_strerror():
    lib = L.load_library()
    assert lib.clq_strerror(L.TRACEBACK_DIVERGED).decode().startswith("traceback diverged")
    assert lib.clq_strerror(-2).decode() == "CUDA error"


def _c_header_model():
    """functions (name -> parameter count), structs (name -> field names in order) and #define constants of include/clq.h"""
    src = open(os.path.join(ROOT, "include", "clq.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    funcs = {}
    for m in re.finditer(r"\b(clq_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        args = m.group(2).strip()
        funcs[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    structs = {}
    for m in re.finditer(r"typedef struct \{(.*?)\}\s*(clq_[a-z_]+_t)\s*;", src, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if decl:
                fields += [f.strip().lstrip("*") for f in decl.split(None, 1)[1].split(",")]
        structs[m.group(2)] = fields
    consts = {}
    for m in re.finditer(r"#define (CLQ_[A-Z0-9_]+)[ \t]+(\S.*)", src):
        expr = m.group(2).strip().replace("u", "")
        consts[m.group(1)] = int(eval(expr, {"__builtins__": {}}))
    return funcs, structs, consts


def test_rust_sys_crate_matches_header():
    """rust/clq-sys is source only (no Rust toolchain in this image): keep its extern block, #[repr(C)] structs and constants
    in step with include/clq.h mechanically."""
    rs = open(os.path.join(ROOT, "rust", "clq-sys", "src", "lib.rs")).read()
    rs = re.sub(r"//.*", "", rs)
    funcs, structs, consts = _c_header_model()
    assert len(funcs) == len(L.SYMBOLS)
    ext = rs[rs.index('extern "C" {'):]
    rust_funcs = {}
    for m in re.finditer(r"pub fn (clq_[a-z0-9_]+)\s*\((.*?)\)\s*(->\s*[^;]+)?;", ext, flags=re.S):
        args = m.group(2).strip()
        rust_funcs[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    assert rust_funcs == funcs
    for name, fields in structs.items():
        m = re.search(r"pub struct %s \{(.*?)\}" % name, rs, flags=re.S)
        assert m, name
        rust_fields = [f.strip().split(":")[0].replace("pub ", "").strip() for f in m.group(1).split(",") if f.strip()]
        assert [f.rstrip("_") for f in rust_fields] == fields, name      # `match` is a Rust keyword: match_
    rust_consts = {m.group(1): m.group(3) for m in re.finditer(r"pub const (CLQ_[A-Z0-9_]+): (u32|i32) = ([^;]+);", rs)}
    for name, value in consts.items():
        assert name in rust_consts, name
        assert int(eval(rust_consts[name], {"__builtins__": {}})) == value, name
