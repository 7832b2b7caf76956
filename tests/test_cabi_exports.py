"""CPU: the C-ABI library loads and exports every symbol include/mma_b200.h declares
(no compute calls without a GPU), argument validation paths return error codes."""
import ctypes
import os
import re

from mma_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|const char \*)\s*(mma\w+|mmconv\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 12, names
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/mma_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding table out of sync with the header"
    assert _lib.lib().mma_b200_version() >= 100


def test_binding_table_matches_every_declaration():
    """Parameter by parameter: the C type of every declared argument against the ctypes type it is bound with."""
    text = open(os.path.join(ROOT, "include", "mma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    checked = 0
    for name, args in re.findall(r"\b(?:int|const char \*)\s*(mma\w+|mmconv\w+)\s*\((.*?)\)\s*;", text, flags=re.S):
        params = [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]
        bound = _lib._SIGS[name][0]
        assert len(params) == len(bound), f"{name}: {len(params)} declared, {len(bound)} bound"
        for prm, ct in zip(params, bound):
            prm = re.sub(r"^const\s+", "", prm)
            if "*" in prm:
                want = (ctypes.c_void_p, ctypes.c_char_p)
                ok = ct in want or (hasattr(ct, "_type_"))          # POINTER(c_size_t) for out-parameters
            elif prm.startswith("uint64_t"):
                ok = ct is ctypes.c_uint64
            elif prm.startswith("int64_t"):
                ok = ct is ctypes.c_int64
            elif prm.startswith("size_t"):
                ok = ct is ctypes.c_size_t
            elif prm.startswith("uint32_t"):
                ok = ct is ctypes.c_uint32
            elif prm.startswith("float"):
                ok = ct is ctypes.c_float
            elif prm.startswith("mma_stream_t"):
                ok = ct is ctypes.c_void_p
            else:
                ok = ct is ctypes.c_int and re.match(r"(int|int32_t)\b", prm) is not None
            assert ok, f"{name}: parameter '{prm}' bound as {ct}"
            checked += 1
    assert checked > 250


def test_argument_validation_without_gpu():
    l = _lib.lib()
    n = ctypes.c_size_t(0)
    assert l.mma_csr_build_workspace_bytes(-1, 4, ctypes.byref(n)) == _lib.ERR_INVALID
    assert l.mma_csr_build_workspace_bytes(2**31, 4, ctypes.byref(n)) == _lib.ERR_UNSUPPORTED
    ak = _lib.i32_array([0]); sk = _lib.i32_array([0])
    # all data pointers NULL -> invalid, before any CUDA call
    rc = l.mmconv_aggregate_fwd(None, None, None, None, 0, None, None, 0, None, 0, None, 0, None, None, 0, None, 4, 0, None, 0, None, 0, None, 0, None, 0,
                                0.0, 0, None, 1, 4, 1, ak, 1, sk, None, 0, None, 0, None, None, None, None, 0, 0, 0, None)
    assert rc == _lib.ERR_INVALID
    assert l.mma_dropout_keep_scale_rows(None, None, None, 0, 4, 4, 0.5, 0, 4, None, 4, None) == _lib.ERR_INVALID
    assert l.mma_segment_sum_rows(None, None, None, 4, None, 0, 4, None, 0, None) == _lib.ERR_INVALID


def test_no_oracle_import_in_product_path():
    pkg = os.path.join(ROOT, "mma_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "from oracle" not in src and "import oracle" not in src, f


def test_integration_stub_matches_the_abi():
    """The ctypes stub a maintainer of the reference would paste must carry the argument list
    of include/mma_b200.h as bound in mma_b200/_lib.py -- same types, same count, in the declaration and in the call."""
    import ctypes as C
    import os
    import re
    from mma_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "INTEGRATION.md")).read()
    body = re.search(r"lib\.mmconv_aggregate_fwd\.argtypes = \[(.*?)\]\s*#[^\n]*\nKIND", txt, re.S).group(1)
    toks = [t.strip() for t in re.sub(r"#[^\n]*", "", body).replace("\n", " ").split(",") if t.strip()]
    names = {"vp": C.c_void_p, "i64": C.c_int64, "i32": C.c_int, "f32": C.c_float, "u64": C.c_uint64}
    assert [names[t] for t in toks] == list(_lib._SIGS["mmconv_aggregate_fwd"][0])
    call = re.search(r"rc = lib\.mmconv_aggregate_fwd\((.*?)\)\n    assert rc == 0", txt, re.S).group(1)
    call = re.sub(r"#[^\n]*", "", call)
    depth, n = 0, 1
    for ch in call:
        depth += ch in "([" 
        depth -= ch in ")]"
        n += (ch == "," and depth == 0)
    assert n == len(toks)
    # the header declares exactly the parameters that are bound
    hdr = open(os.path.join(root, "include", "mma_b200.h")).read()
    decl = re.search(r"\nint mmconv_aggregate_fwd\((.*?)\);", hdr, re.S).group(1)
    assert len([a for a in decl.split(",") if a.strip()]) == len(toks)



def test_k1_args_struct_layout_matches_the_header(tmp_path):
    """mma_k1_args_t as gcc lays it out from include/mma_b200.h against the ctypes Structure that binds it: size and
    the offset of every field (the versioned block is what a hand-written binding uses instead of 47 positional
    arguments, so its layout IS the ABI)."""
    import subprocess
    fields = [f[0] for f in _lib.K1Args._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "mma_b200.h")}"',
             'int main(void) {', '  printf("%zu\\n", sizeof(mma_k1_args_t));']
    lines += [f'  printf("%zu\\n", offsetof(mma_k1_args_t, {f}));' for f in fields]
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-o", str(exe), str(src)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(_lib.K1Args)
    for f, off in zip(fields, out[1:]):
        assert getattr(_lib.K1Args, f).offset == off, f


def test_k1_args_versioning_without_gpu():
    l = _lib.lib()
    a = _lib.K1Args()
    assert l.mmconv_aggregate_fwd_args(None, None) == _lib.ERR_INVALID
    a.struct_size = 16                                       # shorter than the first published layout
    assert l.mmconv_aggregate_fwd_args(ctypes.byref(a), None) == _lib.ERR_INVALID
    a.struct_size = ctypes.sizeof(_lib.K1Args)              # well-formed but empty: rejected by the argument checks
    assert l.mmconv_aggregate_fwd_args(ctypes.byref(a), None) == _lib.ERR_INVALID
    assert l.mmconv_aggregate_bwd_dst_args(ctypes.byref(a), None) == _lib.ERR_INVALID

    class Newer(ctypes.Structure):                           # a caller built against a later header
        _fields_ = [("base", _lib.K1Args), ("new_field", ctypes.c_int64)]
    n = Newer()
    n.base.struct_size = ctypes.sizeof(Newer)
    fn = l.mmconv_aggregate_fwd_args
    saved = fn.argtypes
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    try:
        assert fn(ctypes.addressof(n), None) == _lib.ERR_INVALID        # unknown field unset: accepted, then the usual checks
        n.new_field = 7
        assert fn(ctypes.addressof(n), None) == _lib.ERR_UNSUPPORTED    # unknown field SET: this library cannot honour it
    finally:
        fn.argtypes = saved
