"""include/cia.h parsed as text against the ctypes table of the package (cell_image_analysis_b200._lib.SIGNATURES and
the Structure classes): parameter count, width and kind (integer / floating point / pointer) of every entry point, the
return types, and the layout of every struct.  A mismatch here is a silent wrong-argument bug on the GPU box -- no GPU
needed, nothing is called."""
import ctypes as C
import os
import re

from cell_image_analysis_b200 import _lib
from cell_image_analysis_b200.stardist import SegConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "cia.h")).read()
HDR = re.sub(r"/\*.*?\*/", " ", HDR, flags=re.S)
HDR = re.sub(r"//[^\n]*", " ", HDR)

SCALARS = {"int": ("i", 4), "int32_t": ("i", 4), "uint32_t": ("i", 4), "int64_t": ("i", 8), "uint64_t": ("i", 8),
           "size_t": ("i", 8), "long long": ("i", 8), "float": ("f", 4), "double": ("f", 8), "int8_t": ("i", 1),
           "uint16_t": ("i", 2), "int16_t": ("i", 2), "cia_handle": ("p", 8)}


def _kind_c(decl):
    """('p' | 'i' | 'f', bytes) of one C parameter / return declaration"""
    d = decl.replace("const", " ").strip()
    if "*" in d or "[" in d:
        return ("p", 8)
    d = re.sub(r"\s+", " ", d)
    for t in sorted(SCALARS, key=len, reverse=True):
        if d == t or d.startswith(t + " "):
            return SCALARS[t]
    raise AssertionError(f"unparsed C type: {decl!r}")


def _kind_ctypes(t):
    if t is None:
        return None
    if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or issubclass(t, C._Pointer):
        return ("p", 8)
    if t in (C.c_float, C.c_double):
        return ("f", C.sizeof(t))
    return ("i", C.sizeof(t))


def _declarations():
    text = re.sub(r"typedef struct \w+\s*\{.*?\}\s*\w+\s*;", ";", HDR, flags=re.S)     # struct bodies hold ';'
    text = re.sub(r'extern\s+"C"\s*\{', ";", text)
    text = re.sub(r"^\s*#[^\n]*", "", text, flags=re.M)
    out = {}
    for stmt in text.split(";"):
        m = re.match(r"\s*([A-Za-z_][\w\s\*]*?)\b(cia_[a-z0-9_]+)\s*\((.*)\)\s*$", stmt, flags=re.S)
        if not m:
            continue
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [re.sub(r"\s+", " ", p.strip()) for p in params.split(",")]
        out[name] = (ret, plist)
    return out


def test_every_prototype_matches_the_ctypes_table():
    decls = _declarations()
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, (ret, params) in sorted(decls.items()):
        res, args = _lib.SIGNATURES[name]
        assert len(params) == len(args), (name, len(params), len(args))
        for i, (p, a) in enumerate(zip(params, args)):
            assert _kind_c(p) == _kind_ctypes(a), (name, i, p, a)
        if ret == "void":
            assert res is None, name
        else:
            assert _kind_c(ret) == _kind_ctypes(res), (name, ret, res)


def _struct_fields(name):
    m = re.search(rf"typedef struct {name}\s*\{{(.*?)\}}\s*{name}\s*;", HDR, flags=re.S)
    assert m, name
    fields = []
    for stmt in m.group(1).split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        t, names = re.match(r"((?:const\s+)?[A-Za-z_]\w*(?:\s+long)?\s*\**)\s*(.*)", stmt, flags=re.S).groups()
        for n in names.split(","):
            n = n.strip()
            ptr = n.startswith("*") or "*" in t
            fields.append((n.lstrip("* "), ("p", 8) if ptr else _kind_c(t)))
    return fields


def test_every_struct_layout_matches_the_ctypes_classes():
    for cname, cls in (("cia_region", _lib.Region), ("cia_cell", _lib.Cell), ("cia_params", _lib.Params),
                       ("cia_scores", _lib.Scores), ("cia_seg_config", SegConfig)):
        hdr = _struct_fields(cname)
        got = [(n, _kind_ctypes(t)) for n, t in cls._fields_]
        assert hdr == got, (cname, hdr, got)
        # natural alignment of the header's fields gives the same size the ctypes class has
        off = 0
        for _, (_, size) in hdr:
            off = (off + size - 1) // size * size + size
        align = max(s for _, (_, s) in hdr)
        assert (off + align - 1) // align * align == C.sizeof(cls), cname


def test_every_entry_point_answers_a_null_handle_and_null_buffers_with_an_error_code():
    """error behaviour at the boundary: CIA_E_ARG (-2) for a null handle on every handle-taking entry point, and the
    host-only functions refuse null buffers -- nothing dereferences, nothing needs a GPU"""
    lib = _lib.load()
    raw = open(os.path.join(ROOT, "include", "cia.h")).read()
    swept = 0
    for name, (res, args) in _lib.SIGNATURES.items():
        if not re.search(rf"\b{name}\s*\(\s*cia_handle\s+h\b", raw):
            continue
        vals = [0.0 if a in (C.c_double, C.c_float) else 0 if _kind_ctypes(a)[0] == "i" else None for a in args]
        r = getattr(lib, name)(*vals)
        swept += 1
        if name == "cia_last_error":
            assert r == b"null handle"
        elif name == "cia_launch_count":
            assert r == 0
        else:
            assert r == _lib.CIA_E_ARG, (name, r)
    assert swept >= 35
    assert lib.cia_rle_encode_fields(None, 1, 8, 8, None, 0, None, None, 1) == _lib.CIA_E_ARG
    assert lib.cia_rle_encode_pack_fields(None, None, 1, 8, 8, None, 0, None, None, 4, None, 0, None, 1) == _lib.CIA_E_ARG
    assert lib.cia_tiff_lzw_decode(None, 0, None, 0) == -1 and lib.cia_tiff_packbits_decode(None, 0, None, 0) == -1
    assert lib.cia_host_read_probe(None, 1 << 20, 1, 1) == 0.0
    assert lib.cia_create(0, None) == _lib.CIA_E_ARG
