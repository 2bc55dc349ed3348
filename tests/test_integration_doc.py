"""INTEGRATION.md is the binding a maintainer of the reference would copy: its ctypes stubs must agree with
include/cia.h as bound by the package (_lib.SIGNATURES / the Structure classes), every entry point it names must
be exported by libcia.so, and its Python blocks must at least compile -- no GPU needed (nothing is called)."""
import ctypes as C
import os
import re

import numpy as np

from cell_image_analysis_b200 import _lib
from cell_image_analysis_b200.stardist import SegConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOC = open(os.path.join(ROOT, "INTEGRATION.md")).read()
BLOCKS = re.findall(r"```python\n(.*?)```", DOC, flags=re.S)


def _class_source(name):
    for b in BLOCKS:
        m = re.search(rf"^class {name}\(C\.Structure\):.*?(?=^\S)", b, flags=re.S | re.M)
        if m:
            return m.group(0)
    raise AssertionError(f"no ctypes class {name} in INTEGRATION.md")


def test_python_blocks_compile():
    assert len(BLOCKS) >= 4
    for i, b in enumerate(BLOCKS):
        compile(b, f"INTEGRATION.md block {i}", "exec")


def test_every_named_entry_point_is_exported():
    lib = C.CDLL(_lib.LIB_PATH)
    names = set(re.findall(r"\bcia_[a-z0-9_]+\b", DOC)) - {"cia_params", "cia_scores", "cia_seg_config", "cia_cell", "cia_region",
                                                         "cia_handle", "cia_ctx"}
    assert len(names) > 25
    for n in sorted(names):
        assert n in _lib.SIGNATURES, f"{n} is named in INTEGRATION.md but not declared in include/cia.h"
        assert hasattr(lib, n), n


def test_stub_structures_match_the_header():
    ns = {"C": C, "np": np}
    for doc_name, bound in (("Params", _lib.Params), ("Scores", _lib.Scores), ("SegConfig", SegConfig)):
        exec(_class_source(doc_name), ns)
        doc = ns[doc_name]
        assert C.sizeof(doc) == C.sizeof(bound), (doc_name, C.sizeof(doc), C.sizeof(bound))
        assert [(n, getattr(doc, n).offset) for n, _ in doc._fields_] == \
               [(n, getattr(bound, n).offset) for n, _ in bound._fields_], doc_name


def test_stub_prototypes_match_the_header():
    for name in ("cia_create", "cia_screen_fields_host"):
        m = re.search(rf"lib\.{name}\.argtypes = \[(.*?)\]\n", DOC, flags=re.S)
        assert m, name
        doc_types = eval("[" + m.group(1) + "]", {"C": C})
        bound = _lib.SIGNATURES[name][1]
        assert len(doc_types) == len(bound), (name, len(doc_types), len(bound))
        for d, b in zip(doc_types, bound):       # same width and kind; pointer flavours may differ
            assert C.sizeof(d) == C.sizeof(b), (name, d, b)


def test_stub_cell_record_matches_the_header():
    m = re.search(r"cells = np\.empty\(cap, dtype=(\[.*?\])\)", DOC, flags=re.S)
    dt = np.dtype(eval(m.group(1)))
    assert dt.itemsize == _lib.CELL_DTYPE.itemsize == 56
    assert [dt.fields[k][1] for k in dt.names] == [_lib.CELL_DTYPE.fields[k][1] for k in _lib.CELL_DTYPE.names]
