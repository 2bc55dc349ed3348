"""The oracle is test infrastructure: only tests/, __graft_entry__.smoke() / build() and bench.py's CPU arms may
import it.  This walks the syntax trees of everything else in the repository and fails on an `oracle` import, and
checks that bench.py reaches the oracle only from the functions of its CPU arms -- no GPU needed."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cell-image-analysis_b200")
CPU_ARM_FUNCTIONS = {"_ref_field", "cpu_baseline_inline", "run_reference", "segmentation_extras"}


def _oracle_imports(path):
    """[(line, enclosing function or None)] of every `import oracle...` / `from oracle... import` in a file"""
    tree = ast.parse(open(path).read(), path)
    hits = []

    def walk(node, fn):
        for child in ast.iter_child_nodes(node):
            inner = child.name if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)) else fn
            if isinstance(child, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in child.names):
                hits.append((child.lineno, fn))
            if isinstance(child, ast.ImportFrom) and child.level == 0 and (child.module or "").split(".")[0] == "oracle":
                hits.append((child.lineno, fn))
            walk(child, inner)

    walk(tree, None)
    return hits


def _python_files(top):
    for base, dirs, files in os.walk(top):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", ".git", "gpurun_out")]
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(base, f)


def test_product_package_tools_and_profiles_never_import_the_oracle():
    offenders = []
    for top in (PKG, os.path.join(ROOT, "tools"), os.path.join(ROOT, "profiles")):
        for p in _python_files(top):
            offenders += [(os.path.relpath(p, ROOT), line) for line, _ in _oracle_imports(p)]
    offenders += [("cell_image_analysis_b200.py", line)
                  for line, _ in _oracle_imports(os.path.join(ROOT, "cell_image_analysis_b200.py"))]
    assert not offenders, offenders


def test_native_sources_do_not_reach_for_the_oracle():
    for p in sorted(os.listdir(os.path.join(PKG, "csrc"))):
        text = open(os.path.join(PKG, "csrc", p), errors="replace").read()
        assert "#include \"oracle" not in text and "../oracle" not in text and "dlopen(\"oracle" not in text, p


def test_bench_reaches_the_oracle_only_from_its_cpu_arms():
    hits = _oracle_imports(os.path.join(ROOT, "bench.py"))
    assert hits, "bench.py's cpu_baseline / --impl reference legs time the oracle port"
    assert {fn for _, fn in hits} <= CPU_ARM_FUNCTIONS, hits
    # the CPU leg inside segmentation_extras is the one behind `if with_cpu:` (a reported baseline, never the value)
    src = open(os.path.join(ROOT, "bench.py")).read()
    seg = src[src.index("def segmentation_extras"):src.index("def run_native")]
    assert seg.index("if with_cpu:") < seg.index("from oracle")


def test_graft_entry_uses_the_oracle_as_builder_and_checker_only():
    hits = _oracle_imports(os.path.join(ROOT, "__graft_entry__.py"))
    assert {fn for _, fn in hits} <= {"build", "smoke"}, hits
