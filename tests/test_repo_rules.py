"""Structural rules of the repository, checked on the source tree (no GPU, no imports of the product):

* `oracle/` is test infrastructure: only tests/, __graft_entry__.py (smoke) and bench.py (CPU legs) may import it;
* nothing that runs on the GPU box may read /root/reference: only the fixture generator, its shims and the
  `reference_available()`-gated live checks in tests/ mention it;
* the product package has no CPU fallback: it never imports the oracle and every measure / operator goes through _lib.
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMPORTS_ORACLE = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)


def _python_files(*tops):
    for top in tops:
        base = os.path.join(ROOT, top)
        if os.path.isfile(base):
            yield top
            continue
        for dirpath, dirnames, files in os.walk(base):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", "build")]
            for f in files:
                if f.endswith(".py"):
                    yield os.path.relpath(os.path.join(dirpath, f), ROOT)


def _read(rel):
    with open(os.path.join(ROOT, rel)) as f:
        return f.read()


def test_only_tests_smoke_and_bench_import_the_oracle():
    offenders = [p for p in _python_files("acav100m_b200", "tools") if IMPORTS_ORACLE.search(_read(p))]
    assert offenders == []
    assert IMPORTS_ORACLE.search(_read("bench.py")) and IMPORTS_ORACLE.search(_read("__graft_entry__.py"))
    # bench.py touches the oracle only inside its CPU legs, __graft_entry__ only inside build() / smoke()
    bench = _read("bench.py")
    for m in IMPORTS_ORACLE.finditer(bench):
        head = bench[:m.start()]
        assert head.rfind("\ndef cpu_") > head.rfind("\ndef run_"), "oracle import outside the cpu_* legs of bench.py"


def test_reference_tree_is_only_read_by_fixture_tooling():
    allowed = {"oracle/gen_golden.py", "oracle/ref_shims.py"}
    mention = re.compile(r"/root/reference")
    for p in _python_files("acav100m_b200", "tools", "bench.py", "__graft_entry__.py", "oracle"):
        if p in allowed:
            continue
        src = _read(p)
        code = "\n".join(line.split("#", 1)[0] for line in src.splitlines())
        code = re.sub(r'"""(.|\n)*?"""', "", code)                 # docstrings may cite the reference's paths
        assert not mention.search(code), p
    for p in _python_files("tests"):
        src = _read(p)
        if "ref_shims.load_reference" in src or "REFERENCE_ROOT" in src:
            assert "reference_available()" in src, p + " uses the reference without the availability gate"


def test_product_operators_have_no_cpu_path():
    for p in ("acav100m_b200/subset_selection/measures/mem_mi.py", "acav100m_b200/subset_selection/measures/dense_mi.py",
              "acav100m_b200/subset_selection/measures/batch_mi.py", "acav100m_b200/subset_selection/measures/pairs_engine.py",
              "acav100m_b200/clustering/sgd_clustering.py"):
        src = _read(p)
        assert "_lib.call(" in src, p
        assert "require_cuda" in src or "no CPU path" in src or "pairs_engine" in p, p


def test_no_undefined_names_in_gpu_only_code_paths():
    """Most of the host code only runs on the GPU box; a name that is never bound anywhere in its module (a missing
    import, a typo) must not wait for that run to show up."""
    import ast
    import builtins
    known = set(dir(builtins)) | {"__file__"}
    offenders = []
    for rel in _python_files("acav100m_b200", "tools", "tests", "oracle", "bench.py", "__graft_entry__.py"):
        tree = ast.parse(_read(rel))
        bound = set(known)
        for n in ast.walk(tree):
            if isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
                bound.add(n.name)
            elif isinstance(n, ast.Import):
                bound.update((a.asname or a.name).split(".")[0] for a in n.names)
            elif isinstance(n, ast.ImportFrom):
                bound.update(a.asname or a.name for a in n.names)
            elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                bound.add(n.id)
            elif isinstance(n, ast.arg):
                bound.add(n.arg)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                bound.add(n.name)
            elif isinstance(n, (ast.Global, ast.Nonlocal)):
                bound.update(n.names)
        offenders += [(rel, n.id, n.lineno) for n in ast.walk(tree)
                      if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in bound]
    assert offenders == []
