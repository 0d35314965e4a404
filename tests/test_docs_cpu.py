"""Doc drift guard: every C-ABI symbol and adapter entry point named in INTEGRATION.md / DESIGN.md / README.md
exists (in include/vidmem.h + the ctypes binding, or as an attribute of the Python module it is quoted from)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["INTEGRATION.md", "DESIGN.md", "README.md"]


def _text(name):
    with open(os.path.join(ROOT, name), encoding="utf-8") as f:
        return f.read()


def test_abi_symbols_named_in_docs_exist():
    import vidmem_b200
    from vidmem_b200 import _lib
    header = _text("include/vidmem.h")
    known_types = {"vm_store", "vm_comm", "vm_status", "vm_dtype", "vm_mem", "vm_score_mode", "vm_sum_mode", "vm_topk_stats", "vm_oracle"}
    for doc in DOCS:
        for sym in sorted(set(re.findall(r"\bvm_[a-z0-9_]+\b", _text(doc)))):
            if sym in known_types or sym.endswith("_"):
                continue
            assert sym in _lib.EXPORTS and re.search(r"\b%s\s*\(" % sym, header), f"{doc} names {sym}, which the ABI does not export"


@pytest.mark.parametrize("module,pattern", [("adapters", r"\badapters\.([A-Za-z_]+)"), ("sharded", r"\bsharded\.([A-Za-z_]+)"),
                                            ("dedup", r"\bdedup\.([A-Za-z_]+)")])
def test_python_entry_points_named_in_docs_exist(module, pattern):
    import importlib
    import vidmem_b200
    mod = importlib.import_module(f"vidmem_b200.{module}")
    for doc in DOCS:
        for name in sorted(set(re.findall(pattern, _text(doc)))):
            if name in ("py",):
                continue
            assert hasattr(mod, name), f"{doc} names {module}.{name}, which does not exist"


def test_methods_quoted_in_integration_md_exist():
    """`obj.method(` calls shown in INTEGRATION.md resolve on the classes they are shown on."""
    import vidmem_b200
    from vidmem_b200 import adapters, sharded, store
    text = _text("INTEGRATION.md")
    shown = {"backend": adapters.ChunkSimilarityBackend, "resident": adapters.ResidentChunkStore, "comm": sharded.Communicator,
             "store2": store.EmbeddingStore, "st": adapters.ResidentChunkStore}
    checked = 0
    for var, cls in shown.items():
        for name in sorted(set(re.findall(r"\b%s\.([a-z_]+)\(" % var, text))):
            assert hasattr(cls, name), f"INTEGRATION.md calls {var}.{name}(), which {cls.__name__} does not have"
            checked += 1
    assert checked >= 5


def test_bench_self_launch_line_follows_the_contract():
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    argv = bench.torchrun_argv(4, ["--gpus", "4", "--steps", "10"], 29517)
    assert argv[:3] == [sys.executable, "-m", "torch.distributed.run"]
    assert "--nnodes=1" in argv and "--nproc-per-node=4" in argv
    assert argv[argv.index("--master-addr") + 1] == "127.0.0.1" and argv[argv.index("--master-port") + 1] == "29517"
    assert argv[-5].endswith("bench.py") and argv[-4:] == ["--gpus", "4", "--steps", "10"]
