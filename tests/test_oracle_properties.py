"""Property tests (hypothesis) of the C oracle against an inline restatement of the reference's three
scalar formulas in plain Python floats -- the same expressions as src/components/pre_llm_injector.py:374-388,
src/pipeline/retriever_hybrid.py:655-664 and src/utils/embedding_utils.py:29-39, evaluated by THIS
interpreter's builtin sum() (Neumaier on CPython >= 3.12)."""
import math
import sys

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle, synth

MODE = oracle.SUM_NEUMAIER if sys.version_info >= (3, 12) else oracle.SUM_NAIVE
finite = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, allow_infinity=False, width=64)
vec = st.lists(finite, min_size=0, max_size=40)


def py_injector(vec1, vec2):
    if len(vec1) != len(vec2):
        return 0.0
    dot_product = sum(a * b for a, b in zip(vec1, vec2))
    norm1 = math.sqrt(sum(a * a for a in vec1))
    norm2 = math.sqrt(sum(b * b for b in vec2))
    if norm1 == 0 or norm2 == 0:
        return 0.0
    return dot_product / (norm1 * norm2)


def py_retriever(vec1, vec2):
    dot_product = sum(a * b for a, b in zip(vec1, vec2))
    mag1 = math.sqrt(sum(a * a for a in vec1))
    mag2 = math.sqrt(sum(b * b for b in vec2))
    if mag1 * mag2 == 0:
        return 0.0
    return dot_product / (mag1 * mag2)


def py_utils(vec1, vec2):
    dot_product = sum(a * b for a, b in zip(vec1, vec2))
    magnitude1 = sum(a * a for a in vec1) ** 0.5
    magnitude2 = sum(b * b for b in vec2) ** 0.5
    if magnitude1 == 0 or magnitude2 == 0:
        return 0.0
    return dot_product / (magnitude1 * magnitude2)


@settings(max_examples=300, deadline=None)
@given(vec, vec)
def test_scalar_variants_bit_exact(v1, v2):
    for name, fn in (("injector", py_injector), ("retriever", py_retriever), ("utils", py_utils)):
        got = oracle.cosine(v1, v2, name, MODE)
        want = fn(v1, v2)
        assert got == want or (math.isnan(got) and math.isnan(want)), (name, v1, v2, got, want)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 40), st.integers(1, 12), st.integers(1, 5), st.integers(1, 12), st.integers(0, 2 ** 31 - 1))
def test_topk_matches_python_sort(n, d, q, k, seed):
    rng = np.random.default_rng(seed)
    X = rng.integers(-3, 4, size=(n, d)).astype(np.float64) / 4.0     # few distinct values -> many exact ties
    Q = rng.integers(-3, 4, size=(q, d)).astype(np.float64) / 4.0
    got = oracle.batch_similarities(Q, X, k, sum_mode=MODE)
    for qi in range(q):
        sims = [(i, float(py_injector(list(Q[qi]), list(X[i])))) for i in range(n)]
        sims.sort(key=lambda x: x[1], reverse=True)                    # stable, like the reference (:369)
        assert got[qi] == sims[:k]


def test_streamed_tier_equals_blocked_tier():
    """oracle.topk_streamed (float32 pre-ranking + completeness proof, rows fetched block by block) == oracle.topk_blocked
    (float64 pre-ranking over one array) == the plain reference loop, on the same rows."""
    n, d, k = 20_000, 96, 7
    X = oracle.synth_rows_c(9, 0, n, d)
    Q = np.random.default_rng(4).standard_normal((5, d)).astype(np.float32)
    Q[0] = X[777]
    a = oracle.topk_blocked(Q, X, k)
    b = oracle.topk_streamed(Q, n, lambda b0, b1: X[b0:b1], lambda r: X[r], k, block=3000)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    ref = oracle.batch_similarities(Q, X, k)
    assert [[r for r, _ in lst] for lst in ref] == b[0].tolist()
    assert [[s for _, s in lst] for lst in ref] == b[1].tolist()



def test_interpreter_loop_equals_the_c_restatement():
    """oracle/pyloop.py (what CPython computes, generator expressions and the interpreter's own sum()) against
    vm_oracle.c: same rows, bit-identical scores -- incl. a zero row, a falsy row, a wrong-length row, duplicates and an
    Exception query.  The C side uses the summation order of the running interpreter (Neumaier from 3.12 on)."""
    import sys
    from oracle import pyloop
    mode = oracle.SUM_NEUMAIER if sys.version_info >= (3, 12) else oracle.SUM_NAIVE
    X = synth.synth_rows(31, 0, 400, 96).astype(np.float64)
    X *= 1.0 + 1e-9 * np.random.default_rng(3).standard_normal(X.shape)       # off the bf16 / fp32 grid
    X[7] = 0.0
    X[200] = X[20]
    Q = synth.synth_queries(32, 5, 96, 31, 400).astype(np.float64)
    store = {f"c{i}": [float(v) for v in X[i]] for i in range(len(X))}
    store["c11"] = []                    # falsy: skipped (:363)
    store["c12"] = store["c12"][:50]     # wrong length: scores 0.0 (:378-379)
    row_ok = np.ones(len(X), np.uint8); row_ok[11] = 0
    Xc = X.copy(); Xc[12] = 0.0          # a zero row scores 0.0 as well
    queries = [[float(v) for v in q] for q in Q]
    queries.insert(2, RuntimeError("embedding failed"))
    got = pyloop.batch_similarities(queries, store, 7)
    qok = np.ones(len(queries), np.uint8); qok[2] = 0
    Qc = np.insert(Q, 2, 0.0, axis=0)
    want = oracle.batch_similarities(Qc, Xc, 7, query_ok=qok, row_ok=row_ok, sum_mode=mode)
    assert [[(int(c[1:]), s) for c, s in lst] for lst in got] == want
    assert got[2] == []


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 220), st.sampled_from([8, 33, 96]), st.integers(0, 9), st.sampled_from([0.0, 0.35, 0.8, 0.9, 0.97]),
       st.sampled_from([16, 64, 100, 4096]), st.integers(0, 2 ** 31 - 1))
def test_streamed_pair_tier_property(n, d, dup, thr, block, seed):
    """oracle.pairs_above_streamed == oracle.pairs_above for random shapes, duplicate densities, thresholds and block
    sizes (ragged last blocks, a single block), with general float32 values, zero rows and exact duplicates mixed in."""
    rng = np.random.default_rng(seed)
    E = oracle.synth_rows_c(seed % 1000, 0, n, d, dup).astype(np.float32)
    E *= (1.0 + 1e-3 * rng.standard_normal(E.shape)).astype(np.float32)      # off the 1/128 grid
    if n > 4:
        E[rng.integers(n)] = 0.0
        E[rng.integers(n)] = E[rng.integers(n)]
    a = oracle.pairs_above(E, thr)
    b = oracle.pairs_above_streamed(E, thr, block=block)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
