"""CPU (no GPU): the C-ABI library loads and exports exactly what include/prs.h declares, fails
loudly without a device (no fallback), and the host-side logic mirrors the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def P():
    import persian_rag_system_b200 as P
    return P


def test_library_exports_every_declared_symbol(P):
    hdr = open(os.path.join(ROOT, "include", "prs.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(prs_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 28
    lib = ctypes.CDLL(P._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/prs.h but not exported by libprs.so"
    assert declared == set(P._lib.SIGNATURES), declared ^ set(P._lib.SIGNATURES)


def test_library_is_sm100a_only_with_tcgen05_and_tma():
    """Evidence that the shipped binary is the Blackwell-native path (B200_PROFILING.md table)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    import persian_rag_system_b200 as P
    elf = subprocess.run(["cuobjdump", "-lelf", P._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf)
    sass = subprocess.run(["cuobjdump", "-sass", P._lib.LIB_PATH], capture_output=True, text=True).stdout
    # tcgen05.mma / tcgen05.ld / tcgen05.st / cp.async.bulk (the T64 corpus layout makes every TMA
    # transfer a contiguous bulk copy, so no tensor-map UTMALDG is needed)
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_scan_kernel_hot_loop_keeps_uniform_datapath_branches():
    """Regression guard for a codegen cliff measured on B200 (one-launch search): when the kernel's parameter
    block has its address taken / is indexed dynamically, or too much merge code becomes reachable from the scan kernel,
    ptxas drops the uniform-datapath branches (BRA.U / UISETP) of the epilogue loop and guards its __syncwarp with
    WARPSYNC.ALL -- the scan kernel then runs ~15 % slower.  Every instantiation must keep them."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    import persian_rag_system_b200 as P
    sass = subprocess.run(["cuobjdump", "-sass", P._lib.LIB_PATH], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    seen = 0
    for f in funcs:
        if "flat_scan_umma_kernel" not in f.split("\n", 1)[0]:
            continue
        seen += 1
        lines = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l)]
        first = next(i for i, l in enumerate(lines) if "LDTM" in l)
        window = "\n".join(lines[max(0, first - 40): first + 500])
        assert "BRA.U" in window and "UISETP" in window, f.split("\n", 1)[0]
        assert "WARPSYNC.ALL" not in "\n".join(lines[max(0, first - 40): first]), f.split("\n", 1)[0]
    assert seen == 6


def test_no_cpu_fallback_without_a_device(P):
    arch = P.lib().prs_device_arch(0)
    if arch == 100:
        pytest.skip("a B200 is present")
    with pytest.raises(P.PrsError) as e:
        P.IndexFlatL2(8)
    assert e.value.code == -2 and "CUDA" in str(e.value)
    with pytest.raises(P.PrsError):
        P.read_index(os.path.join(ROOT, "tests", "golden", "indices", "drugs_sentence_chunks.index"))
    with pytest.raises(P.PrsError):
        P.BM25Index([["a", "b"], ["b"]])
    r = P.RetrievalSystem(method="bm25")
    assert r.load_chunks([{"id": "c0", "text": "a b"}]) is False and r.retrieve("a") == []


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "persian-rag-system_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f


# ------------------------------------------------------------------ sparse host builders
def test_bm25_host_weights_are_rank_bm25_bit_for_bit(P, gold_dir, golden_texts):
    from persian_rag_system_b200.sparse import build_bm25_csr
    chunks, queries = golden_texts
    docs = [c["text"].split() for c in chunks]
    b = build_bm25_csr(docs)
    ob = O.BM25OkapiOracle(docs)
    assert b["avgdl"] == ob.avgdl and b["average_idf"] == ob.average_idf
    assert {w: b["idf"][i] for w, i in b["vocab"].items()} == ob.idf
    gold = np.load(os.path.join(gold_dir, "bm25_golden.npz"))["scores"]
    import scipy.sparse as sp
    M = sp.csr_matrix((b["weights"], b["indices"], b["indptr"]), shape=(len(docs), len(b["vocab"]))).tocsc()
    for r, q in enumerate(queries):
        s = np.zeros(len(docs))
        for tok in q.split():
            t = b["vocab"].get(tok, -1)
            if t >= 0:
                s += M[:, t].toarray().ravel()
        assert np.array_equal(s, gold[r])


def test_tfidf_host_vectoriser_is_sklearn_bit_for_bit(P, gold_dir, golden_texts):
    from persian_rag_system_b200.sparse import TfidfVectorizerHost
    chunks, queries = golden_texts
    texts = [c["text"] for c in chunks]
    vec, mat = O.tfidf_fit(texts)
    mat.sort_indices()
    h = TfidfVectorizerHost(10000, (1, 2))
    indptr, indices, data = h.fit(texts)
    assert h.vocabulary_ == {k: int(v) for k, v in vec.vocabulary_.items()}
    assert np.array_equal(h.idf_, vec.idf_)
    assert np.array_equal(indptr, mat.indptr) and np.array_equal(indices, mat.indices)
    assert np.array_equal(h.tfidf_data_, mat.data)
    gold = np.load(os.path.join(gold_dir, "tfidf_golden.npz"))["scores"]
    import scipy.sparse as sp
    M = sp.csr_matrix((data, indices, indptr), shape=(len(texts), h.n_features)).tocsc()
    qi, qt, qw = h.encode_queries(queries)
    for r in range(len(queries)):
        s = np.zeros(len(texts))
        for e in range(qi[r], qi[r + 1]):
            s += qw[e] * M[:, qt[e]].toarray().ravel()
        assert np.array_equal(s, gold[r])


def test_tfidf_max_features_pruning_matches_sklearn(P):
    from persian_rag_system_b200.sparse import TfidfVectorizerHost
    rng = np.random.default_rng(3)
    words = [f"tok{i}" for i in range(400)]
    texts = [" ".join(rng.choice(words, size=60)) for _ in range(80)]
    from sklearn.feature_extraction.text import TfidfVectorizer
    vec = TfidfVectorizer(max_features=300, ngram_range=(1, 2))
    mat = vec.fit_transform(texts)
    mat.sort_indices()
    h = TfidfVectorizerHost(300, (1, 2))
    indptr, indices, _ = h.fit(texts)
    assert h.vocabulary_ == {k: int(v) for k, v in vec.vocabulary_.items()}
    assert np.array_equal(indices, mat.indices) and np.allclose(h.tfidf_data_, mat.data, rtol=0, atol=1e-15)


# ------------------------------------------------------------------ retriever host logic
class _StubDense:
    ntotal = 6

    def search(self, q, k):
        d = np.array([[0.1, 0.2, 0.4, 0.8, 1.6, 3.2]], np.float32)[:, :k]
        i = np.array([[4, 2, 0, 5, 9, -1]], np.int64)[:, :k]          # 9 is out of range, -1 is padding
        return np.repeat(d, len(q), 0), np.repeat(i, len(q), 0)


class _StubSparse:
    def get_top_k(self, tokens, k):
        return np.array([3.0, 2.0, 0.0, 0.0])[:k], np.array([2, 1, 5, 4])[:k]


class _Enc:
    def encode(self, s, device=None):                     # the bare reference signature (no convert_to_tensor)
        return np.zeros((len(s), 4), np.float32)


def _stub_system(P, method):
    r = P.RetrievalSystem(method=method, encoder=_Enc())
    r.chunks = [{"id": f"word_chunk_{i}", "text": "x" * (900 + i), "chunk_type": "word_based"} for i in range(6)]
    r.faiss_index, r.bm25_index, r.is_ready = _StubDense(), _StubSparse(), True
    return r


def test_dense_filters_out_of_range_ids_and_scores_like_the_reference(P):
    r = _stub_system(P, "dense")
    res = r.retrieve("q", top_k=6)
    assert [c["id"] for c, _ in res] == ["word_chunk_4", "word_chunk_2", "word_chunk_0", "word_chunk_5"]
    assert [float(s) for _, s in res] == [float(1 / (1 + d)) for d in np.array([0.1, 0.2, 0.4, 0.8], np.float32)]


def test_context_packing_and_dispatch_match_the_reference_restatement(P):
    """(The hybrid fusion itself is a device kernel now: its parity test is tests/test_retrieval_gpu.py.)"""
    r = _stub_system(P, "dense")
    ctx, meta = r.get_contexts_for_rag("q", top_k=5, max_context_length=2000)
    want_ctx, want_meta = O.pack_contexts(r.retrieve("q", 5), 2000)
    assert ctx == want_ctx and meta == want_meta and ctx[-1].endswith("...") and len(ctx) == 3
    # batched variants: one engine pass, element i == the per-query call
    assert r.retrieve_batch(["q", "q"], 5) == [r.retrieve("q", 5)] * 2
    assert r.get_contexts_for_rag_batch(["q"], 5, 2000) == [(ctx, meta)]
    rep = r.evaluate_retrieval_quality([{"id": "a", "question": "q"}, {"id": "b", "question": "q"}],
                                       {"a": ["word_chunk_2"], "b": []})
    assert rep == {"hit_at_1": 0.0, "hit_at_3": 1.0, "hit_at_5": 1.0, "mrr": 0.5, "total_queries": 2}
    r.method = "nope"
    assert r.retrieve("q") == [] and r.retrieve_batch(["q"]) == [[]]
    r.is_ready = False
    assert r.retrieve_batch(["q", "q"]) == [[], []]


def test_shard_bounds_and_merge_rule(P):
    from persian_rag_system_b200.sharded import merge_topk_host_lists, shard_bounds
    assert [shard_bounds(10, 4, g) for g in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(2, 4, 3) == (2, 2)
    rng = np.random.default_rng(0)
    x = rng.integers(0, 4, size=(60, 3)).astype(np.float32)          # many exact ties
    q = x[:5].copy()
    k = 7
    Dw, Iw = O.flat_search_c(x, q, k, O.METRIC_L2, form=1)
    Dp, Ip = [], []
    for g in range(4):
        lo, hi = shard_bounds(60, 4, g)
        D, I = O.flat_search_c(x[lo:hi], q, k, O.METRIC_L2, form=1)
        Dp.append(D)
        Ip.append(np.where(I >= 0, I + lo, -1))
    D, I = merge_topk_host_lists(np.stack(Dp), np.stack(Ip), largest=False)
    assert np.array_equal(I, Iw) and np.array_equal(D, Dw)


def test_t64_corpus_layout_is_a_swizzled_bijection(tmp_path):
    """The HBM layout of 16-bit corpora (csrc/common.cuh::t64_offset) checked on the host: bijective per
    64-row block, k-block-major, chunk ^ (row & 7) swizzle -- the image tcgen05's SWIZZLE_128B descriptor expects."""
    import shutil
    import subprocess
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    src = os.path.join(ROOT, "tests", "native", "t64_layout_check.cu")
    exe = str(tmp_path / "t64_layout_check")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-o", exe, src], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr


# ------------------------------------------------------------------ index build mirror (§8 a-2), host logic
def test_index_build_naming_skip_and_missing_file(P, tmp_path, capsys):
    """create_model_embeddings keeps the reference's contract without touching the GPU: file naming
    (src/create_embeddings.py:62), skip-if-exists -> True (:64-66), missing chunk file -> False (:68-70)."""
    d = str(tmp_path / "faiss")
    assert P.index_path_for("models/e5-base-ft", "sentence") == "results/faiss/e5-base-ft_drugs_sentence_chunks.index"
    path = P.index_path_for("models/e5-base-ft", "word", d)
    os.makedirs(d)
    open(path, "wb").write(b"already here")
    assert P.create_model_embeddings("models/e5-base-ft", "no_such_chunks.csv", "word", encoder=object(), faiss_dir=d) is True
    assert open(path, "rb").read() == b"already here"                      # untouched
    assert P.create_model_embeddings("models/e5-base-ft", "no_such_chunks.csv", "sentence", encoder=object(), faiss_dir=d) is False
    assert "not found" in capsys.readouterr().out
