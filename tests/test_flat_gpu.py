"""GPU parity: flat dense search through the C ABI (ctypes) vs the CPU oracle.

fp32 storage is the exact-parity mode: id lists identical to the oracle, distances within 1e-5
relative (summation order differs).  fp16/bf16 storage is checked tie-aware against the fp32
oracle run on the STORED (rounded) rows with the north-star tolerance 1e-3 relative.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL_F32 = 1e-5          # north_star: <= 1e-5 for fp32
RTOL_16 = 1e-3           # north_star: <= 1e-3 relative for fp16/bf16 storage


@pytest.fixture(scope="module")
def P():
    import persian_rag_system_b200 as P
    assert P.lib().prs_device_arch(0) == 100, P._lib.last_error()
    return P


def _round_to(x, storage):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if storage == "fp16":
        return t.half().float().numpy()
    if storage == "bf16":
        return t.bfloat16().float().numpy()
    return x.astype(np.float32)


def _check_exact(D, I, x, q, k, metric, what):
    Dr, Ir = O.flat_search_c(x, q, k, metric, form=1)
    S = O.flat_scores_f64(x, q, metric)
    flips = 0
    for r in range(q.shape[0]):
        flips += O.check_topk_against_scores(I[r], D[r], S[r], k, metric == O.METRIC_IP, rtol=RTOL_F32, atol=1e-6,
                                             what=f"{what} q{r}")
    return flips, Ir


# ------------------------------------------------------------------ the reference's own indices
@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
def test_golden_indices_fp32_ids_identical(P, gold_dir, golden_indices, metric):
    g = np.load(os.path.join(gold_dir, "flat_golden.npz"))
    tag = "l2" if metric == O.METRIC_L2 else "ip"
    for t, f in enumerate(g["files"].tolist()):
        x, _ = golden_indices[f]
        idx = P.FlatIndex(x.shape[1], metric, "fp32")
        idx.add(x)
        assert idx.ntotal == x.shape[0] and idx.d == x.shape[1]
        q = g[f"q_{t}"]
        for k in (1, 5, 10, 20):
            D, I = idx.search(q, k)
            assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (64, k)
            assert np.array_equal(I, g[f"{tag}_I_{t}_{k}"]), f"{f} k={k}: id lists differ from the oracle"
            np.testing.assert_allclose(D, g[f"{tag}_D_{t}_{k}"], rtol=RTOL_F32, atol=1e-6)
        # the reference only ever searches one query at a time (src/retrieval.py:98-102)
        for r in range(8):
            D1, I1 = idx.search(q[r:r + 1], 5)
            assert np.array_equal(I1[0], g[f"{tag}_I_{t}_5"][r])
        assert idx.last_path == "cuda-core"


def test_read_index_of_reference_file_and_self_search(P, gold_dir):
    d = os.path.join(gold_dir, "indices")
    for f in os.listdir(d):
        idx = P.read_index(os.path.join(d, f))
        x, metric = O.read_faiss_flat(os.path.join(d, f))
        assert idx.metric_type == metric == P.METRIC_L2 and idx.ntotal == x.shape[0]
        D, I = idx.search(x, 1)
        assert (I[:, 0] == np.arange(x.shape[0])).all() and (D[:, 0] == 0).all()
        assert np.array_equal(idx.reconstruct_n(0, idx.ntotal), x)


def test_write_index_is_byte_exact_faiss_file(P, gold_dir, tmp_path):
    d = os.path.join(gold_dir, "indices")
    for f in os.listdir(d):
        idx = P.read_index(os.path.join(d, f))
        out = tmp_path / f
        P.write_index(idx, str(out))
        assert open(out, "rb").read() == open(os.path.join(d, f), "rb").read()
    # IP index gets fourcc IxFI and round-trips
    x = np.random.default_rng(0).standard_normal((10, 32)).astype(np.float32)
    ip = P.IndexFlatIP(32)
    ip.add(x)
    P.write_index(ip, str(tmp_path / "ip.index"))
    assert open(tmp_path / "ip.index", "rb").read()[:4] == b"IxFI"
    x2, m2 = O.read_faiss_flat(str(tmp_path / "ip.index"))
    assert m2 == O.METRIC_IP and np.array_equal(x2, x)


# ------------------------------------------------------------------ BASELINE configs[0] on 16-bit storage
@pytest.mark.parametrize("storage", ["fp16", "bf16"])
def test_golden_indices_16bit_l2_c1_queries(P, golden_indices, storage):
    """The reference's own indices (125/121 rows, d = 384/512/768, ||x||^2 up to ~25), squared L2 like
    the reference (src/retrieval.py:102), 1 000 C1 queries (seeded Gaussian perturbations of
    cyclic rows), k = 5, 16-bit storage on the tcgen05 scan.  Checked tie-aware at the north-star
    tolerance (1e-3 relative) against float64 scores of the STORED rows and rounded queries."""
    from oracle.make_golden import make_queries
    for t, (name, (x, _)) in enumerate(sorted(golden_indices.items())):
        q = make_queries(x, 1000, seed=100 + t)
        idx = P.FlatIndex(x.shape[1], P.METRIC_L2, storage)
        idx.add(x)
        D, I = idx.search(q, 5)
        assert idx.last_path == "tcgen05"
        xs, qs = _round_to(x, storage), _round_to(q, storage)
        S = O.flat_scores_f64(xs, qs, O.METRIC_L2)
        flips = 0
        for r in range(q.shape[0]):
            flips += O.check_topk_against_scores(I[r], D[r], S[r], 5, False, rtol=RTOL_16, atol=1e-6, what=f"{name} {storage} q{r}")
        assert flips <= 50, (name, flips)                   # of 5 000 positions: two distances tying within fp32 rounding
        # the reference's call shape: one query per search (nq = 1 takes the same kernel, no cluster)
        for r in range(0, 1000, 97):
            D1, I1 = idx.search(q[r:r + 1], 5)
            assert np.array_equal(I1[0], I[r]) and np.array_equal(D1[0], D[r])
        # the fp32 exact-parity index on the same queries: same neighbours wherever 16-bit rounding keeps
        # the distances apart (sanity of the storage mode, not a parity claim)
        ref = P.FlatIndex(x.shape[1], P.METRIC_L2, "fp32")
        ref.add(x)
        D32, I32 = ref.search(q, 5)
        assert (I32[:, 0] == I[:, 0]).mean() > (0.98 if storage == "fp16" else 0.6)      # bf16 keeps 8 bits: near-duplicate rows merge


@pytest.mark.parametrize("storage", ["fp16", "bf16"])
def test_16bit_l2_near_duplicate_rows_get_direct_form_distances(P, golden_indices, storage):
    """SURVEY findings 2/7: fine-tuned indices hold near-duplicate rows (squared distances down to 5e-5
    at ||x||^2 ~ 25) where the expanded form ||q||^2 + ||x||^2 - 2q.x loses every digit in fp32.  The
    merge recomputes the selected rows' distances in the direct form: self-distance is exactly 0 and the
    neighbours' distances are right to 1e-3 relative even at 1e-5 absolute."""
    name = "paraphrase-multilingual-MiniLM-L12-v2_finetuned_drugs_word_chunks.index"
    x, _ = golden_indices[name]
    rng = np.random.default_rng(5)
    xs = _round_to(x, storage)
    # rows 0..19 get a twin that differs by a few storage ulps in 8 coordinates
    twins = xs[:20].copy()
    cols = rng.integers(0, x.shape[1], size=(20, 8))
    for i in range(20):
        twins[i, cols[i]] = np.nextafter(_round_to(twins[i, cols[i]] * (1 + 2.0 ** (-7 if storage == "bf16" else -10)), storage),
                                         np.float32(np.inf)).astype(np.float32)
    twins = _round_to(twins, storage)
    corpus = np.concatenate([xs, twins])
    idx = P.FlatIndex(x.shape[1], P.METRIC_L2, storage)
    idx.add(corpus)
    q = corpus[:40].copy()
    q[20:40] = twins
    D, I = idx.search(q, 4)
    assert idx.last_path == "tcgen05"
    S = O.flat_scores_f64(corpus, q, O.METRIC_L2)
    for r in range(40):
        O.check_topk_against_scores(I[r], D[r], S[r], 4, False, rtol=RTOL_16, atol=1e-7, what=f"twin q{r}")
    assert (D[:, 0] == 0).all()                              # direct form: the row itself is at exactly 0
    assert (I[:20, 0] == np.arange(20)).all() and (I[20:40, 0] == corpus.shape[0] - 20 + np.arange(20)).all()
    assert (I[:20, 1] == corpus.shape[0] - 20 + np.arange(20)).all()          # then its twin
    assert ((D[:20, 1] > 0) & (D[:20, 1] < 1e-2)).all()


# ------------------------------------------------------------------ shapes and edge cases
@pytest.mark.parametrize("n,d,nq,k", [(1, 8, 1, 1), (3, 5, 2, 5), (257, 100, 3, 7), (1000, 384, 1, 10),
                                      (1000, 384, 17, 10), (4099, 512, 9, 100), (2500, 768, 5, 33),
                                      (5000, 64, 4, 1024), (300, 1000, 2, 3)])
@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
def test_random_fp32_exact(P, n, d, nq, k, metric):
    rng = np.random.default_rng(n * 31 + d)
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    idx = P.FlatIndex(d, metric, "fp32")
    idx.add(x[: n // 2])
    idx.add(x[n // 2:])                      # growth path
    D, I = idx.search(q, k)
    flips, Ir = _check_exact(D, I, x, q, k, metric, f"n{n} d{d}")
    # against the fp32 C oracle: identical ids except where two fp32 sums in different orders tie
    Dr = O.flat_search_c(x, q, k, metric, form=1)[0]
    nflip = O.check_topk_lists(I, D, Ir, Dr, rtol=RTOL_F32, atol=1e-6, what=f"n{n} d{d}")
    assert flips <= 2 and nflip <= max(2, (nq * min(k, n)) // 200)


@pytest.mark.parametrize("n", [125, 170, 171, 400])          # 170 x 384 fp32 is the last size one CTA takes (256 KB)
@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
def test_small_index_single_cta_search(P, n, metric):
    """The reference's own call shape (src/retrieval.py:102: nq = 1, k = 5 on a 125-row fp32 index) goes through ONE
    CTA that writes D / I itself (no merge launch), with the queries read from page-locked memory: same answers as
    the oracle for pageable numpy, pinned and device inputs, several query groups, k beyond the corpus and k > 128."""
    import torch
    d = 384
    rng = np.random.default_rng(n + metric)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[n // 2] = x[3]                                          # an exact duplicate: lower id first
    idx = P.FlatIndex(d, metric, "fp32")
    idx.add(x)
    for nq, k in ((1, 5), (9, 5), (20, 16), (2, 150), (3, n + 7)):
        q = (x[rng.integers(0, n, nq)] + 0.05 * rng.standard_normal((nq, d))).astype(np.float32)
        Dr, Ir = O.flat_search_c(x, q, k, metric, form=1)
        D, I = idx.search(q, k)                               # pageable numpy
        assert idx.last_path == "cuda-core"
        # k reaches past the corpus: inner products near zero are sums of 384 O(1) terms, hence the absolute floor
        O.check_topk_lists(I, D, Ir, Dr, rtol=RTOL_F32, atol=5e-5, what=f"small n{n} nq{nq} k{k}")
        assert (I[:, min(k, n):] == -1).all()
        Dd, Id = idx.search(torch.from_numpy(q).cuda(), k)    # device tensors
        assert np.array_equal(Id.cpu().numpy(), I) and np.array_equal(Dd.cpu().numpy(), D)
        qp = torch.from_numpy(q).pin_memory()                 # caller-pinned host buffers
        Dp, Ip = idx.search(qp.numpy(), k)
        assert np.array_equal(Ip, I) and np.array_equal(Dp, D)


def test_empty_index_and_padding(P):
    idx = P.IndexFlatL2(16)
    D, I = idx.search(np.zeros((2, 16), np.float32), 3)
    assert (I == -1).all() and (D == np.finfo(np.float32).max).all()
    ip = P.IndexFlatIP(16)
    D, I = ip.search(np.zeros((1, 16), np.float32), 2)
    assert (I == -1).all() and (D == -np.finfo(np.float32).max).all()
    idx.add(np.eye(2, 16, dtype=np.float32))
    D, I = idx.search(np.eye(1, 16, dtype=np.float32), 4)
    assert I[0].tolist() == [0, 1, -1, -1] and D[0, 0] == 0 and D[0, 1] == 2 and D[0, 2] == np.finfo(np.float32).max
    D, I = idx.search(np.zeros((0, 16), np.float32), 4)
    assert D.shape == (0, 4)


def test_exact_duplicates_tie_order_is_id_ascending(P):
    rng = np.random.default_rng(3)
    base = rng.standard_normal((5, 48)).astype(np.float32)
    x = np.concatenate([base, base, base])           # every row appears 3 times
    for metric in (O.METRIC_L2, O.METRIC_IP):
        idx = P.FlatIndex(48, metric)
        idx.add(x)
        D, I = idx.search(base, 3)
        for r in range(5):
            assert I[r].tolist() == [r, r + 5, r + 10]


def test_error_codes(P):
    idx = P.IndexFlatL2(8)
    idx.add(np.zeros((4, 8), np.float32))
    with pytest.raises(P.PrsError) as e:
        idx.search(np.zeros((1, 9), np.float32), 1)          # dimension mismatch (faiss asserts)
    assert e.value.code == -1
    with pytest.raises(P.PrsError):
        idx.search(np.zeros((1, 8), np.float32), 0)
    with pytest.raises(P.PrsError):
        idx.search(np.zeros((1, 8), np.float32), 1025)
    with pytest.raises(P.PrsError):
        idx.add(np.zeros((1, 7), np.float32))
    with pytest.raises(P.PrsError) as e:
        P.read_index("/nonexistent/file.index")
    assert e.value.code == -3


# ------------------------------------------------------------------ 16-bit storage, both kernel families
@pytest.mark.parametrize("storage", ["fp16", "bf16"])
@pytest.mark.parametrize("path", ["cuda-core", "tcgen05"])
@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
@pytest.mark.parametrize("n,d,nq,k", [(4000, 384, 8, 10), (9000, 768, 64, 10), (3001, 512, 130, 5), (700, 100, 33, 16)])
def test_16bit_storage_vs_oracle(P, storage, path, metric, n, d, nq, k):
    rng = np.random.default_rng(n + d + nq)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    idx = P.FlatIndex(d, metric, storage)
    idx.add(x)
    idx.set_path(path)
    D, I = idx.search(q, k)
    assert idx.last_path == path
    xs = _round_to(x, storage)
    # the tensor-core path also rounds the queries to the storage type
    qs = _round_to(q, storage) if path == "tcgen05" else q
    S = O.flat_scores_f64(xs, qs, metric)
    flips = 0
    for r in range(nq):
        flips += O.check_topk_against_scores(I[r], D[r], S[r], k, metric == O.METRIC_IP, rtol=RTOL_16, atol=2e-5,
                                             what=f"{storage}/{path} q{r}")
    # and against fp32 queries (what a caller sees) with the north-star tolerance
    S32 = O.flat_scores_f64(xs, q, metric)
    for r in range(nq):
        O.check_topk_against_scores(I[r], D[r], S32[r], k, metric == O.METRIC_IP, rtol=RTOL_16, atol=2e-3 if storage == "bf16" else 3e-4,
                                    what=f"{storage}/{path} vs fp32 queries q{r}")
    assert flips <= nq * k * 0.05


@pytest.mark.parametrize("n,d,nq,k,metric,storage", [
    (30000, 384, 129, 10, O.METRIC_IP, "fp16"),      # cluster of 2, second query block nearly empty; two row blocks per tile
    (30000, 512, 256, 16, O.METRIC_L2, "fp16"),      # cluster of 2, both blocks full
    (30000, 768, 257, 10, O.METRIC_IP, "bf16"),      # cluster of 4, blocks 3 and 4 (almost) empty; one row block per tile
    (60000, 768, 700, 10, O.METRIC_L2, "fp16"),      # cluster of 4, two passes
    (100, 64, 600, 5, O.METRIC_L2, "fp16"),          # fewer tiles than clusters
    (65, 384, 3, 16, O.METRIC_IP, "fp16"),           # a tile whose second row block holds one row
    (8256, 384, 1, 10, O.METRIC_L2, "bf16"),         # odd number of row blocks (129): the last tile is half empty
])
def test_tcgen05_clusters_and_tile_shapes_vs_oracle(P, n, d, nq, k, metric, storage):
    """nq > 128 runs thread-block clusters (TMA multicast of the corpus stages, one 128-query block
    per CTA); pitch <= 512 uses two T64 row blocks per MMA tile.  Every shape must match the oracle."""
    rng = np.random.default_rng(n + d + nq)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    idx = P.FlatIndex(d, metric, storage)
    idx.add(x)
    D, I = idx.search(q, k)
    assert idx.last_path == "tcgen05"
    xs, qs = _round_to(x, storage), _round_to(q, storage)
    S = O.flat_scores_f64(xs, qs, metric)
    for r in range(nq):
        O.check_topk_against_scores(I[r], D[r], S[r], k, metric == O.METRIC_IP, rtol=RTOL_16, atol=2e-5, what=f"q{r}")
    # the same queries one by one (no cluster, different CTA->query mapping) give the same lists
    for r in (0, nq // 2, nq - 1):
        D1, I1 = idx.search(q[r:r + 1], k)
        assert np.array_equal(I1[0], I[r]) and np.array_equal(D1[0], D[r])


@pytest.mark.parametrize("n,d,nq,k,metric,storage", [
    (200000, 768, 64, 100, O.METRIC_IP, "fp16"),
    (200000, 384, 1, 17, O.METRIC_L2, "fp16"),
    (100000, 512, 300, 100, O.METRIC_L2, "bf16"),     # several 128-query passes
    (300000, 384, 5, 1024, O.METRIC_IP, "fp16"),      # k at the API maximum
    (40000, 768, 130, 33, O.METRIC_L2, "fp16"),
])
def test_tcgen05_wide_k_vs_oracle(P, n, d, nq, k, metric, storage):
    """16 < k <= 1024 on 16-bit corpora: sampled per-part lists -> threshold -> collecting scan -> select."""
    rng = np.random.default_rng(n + d + nq + k)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    idx = P.FlatIndex(d, metric, storage)
    idx.add(x)
    D, I = idx.search(q, k)
    assert idx.last_path == "tcgen05"
    xs, qs = _round_to(x, storage), _round_to(q, storage)
    S = O.flat_scores_f64(xs, qs, metric)
    for r in range(0, nq, max(1, nq // 24)):
        O.check_topk_against_scores(I[r], D[r], S[r], k, metric == O.METRIC_IP, rtol=RTOL_16, atol=2e-5, what=f"wide q{r}")


def test_tcgen05_wide_k_pathological_duplicates_stay_exact(P):
    """Tens of thousands of identical rows tie with the threshold: the per-(CTA, query) collection slices
    fill up, their owner threads compact them to their k best and raise the threshold, and the result is
    still exact (ties ordered by ascending row id) -- on the tcgen05 path, with no host synchronisation."""
    rng = np.random.default_rng(3)
    base = rng.standard_normal((120000, 128)).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    x = base.copy()
    x[5000:105000] = x[0]                              # 100 000 copies of row 0: ~675 per CTA against slices of 256
    idx = P.FlatIndex(128, P.METRIC_IP, "fp16")
    idx.add(x)
    for k in (100, 17):
        D, I = idx.search(x[:3], k)
        assert idx.last_path == "tcgen05"
        want = [0] + list(range(5000, 5000 + k - 1))
        assert I[0].tolist() == want and np.allclose(D[0], D[0, 0])
        xs, qs = _round_to(x, "fp16"), _round_to(x[:3], "fp16")
        S = O.flat_scores_f64(xs, qs, O.METRIC_IP)
        for r in (1, 2):
            O.check_topk_against_scores(I[r], D[r], S[r], k, True, rtol=RTOL_16, atol=2e-5, what=f"dups q{r}")
    # L2 as well (the merge re-ranks the selected rows in the direct form: all duplicates at exactly 0)
    l2 = P.FlatIndex(128, P.METRIC_L2, "fp16")
    l2.add(x)
    D, I = l2.search(x[:1], 40)
    assert I[0].tolist() == [0] + list(range(5000, 5039)) and (D[0] == 0).all()


@pytest.mark.parametrize("storage", ["fp16", "bf16"])
@pytest.mark.parametrize("d", [40, 128, 384])
def test_16bit_t64_layout_incremental_add_and_reconstruct(P, storage, d):
    """16-bit corpora live in HBM in the T64 block layout: rows added in ragged pieces (crossing 64-row
    block boundaries, growing the buffer) must reconstruct exactly and search like a single add."""
    rng = np.random.default_rng(d)
    x = rng.standard_normal((333, d)).astype(np.float32)
    a = P.FlatIndex(d, P.METRIC_L2, storage)
    for lo, hi in [(0, 1), (1, 63), (63, 65), (65, 200), (200, 333)]:
        a.add(x[lo:hi])
    b = P.FlatIndex(d, P.METRIC_L2, storage)
    b.add(x)
    xs = _round_to(x, storage)
    assert np.array_equal(a.reconstruct_n(0, 333), xs) and np.array_equal(b.reconstruct_n(5, 100), xs[5:105])
    q = x[[0, 62, 63, 64, 199, 332]] + 0.01
    for path in ("tcgen05", "cuda-core"):
        a.set_path(path); b.set_path(path)
        Da, Ia = a.search(q, 7)
        Db, Ib = b.search(q, 7)
        assert np.array_equal(Ia, Ib) and np.array_equal(Da, Db)
        assert Ia[:, 0].tolist() == [0, 62, 63, 64, 199, 332]


@pytest.mark.parametrize("storage", ["fp16", "bf16"])
def test_16bit_container_roundtrip(P, storage, tmp_path):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((77, 96)).astype(np.float32)
    idx = P.FlatIndex(96, P.METRIC_L2, storage)
    idx.add(x)
    p = str(tmp_path / "h.index")
    P.write_index(idx, p)
    assert open(p, "rb").read()[:4] == (b"PRSh" if storage == "fp16" else b"PRSb")
    assert os.path.getsize(p) == 45 + 2 * 77 * 96
    back = P.read_index(p, storage=storage)
    assert np.array_equal(back.reconstruct_n(), _round_to(x, storage))
    q = rng.standard_normal((3, 96)).astype(np.float32)
    assert np.array_equal(back.search(q, 5)[1], idx.search(q, 5)[1])


@pytest.mark.parametrize("storage", ["fp16", "bf16", "fp32"])
def test_sharded_container_roundtrip_and_mmap(P, storage, tmp_path):
    """f-2: shard files hold the HBM image (T64 for 16-bit) + norms, page aligned; a load is a straight copy."""
    import json
    import struct
    from persian_rag_system_b200 import container as C
    rng = np.random.default_rng(8)
    n, d = 1000, 200                                         # d not a multiple of 64, rows not a multiple of 64
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((9, d)).astype(np.float32)
    whole = P.FlatIndex(d, P.METRIC_L2, storage)
    whole.add(x)
    Dw, Iw = whole.search(q, 7)
    # one shard
    C.write_sharded(whole, str(tmp_path / "one"))
    man = json.load(open(tmp_path / "one" / "manifest.json"))
    assert man["ntotal"] == n and man["storage"] == storage and len(man["shards"]) == 1
    back = C.read_sharded(str(tmp_path / "one"))
    assert back.ntotal == n and back.d == d and back.storage == storage and back.metric_type == P.METRIC_L2
    D, I = back.search(q, 7)
    assert np.array_equal(I, Iw) and np.array_equal(D, Dw)
    assert np.array_equal(back.reconstruct_n(0, n), whole.reconstruct_n(0, n))
    # the file layout: 128-byte header, page-aligned payload and norms; norms are those of the STORED rows
    raw = open(tmp_path / "one" / man["shards"][0]["file"], "rb").read()
    assert raw[:4] == b"PRST"
    ver, dd, pitch, metric, st = struct.unpack_from("<5i", raw, 4)
    rows, rows_pad, id_off, p_off, p_bytes, n_off, n_bytes = struct.unpack_from("<7q", raw, 24)
    es = 4 if storage == "fp32" else 2
    assert (ver, dd, pitch, rows, id_off) == (1, d, 256, n, 0) and p_off % 4096 == 0 and n_off % 4096 == 0
    assert rows_pad == (n if storage == "fp32" else 1024) and p_bytes == rows_pad * pitch * es and n_bytes == 4 * n
    norms = np.memmap(tmp_path / "one" / man["shards"][0]["file"], dtype="<f4", mode="r", offset=n_off, shape=(n,))
    xs = whole.reconstruct_n(0, n).astype(np.float64)
    np.testing.assert_allclose(norms, (xs * xs).sum(1), rtol=1e-5)
    # three shards with global ids (what three ranks would write), loaded by 1 rank (appended) and by rank 1 of 3
    bounds = [(0, 384), (384, 768), (768, n)]                  # non-final shards end on 64-row block boundaries
    os.makedirs(tmp_path / "three")
    for i, (lo, hi) in enumerate(bounds):
        part = P.FlatIndex(d, P.METRIC_L2, storage)
        part.add(x[lo:hi])
        part.set_id_offset(lo)
        C.write_shard(part, str(tmp_path / "three" / f"shard_{i:05d}.prst"))
    json.dump({"format": C.FORMAT, "d": d, "metric": P.METRIC_L2, "storage": storage, "ntotal": n,
               "shards": [{"file": f"shard_{i:05d}.prst", "rows": hi - lo, "id_offset": lo} for i, (lo, hi) in enumerate(bounds)]},
              open(tmp_path / "three" / "manifest.json", "w"))
    allin = C.read_sharded(str(tmp_path / "three"))
    D, I = allin.search(q, 7)
    assert allin.ntotal == n and np.array_equal(I, Iw) and np.array_equal(D, Dw)
    mid = C.read_sharded(str(tmp_path / "three"), rank=1, world=3)
    Dm, Im = mid.search(q, 7)
    ref = P.FlatIndex(d, P.METRIC_L2, storage)
    ref.add(x[384:768])
    Dr, Ir = ref.search(q, 7)
    assert mid.ntotal == 384 and np.array_equal(Im, Ir + 384) and np.array_equal(Dm, Dr)
    # errors: truncated file, append at a non-block boundary, foreign file
    open(tmp_path / "trunc.prst", "wb").write(raw[: len(raw) // 2])
    with pytest.raises(P.PrsError):
        C.read_shard(str(tmp_path / "trunc.prst"))
    if storage != "fp32":
        odd = P.FlatIndex(d, P.METRIC_L2, storage)
        odd.add(x[:100])
        with pytest.raises(P.PrsError):
            C.read_shard(str(tmp_path / "three" / "shard_00001.prst"), into=odd)
    with pytest.raises(P.PrsError):
        C.read_shard(os.path.join(os.path.dirname(__file__), "golden", "indices", "drugs_sentence_chunks.index"))


def test_one_launch_search_equals_three_kernel_sequence(P):
    """The cooperative one-launch search (queries converted by the epilogue threads, grid barrier, CTAs merge the
    queries among themselves) must return bit for bit what prep -> scan -> merge return, search after search
    (the bootstrap words and the grid barrier reset themselves), for fp32 / fp16 / bf16 queries and both metrics."""
    import torch
    dev = torch.device("cuda", 0)
    for (n, d, nq, k, storage, metric) in [(1000, 200, 9, 7, "bf16", P.METRIC_L2), (20000, 768, 64, 10, "fp16", P.METRIC_INNER_PRODUCT),
                                           (100000, 384, 128, 16, "bf16", P.METRIC_L2), (125, 384, 1, 5, "fp16", P.METRIC_L2),
                                           (5000, 64, 33, 3, "fp16", P.METRIC_INNER_PRODUCT), (3000, 100, 5, 4, "fp16", P.METRIC_L2)]:
        rng = np.random.default_rng(n + d)
        x = rng.standard_normal((n, d)).astype(np.float32)
        one = P.FlatIndex(d, metric, storage)
        one.add(x)
        three = P.FlatIndex(d, metric, storage)
        three.add(x)
        three.set_fused(False)
        two = P.FlatIndex(d, metric, storage)                      # scan with its own query preparation + separate merge kernel
        two.add(x)
        two.set_fused(3)
        for rep in range(25):
            qh = rng.standard_normal((nq, d)).astype(np.float32)
            q = torch.from_numpy(qh).to(dev)
            if rep % 3 == 1:
                q = q.half()
            elif rep % 3 == 2:
                q = q.bfloat16()
            D, I = one.search(q, k)
            Dr, Ir = three.search(q, k)
            assert one.last_path == "tcgen05" and not three.last_fused
            assert one.last_fused                                  # (d % 8 != 0 keeps the preparation kernel, the merge is still fused)
            assert torch.equal(I, Ir) and torch.equal(D, Dr), (n, d, nq, k, storage, metric, rep)
            D2, I2 = two.search(q, k)
            assert not two.last_fused and torch.equal(I2, Ir) and torch.equal(D2, Dr), ("two launches", n, d, nq, k, storage, metric, rep)
        # numpy (pageable host) queries take the staged copy and the same kernels
        D, I = one.search(qh, k)
        Dr, Ir = three.search(qh, k)
        assert np.array_equal(I, Ir) and np.array_equal(D, Dr)
    # more than 128 queries or k > 16 are not one launch
    big = P.FlatIndex(64, P.METRIC_INNER_PRODUCT, "fp16")
    big.add(np.random.default_rng(0).standard_normal((40000, 64)).astype(np.float32))
    big.search(np.zeros((129, 64), np.float32), 5)
    assert not big.last_fused
    big.search(np.zeros((4, 64), np.float32), 17)
    assert not big.last_fused
    big.search(np.zeros((4, 64), np.float32), 16)
    assert big.last_fused


def test_search_to_host_equals_search(P):
    """CUDA-tensor queries, results written by the kernels straight into page-locked host memory: same answers as the
    tensor-out search on every path (one launch, clusters, wide k, the fp32 scan, a tiny index)."""
    import torch
    rng = np.random.default_rng(77)
    for (n, d, nq, k, storage) in [(20000, 768, 64, 10, "fp16"), (9000, 384, 200, 5, "bf16"), (40000, 128, 3, 100, "fp16"),
                                   (3000, 384, 2, 7, "fp32"), (125, 384, 1, 5, "fp32")]:
        idx = P.FlatIndex(d, P.METRIC_L2, storage)
        idx.add(rng.standard_normal((n, d)).astype(np.float32))
        q = torch.from_numpy(rng.standard_normal((nq, d)).astype(np.float32)).cuda()
        D, I = idx.search(q, k)
        for _ in range(2):                                   # second call reuses the cached page-locked buffers
            Dh, Ih = idx.search_to_host(q, k)
            assert isinstance(Dh, np.ndarray) and np.array_equal(Ih, I.cpu().numpy()) and np.array_equal(Dh, D.cpu().numpy())


def test_auto_path_selection(P):
    rng = np.random.default_rng(6)
    x = rng.standard_normal((2048, 256)).astype(np.float32)
    idx = P.FlatIndex(256, P.METRIC_IP, "fp16")
    idx.add(x)
    idx.search(rng.standard_normal((1, 256)).astype(np.float32), 10)
    assert idx.last_path == "tcgen05"             # 16-bit corpora stream through the tensor pipe at any batch
    idx.search(rng.standard_normal((64, 256)).astype(np.float32), 10)
    assert idx.last_path == "tcgen05"
    f32 = P.FlatIndex(256, P.METRIC_IP, "fp32")
    f32.add(x)
    f32.search(rng.standard_normal((64, 256)).astype(np.float32), 10)
    assert f32.last_path == "cuda-core"           # fp32 storage is the exact-parity CUDA-core scan
    idx.search(rng.standard_normal((4, 256)).astype(np.float32), 100)
    assert idx.last_path == "cuda-core"           # wide k needs >= 16384 rows for its sampled threshold
    idx.search(rng.standard_normal((64, 256)).astype(np.float32), 100)
    assert idx.last_path == "cuda-core"           # k beyond the in-smem lists


# ------------------------------------------------------------------ device tensors in / out
def test_device_tensor_api_matches_host_api(P):
    import torch
    rng = np.random.default_rng(8)
    x = rng.standard_normal((3000, 384)).astype(np.float32)
    q = rng.standard_normal((12, 384)).astype(np.float32)
    a = P.IndexFlatL2(384)
    a.add(x)
    b = P.IndexFlatL2(384)
    b.add(torch.from_numpy(x).cuda())
    Dh, Ih = a.search(q, 10)
    Dd, Id = b.search(torch.from_numpy(q).cuda(), 10)
    assert Dd.is_cuda and Id.dtype == torch.int64
    assert np.array_equal(Id.cpu().numpy(), Ih) and np.array_equal(Dd.cpu().numpy(), Dh)
    # fp16 query tensors are accepted
    Dq, Iq = b.search(torch.from_numpy(q).cuda().half(), 10)
    S = O.flat_scores_f64(x, torch.from_numpy(q).half().float().numpy(), O.METRIC_L2)
    for r in range(12):
        O.check_topk_against_scores(Iq[r].cpu().numpy(), Dq[r].cpu().numpy(), S[r], 10, False, rtol=RTOL_F32, atol=1e-5)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        D2, I2 = b.search(torch.from_numpy(q).cuda(), 10)
    s.synchronize()
    assert np.array_equal(I2.cpu().numpy(), Ih)


def test_concurrent_readers(P):
    """gradio_luncher.py:361 runs up to 10 threads against one retriever: search must be re-entrant."""
    import threading
    rng = np.random.default_rng(9)
    x = rng.standard_normal((5000, 128)).astype(np.float32)
    qs = rng.standard_normal((10, 4, 128)).astype(np.float32)
    idx = P.IndexFlatL2(128)
    idx.add(x)
    want = [O.flat_search_c(x, qs[t], 5, O.METRIC_L2, form=1)[1] for t in range(10)]
    got = [None] * 10

    def work(t):
        for _ in range(5):
            got[t] = idx.search(qs[t], 5)[1]

    th = [threading.Thread(target=work, args=(t,)) for t in range(10)]
    [t.start() for t in th]
    [t.join() for t in th]
    for t in range(10):
        assert np.array_equal(got[t], want[t])


# ------------------------------------------------------------------ BASELINE sizes: size-independent properties
def test_full_size_1m_x_768_properties(P):
    """configs[1]: 1M x 768.  Generated on device; checked by (a) planted neighbours: a query equal
    to corpus row r must return r first with distance ~0; (b) oracle parity on the 64 queries
    restricted to a 100k-row window around the planted rows via sharded == unsharded reasoning:
    the global top-k restricted to the window equals the window's own top-k."""
    import torch
    n, d, k = 1_000_000, 768, 10
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)
    x = torch.nn.functional.normalize(x, dim=1)
    rows = torch.arange(0, 64, device="cuda") * 15_625 + 7
    q = x[rows].clone()
    for storage, path in (("fp16", "cuda-core"), ("fp16", "tcgen05"), ("bf16", "tcgen05"), ("fp32", "cuda-core")):
        idx = P.FlatIndex(d, P.METRIC_IP, storage)
        idx.add(x)
        idx.set_path(path)
        qq = q if path == "tcgen05" else q[:4]
        D, I = idx.search(qq, k)
        I = I.cpu().numpy()
        D = D.cpu().numpy()
        assert (I[:, 0] == rows[: qq.shape[0]].cpu().numpy()).all(), (storage, path)
        assert np.all(np.abs(D[:, 0] - 1.0) < (1e-2 if storage == "bf16" else 2e-3))
        assert (np.diff(D, axis=1) <= 0).all()
        # window check against the oracle: every returned id inside the first 100k rows must be in
        # the oracle's top-k of that window, in the same relative order (tie-aware)
        xs = x[:100_000].to({"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[storage]).float().cpu().numpy()
        qs = qq.cpu().numpy()
        if path == "tcgen05":
            qs = _round_to(qs, storage)
        S = xs.astype(np.float64) @ qs.astype(np.float64).T
        for r in range(qq.shape[0]):
            inside = I[r][I[r] < 100_000]
            if inside.size == 0:
                continue
            kth = D[r, -1]
            better = np.nonzero(S[:, r] > kth + 2e-3)[0]
            assert set(better.tolist()) <= set(inside.tolist()), (storage, path, r)
        del idx


def test_host_search_with_pinned_buffers_is_zero_copy_and_identical(P):
    """prs_index_search_host with page-locked q / D / I (what bench.py's e2e leg passes) skips the
    staging copies: kernels read the queries from, and write the results to, the caller's pinned memory.
    Results must equal the ordinary pageable-buffer call bit for bit."""
    import torch
    rng = np.random.default_rng(21)
    x = rng.standard_normal((30000, 384)).astype(np.float32)
    q = rng.standard_normal((40, 384)).astype(np.float32)
    for storage, k in (("fp16", 10), ("fp16", 50), ("fp32", 10)):
        idx = P.FlatIndex(384, P.METRIC_L2, storage)
        idx.add(x)
        D0, I0 = idx.search(q, k)                                   # pageable numpy buffers
        qp = torch.from_numpy(q).pin_memory()
        Dp = torch.empty((40, k), dtype=torch.float32).pin_memory()
        Ip = torch.empty((40, k), dtype=torch.int64).pin_memory()
        Dp.fill_(-7.0); Ip.fill_(-7)
        idx.search_into(qp.data_ptr(), 40, k, Dp.data_ptr(), Ip.data_ptr())
        assert np.array_equal(Ip.numpy(), I0) and np.array_equal(Dp.numpy(), D0)
