"""Sparse retrieval (BM25 / TF-IDF) on top of libprs's inverted-index scoring kernel.

Host side = text -> term ids -> CSR weights, restating exactly what the reference's libraries do
before the arithmetic starts; device side (csrc/sparse.cu) = scoring + top-k.

  BM25Index   <- rank_bm25==0.2.2 BM25Okapi(tokenized_chunks)       src/retrieval.py:66-67
                 .get_scores(query_tokens) + argsort[::-1][:k]      src/retrieval.py:127,130
  TfidfIndex  <- TfidfVectorizer(max_features=10000, stop_words=None, ngram_range=(1,2))
                 .fit_transform / .transform + cosine_similarity    src/retrieval.py:78-83,152-159
"""
from __future__ import annotations

import ctypes
import math
import re
from typing import Dict, List, Sequence

import numpy as np

from . import _lib
from ._lib import F32, F64, PrsError, check


class SparseIndex:
    """Doc-by-term CSR weights resident in HBM as term-major postings."""

    def __init__(self, indptr, indices, values, n_terms: int, device: int | None = None, mode: str = "exact"):
        from .flat import _default_device
        self._L = _lib.lib()
        self._h = ctypes.c_void_p()
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        values = np.ascontiguousarray(values)
        if values.dtype == np.float32:
            vd = F32
        elif values.dtype == np.float64:
            vd = F64
        else:
            raise PrsError(_lib.EINVAL, f"values must be float32 or float64, got {values.dtype}")
        if indptr.ndim != 1 or indptr.shape[0] < 1 or indices.shape[0] != values.shape[0] or indptr[-1] != indices.shape[0]:
            raise PrsError(_lib.EINVAL, "malformed CSR")
        self.device = _default_device() if device is None else int(device)
        self.n_terms = int(n_terms)
        check(self._L.prs_sparse_build(indptr.ctypes.data_as(ctypes.c_void_p), indices.ctypes.data_as(ctypes.c_void_p),
                                       values.ctypes.data_as(ctypes.c_void_p), vd, int(indptr.shape[0] - 1), int(n_terms),
                                       self.device, ctypes.byref(self._h)))
        self.set_mode(mode)

    def set_mode(self, mode: str) -> None:
        """"exact": float64 accumulation in query-entry order (rank_bm25 / scipy summation order), one
        query per CTA.  "throughput": 8 queries per CTA share the postings they have in common,
        fixed-point accumulation, k + 16 candidates.  Both re-score their candidates exactly in float64
        (identical scores); id lists can differ only among docs whose scores tie to ~1e-7 relative.
        "auto": exact order below 16 queries per call, throughput from there on."""
        code = {"exact": 0, "throughput": 1, "fast": 1, "auto": 2, 0: 0, 1: 1, 2: 2}[mode]
        check(self._L.prs_sparse_set_mode(self._h, code))

    @property
    def mode(self) -> str:
        return {0: "exact", 1: "throughput", 2: "auto"}[int(self._L.prs_sparse_mode(self._h))]

    @property
    def ndocs(self) -> int:
        return int(self._L.prs_sparse_ndocs(self._h))

    @property
    def nnz(self) -> int:
        return int(self._L.prs_sparse_nnz(self._h))

    @property
    def last_postings(self) -> int:
        return int(self._L.prs_sparse_last_postings(self._h))

    def set_id_offset(self, offset: int) -> None:
        check(self._L.prs_sparse_set_id_offset(self._h, int(offset)))

    def search(self, q_indptr, q_terms, q_weights, k: int):
        """Queries as CSR (entry order is the float64 summation order).  Returns (S float64 [nq,k],
        I int64 [nq,k]) ordered (score desc, doc id desc); zero-score docs are eligible."""
        q_indptr = np.ascontiguousarray(q_indptr, dtype=np.int64)
        q_terms = np.ascontiguousarray(q_terms, dtype=np.int32)
        q_weights = np.ascontiguousarray(q_weights, dtype=np.float64)
        nq = int(q_indptr.shape[0] - 1)
        S = np.zeros((nq, int(k)), dtype=np.float64)
        I = np.full((nq, int(k)), -1, dtype=np.int64)
        # grid.y limit: 65535 queries per call
        for a in range(0, nq, 32768):
            b = min(nq, a + 32768)
            ip = np.ascontiguousarray(q_indptr[a:b + 1] - q_indptr[a])
            e0, e1 = int(q_indptr[a]), int(q_indptr[b])
            qt = np.ascontiguousarray(q_terms[e0:e1])
            qw = np.ascontiguousarray(q_weights[e0:e1])
            Sa = np.empty((b - a, int(k)), dtype=np.float64)
            Ia = np.empty((b - a, int(k)), dtype=np.int64)
            check(self._L.prs_sparse_search_host(self._h, ip.ctypes.data_as(ctypes.c_void_p), qt.ctypes.data_as(ctypes.c_void_p),
                                                 qw.ctypes.data_as(ctypes.c_void_p), b - a, int(k),
                                                 Sa.ctypes.data_as(ctypes.c_void_p), Ia.ctypes.data_as(ctypes.c_void_p)))
            S[a:b], I[a:b] = Sa, Ia
        return S, I

    def search_device(self, q_indptr, q_terms, q_weights, k: int):
        """Same search with the query CSR given as CUDA tensors (int64 / int32 / float64) and the results
        left on the device: (S float64 [nq, k], I int64 [nq, k]) CUDA tensors, asynchronous on the current
        stream.  What the hybrid fusion kernel consumes (at most 65 535 queries per call)."""
        import torch
        dev = q_indptr.device
        q_indptr = q_indptr.to(torch.int64).contiguous()
        q_terms = q_terms.to(torch.int32).contiguous()
        q_weights = q_weights.to(torch.float64).contiguous()
        nq = int(q_indptr.shape[0] - 1)
        S = torch.empty((nq, int(k)), dtype=torch.float64, device=dev)
        I = torch.empty((nq, int(k)), dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        check(self._L.prs_sparse_search_device(self._h, ctypes.c_void_p(q_indptr.data_ptr()), ctypes.c_void_p(q_terms.data_ptr()),
                                               ctypes.c_void_p(q_weights.data_ptr()), nq, int(k), ctypes.c_void_p(S.data_ptr()),
                                               ctypes.c_void_p(I.data_ptr()), ctypes.c_void_p(st)))
        return S, I

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._L.prs_sparse_free(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------
# BM25 (rank_bm25 0.2.2 BM25Okapi semantics)
# --------------------------------------------------------------------------------------------
def build_bm25_csr(corpus: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
    """Host build of the BM25 doc-by-term weight matrix (CSR, float64), rank_bm25 0.2.2 semantics.

    idf(t) = ln(N - df + 0.5) - ln(df + 0.5); negative idfs are replaced by epsilon * mean(idf);
    weight = idf * tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl)) -- same operations, same order as
    `BM25Okapi._calc_idf` / `get_scores`, so the float64 values are rank_bm25's bit for bit.
    Returns dict(indptr, indices, weights, vocab, idf, doc_len, avgdl, average_idf)."""
    corpus_size = len(corpus)
    if corpus_size == 0:
        raise ZeroDivisionError("division by zero")          # rank_bm25 divides by corpus_size
    vocab: Dict[str, int] = {}
    df: List[int] = []
    doc_len = np.empty(corpus_size, dtype=np.int64)
    indptr = np.zeros(corpus_size + 1, dtype=np.int64)
    ind_chunks, tf_chunks = [], []
    for i, document in enumerate(corpus):
        doc_len[i] = len(document)
        freqs: Dict[int, int] = {}
        for word in document:
            tid = vocab.get(word)
            if tid is None:
                tid = len(vocab)
                vocab[word] = tid
                df.append(0)
            freqs[tid] = freqs.get(tid, 0) + 1
        for tid in freqs:
            df[tid] += 1
        ind_chunks.append(np.fromiter(freqs.keys(), dtype=np.int32, count=len(freqs)))
        tf_chunks.append(np.fromiter(freqs.values(), dtype=np.int64, count=len(freqs)))
        indptr[i + 1] = indptr[i] + len(freqs)
    avgdl = int(doc_len.sum()) / corpus_size
    # _calc_idf: accumulation order = vocabulary insertion order, as in rank_bm25
    idf = np.empty(len(vocab), dtype=np.float64)
    idf_sum = 0
    negative = []
    for tid, freq in enumerate(df):
        v = math.log(corpus_size - freq + 0.5) - math.log(freq + 0.5)
        idf[tid] = v
        idf_sum += v
        if v < 0:
            negative.append(tid)
    average_idf = idf_sum / len(vocab) if vocab else 0.0
    eps = epsilon * average_idf
    for tid in negative:
        idf[tid] = eps
    indices = np.concatenate(ind_chunks) if ind_chunks else np.empty(0, np.int32)
    tf = np.concatenate(tf_chunks) if tf_chunks else np.empty(0, np.int64)
    dl = np.repeat(doc_len, np.diff(indptr))
    # same expression, same evaluation order as BM25Okapi.get_scores
    w = idf[indices] * (tf * (k1 + 1) / (tf + k1 * (1 - b + b * dl / avgdl)))
    return dict(indptr=indptr, indices=indices, weights=w, vocab=vocab, idf=idf, doc_len=doc_len, avgdl=avgdl,
                average_idf=average_idf)


class BM25Index:
    """`BM25Okapi(tokenized_corpus)` with the scoring on the GPU.  `dtype="float64"` keeps
    rank_bm25's exact float64 weights (12 B per posting); `dtype="float32"` stores them in fp32
    (8 B per posting, throughput mode)."""

    def __init__(self, corpus: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25,
                 dtype="float64", device: int | None = None, mode: str = "exact"):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        built = build_bm25_csr(corpus, k1, b, epsilon)
        self.corpus_size = len(corpus)
        self.vocab, self.idf, self.doc_len = built["vocab"], built["idf"], built["doc_len"]
        self.avgdl, self.average_idf = built["avgdl"], built["average_idf"]
        self.dtype = np.dtype(dtype)
        self.index = SparseIndex(built["indptr"], built["indices"], built["weights"].astype(self.dtype), len(self.vocab), device, mode)

    def encode_queries(self, queries: Sequence[Sequence[str]]):
        q_indptr = np.zeros(len(queries) + 1, dtype=np.int64)
        terms: List[int] = []
        for i, toks in enumerate(queries):
            for t in toks:                       # repetition kept: repeated tokens add repeatedly
                terms.append(self.vocab.get(t, -1))
            q_indptr[i + 1] = len(terms)
        return q_indptr, np.asarray(terms, dtype=np.int32), np.ones(len(terms), dtype=np.float64)

    def search(self, queries: Sequence[Sequence[str]], k: int):
        """Batch of tokenised queries -> (scores float64 [nq,k], doc ids int64 [nq,k])."""
        return self.index.search(*self.encode_queries(queries), k)

    def search_device(self, queries: Sequence[Sequence[str]], k: int):
        """Same, results left on the device as CUDA tensors (hybrid fusion input)."""
        import torch
        dev = torch.device("cuda", self.index.device)
        ip, qt, qw = self.encode_queries(queries)
        return self.index.search_device(torch.from_numpy(ip).to(dev), torch.from_numpy(qt).to(dev), torch.from_numpy(qw).to(dev), k)

    def get_top_k(self, query_tokens: Sequence[str], k: int):
        S, I = self.search([query_tokens], k)
        valid = I[0] >= 0
        return S[0][valid], I[0][valid]


# --------------------------------------------------------------------------------------------
# TF-IDF (scikit-learn TfidfVectorizer + cosine_similarity semantics)
# --------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"(?u)\b\w\w+\b")


def _analyze(text: str, ngram_range=(1, 2)) -> List[str]:
    """TfidfVectorizer's default analyzer: lowercase, token_pattern (?u)\\b\\w\\w+\\b, word n-grams."""
    tokens = _TOKEN.findall(text.lower())
    lo, hi = ngram_range
    if hi == 1:
        return tokens
    original = tokens
    out = list(original) if lo == 1 else []
    n0 = max(lo, 2)
    n_orig = len(original)
    for n in range(n0, min(hi + 1, n_orig + 1)):
        for i in range(n_orig - n + 1):
            out.append(" ".join(original[i:i + n]))
    return out


def _row_l2_normalize(indptr: np.ndarray, data: np.ndarray) -> None:
    """sklearn inplace_csr_row_normalize_l2: sequential sum of squares per row, sqrt, divide."""
    for r in range(indptr.shape[0] - 1):
        a, b = int(indptr[r]), int(indptr[r + 1])
        s = 0.0
        for v in data[a:b].tolist():
            s += v * v
        if s == 0.0:
            continue
        data[a:b] /= math.sqrt(s)


class TfidfVectorizerHost:
    """Host restatement of `TfidfVectorizer(max_features, stop_words=None, ngram_range)` as the
    reference configures it (src/retrieval.py:78-82; other parameters at scikit-learn defaults:
    lowercase, smooth idf, l2 norm, float64).  `fit` returns the doc-by-term CSR AFTER the second
    row normalisation that `cosine_similarity` applies (src/retrieval.py:156)."""

    def __init__(self, max_features: int | None = 10000, ngram_range=(1, 2)):
        self.max_features, self.ngram_range = max_features, ngram_range

    def fit(self, texts: Sequence[str]):
        n_docs = len(texts)
        vocab: Dict[str, int] = {}
        rows = []
        for doc in texts:
            counts: Dict[int, int] = {}
            for feat in _analyze(doc, self.ngram_range):
                fid = vocab.get(feat)
                if fid is None:
                    fid = len(vocab)
                    vocab[feat] = fid
                counts[fid] = counts.get(fid, 0) + 1
            rows.append(counts)
        if not vocab:
            raise ValueError("empty vocabulary; perhaps the documents only contain stop words")
        # CountVectorizer._sort_features: alphabetical feature order
        sorted_feats = sorted(vocab.items())
        remap = np.empty(len(vocab), dtype=np.int64)
        for new, (_term, old) in enumerate(sorted_feats):
            remap[old] = new
        terms = [t for t, _ in sorted_feats]
        n_feat = len(terms)
        dfs = np.zeros(n_feat, dtype=np.int64)
        tfs = np.zeros(n_feat, dtype=np.int64)
        for counts in rows:
            for old, c in counts.items():
                dfs[remap[old]] += 1
                tfs[remap[old]] += c
        # CountVectorizer._limit_features (max_df=1.0, min_df=1, max_features)
        mask = (dfs <= n_docs) & (dfs >= 1)
        if self.max_features is not None and int(mask.sum()) > self.max_features:
            mask_inds = (-tfs[mask]).argsort()[:self.max_features]
            new_mask = np.zeros(n_feat, dtype=bool)
            new_mask[np.where(mask)[0][mask_inds]] = True
            mask = new_mask
        new_indices = np.cumsum(mask) - 1
        self.vocabulary_ = {t: int(new_indices[i]) for i, t in enumerate(terms) if mask[i]}
        kept = np.where(mask)[0]
        self.n_features = int(kept.shape[0])
        df_kept = dfs[kept].astype(np.float64)
        # TfidfTransformer.fit (smooth_idf=True): idf = ln((1 + n) / (1 + df)) + 1
        self.idf_ = np.log((n_docs + 1) / (df_kept + 1)) + 1
        # Row entries are kept in the order scikit-learn holds them while it normalises: ascending
        # FIRST-SEEN feature id (CountVectorizer._count_vocab sorts by those ids; _sort_features then
        # only relabels).  The sum of squares is accumulated in that order -- the values only depend
        # on the order through the rounding of that sum -- and the CSR is sorted by final id afterwards.
        indptr = np.zeros(n_docs + 1, dtype=np.int64)
        ind_chunks, val_chunks = [], []
        for i, counts in enumerate(rows):
            olds = sorted(o for o in counts if mask[remap[o]])
            ids = np.array([new_indices[remap[o]] for o in olds], dtype=np.int64)
            cs = np.array([counts[o] for o in olds], dtype=np.float64)
            ind_chunks.append(ids)
            val_chunks.append(cs * self.idf_[ids])
            indptr[i + 1] = indptr[i] + ids.shape[0]
        indices = np.concatenate(ind_chunks) if ind_chunks else np.empty(0, np.int64)
        data = np.concatenate(val_chunks) if val_chunks else np.empty(0, np.float64)
        _row_l2_normalize(indptr, data)       # TfidfTransformer norm='l2'
        tfidf_data = data.copy()
        _row_l2_normalize(indptr, data)       # cosine_similarity normalises its inputs again
        for r in range(n_docs):               # sorted CSR indices
            a, b = int(indptr[r]), int(indptr[r + 1])
            order = np.argsort(indices[a:b], kind="stable")
            indices[a:b] = indices[a:b][order]
            data[a:b] = data[a:b][order]
            tfidf_data[a:b] = tfidf_data[a:b][order]
        self.tfidf_data_ = tfidf_data         # == TfidfVectorizer.fit_transform(texts) (sorted indices)
        indices = indices.astype(np.int32)
        return indptr, indices, data

    def transform_query(self, query: str):
        counts: Dict[int, int] = {}
        for feat in _analyze(query, self.ngram_range):
            fid = self.vocabulary_.get(feat)
            if fid is not None:
                counts[fid] = counts.get(fid, 0) + 1
        ids = np.array(sorted(counts), dtype=np.int64)
        vals = np.array([counts[i] for i in ids.tolist()], dtype=np.float64) * self.idf_[ids] if ids.size else np.empty(0)
        ip = np.array([0, ids.shape[0]], dtype=np.int64)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        _row_l2_normalize(ip, vals)
        _row_l2_normalize(ip, vals)
        return ids.astype(np.int32), vals

    def encode_queries(self, queries: Sequence[str]):
        q_indptr = np.zeros(len(queries) + 1, dtype=np.int64)
        ts, ws = [], []
        for i, q in enumerate(queries):
            t, w = self.transform_query(q)
            ts.append(t)
            ws.append(w)
            q_indptr[i + 1] = q_indptr[i] + t.shape[0]
        q_terms = np.concatenate(ts) if ts else np.empty(0, np.int32)
        q_w = np.concatenate(ws) if ws else np.empty(0, np.float64)
        return q_indptr, q_terms, q_w


class TfidfIndex(TfidfVectorizerHost):
    """`TfidfVectorizer(...).fit_transform(texts)` + `cosine_similarity(query_vector, matrix)` +
    argsort top-k, with scoring and selection on the GPU."""

    def __init__(self, texts: Sequence[str], max_features: int | None = 10000, ngram_range=(1, 2), dtype="float64",
                 device: int | None = None, mode: str = "exact"):
        super().__init__(max_features, ngram_range)
        indptr, indices, data = self.fit(texts)
        self.dtype = np.dtype(dtype)
        self.index = SparseIndex(indptr, indices, data.astype(self.dtype), self.n_features, device, mode)

    def search(self, queries: Sequence[str], k: int):
        return self.index.search(*self.encode_queries(queries), k)

    def get_top_k(self, query: str, k: int):
        S, I = self.search([query], k)
        valid = I[0] >= 0
        return S[0][valid], I[0][valid]
