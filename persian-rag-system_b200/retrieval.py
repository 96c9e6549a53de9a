"""Drop-in for the reference's retriever module (`src/retrieval.py`), B200 engine underneath.

Same class names, constructor arguments, method names, return types and error behaviour as
`RetrievalSystem` (src/retrieval.py:12-336) and `MultiModelRetrieval` (:339-389); the third-party
calls at the bottom are replaced:

  faiss.read_index / index.search           -> flat.read_index / FlatIndex.search   (csrc/flat_*.cu*)
  BM25Okapi(...) / get_scores + argsort      -> sparse.BM25Index.search              (csrc/sparse.cu)
  TfidfVectorizer + cosine_similarity + argsort -> sparse.TfidfIndex.search           (csrc/sparse.cu)

Every public method keeps the reference's "print and return [] / False" convention on failure
(src/retrieval.py:47-49,57-59,113-115,141-143,170-172,218-220).  There is no CPU fallback: if
libprs or a B200 is missing, loading fails (returns False after printing why).
"""
from __future__ import annotations

import functools
import gc
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .flat import read_index
from .sparse import BM25Index, TfidfIndex


def _cuda_available() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def _load_chunks_csv(chunk_file: str) -> List[Dict[str, Any]]:
    """`pd.read_csv(chunk_file, encoding='utf-8').to_dict('records')` (src/retrieval.py:44-45)."""
    import pandas as pd
    return pd.read_csv(chunk_file, encoding="utf-8").to_dict("records")


def _guarded(label: str, on_error):
    """The reference wraps every public method in `try/except -> print -> return [] / False`
    (src/retrieval.py:47-49,57-59,113-115,141-143,170-172,218-220).  One decorator keeps that
    contract here; the C ABI reports failures as PrsError, which lands in the same net."""
    def wrap(fn):
        @functools.wraps(fn)
        def inner(self, *args, **kwargs):
            try:
                return fn(self, *args, **kwargs)
            except Exception as exc:                       # noqa: BLE001 - the reference catches everything
                print(f"Error in {label}: {exc}")
                return on_error() if callable(on_error) else on_error
        return inner
    return wrap


_METHODS = ("dense", "bm25", "tfidf", "hybrid")


class RetrievalSystem:
    """`RetrievalSystem(method, model_path, device)` of src/retrieval.py:12-36 with the engine knobs
    added as keyword extensions:

    encoder -- any object with `.encode(list[str], device=...) -> ndarray`; used instead of loading
               `SentenceTransformer(model_path)` (the encoders themselves are unchanged, north-star);
    storage -- how the dense corpus is held in HBM: "fp32" (reference parity), "fp16", "bf16";
    devices -- GPUs the dense corpus is split over inside this ONE process (the reference and its
               Gradio app are single-process): contiguous row blocks, results identical to one device.
    """

    def __init__(self, method="dense", model_path=None, device=None, encoder=None, storage="fp32", devices=None):
        self.method, self.storage, self.devices = method, storage, devices
        self.device = device or ("cuda" if _cuda_available() else "cpu")
        self.embedding_model = encoder
        if encoder is None and method in ("dense", "hybrid") and model_path:
            print(f"Loading embedding model: {model_path}")
            from sentence_transformers import SentenceTransformer
            self.embedding_model = SentenceTransformer(model_path, device=self.device)
        # attributes callers of the reference touch (src/evaluation.py:262-263, gradio launcher)
        self.chunks = self.faiss_index = self.bm25_index = None
        self.tfidf_vectorizer = self.tfidf_matrix = None
        self.is_ready = False

    # ---------------------------------------------------------------- loading, src/retrieval.py:38-90
    def load_chunks_and_index(self, chunk_file: str, faiss_index_file: str = None):
        print(f"Loading chunks from {chunk_file}...")
        try:
            records = _load_chunks_csv(chunk_file)
        except Exception as exc:                            # noqa: BLE001
            print(f"Error loading chunks: {exc}")
            return False
        print(f"✓ Loaded {len(records)} chunks")
        return self.load_chunks(records, faiss_index_file)

    def load_chunks(self, chunks: List[Dict[str, Any]], faiss_index_file: str = None):
        """Extension: the chunk records are already in memory (row i of the index <-> chunks[i])."""
        self.chunks = list(chunks)
        wants_dense = self.method in ("dense", "hybrid")
        # a missing index file is skipped silently, exactly like the reference (:52); dense retrieval
        # then answers [] (:94-95)
        steps = []
        if wants_dense and faiss_index_file and os.path.exists(faiss_index_file):
            steps.append(("FAISS index", lambda: self._open_dense(faiss_index_file)))
        if self.method in ("bm25", "hybrid"):
            steps.append(("BM25 index", self._build_bm25))
        if self.method in ("tfidf", "hybrid"):
            steps.append(("TF-IDF index", self._build_tfidf))
        for what, build in steps:
            try:
                build()
            except Exception as exc:                        # noqa: BLE001
                print(f"Error building/loading {what}: {exc}")
                return False
        self.is_ready = True
        return True

    def _open_dense(self, path):
        self.faiss_index = read_index(path, storage=self.storage, devices=self.devices)
        where = f"{len(self.devices)} GPUs" if self.devices and len(self.devices) > 1 else "HBM"
        print(f"✓ Loaded FAISS index with {self.faiss_index.ntotal} vectors into {where} ({self.storage})")

    def _build_bm25(self):
        # whitespace tokens, no lower-casing, no stop words: src/retrieval.py:66
        # "auto": single queries are scored in rank_bm25's summation order, batches of >= 16 queries by the throughput kernel
        self.bm25_index = BM25Index([record["text"].split() for record in self.chunks], mode="auto")
        print("✓ BM25 postings on device")

    def _build_tfidf(self):
        # TfidfVectorizer(max_features=10000, stop_words=None, ngram_range=(1, 2)): src/retrieval.py:78-83
        self.tfidf_vectorizer = TfidfIndex([record["text"] for record in self.chunks],
                                           max_features=10000, ngram_range=(1, 2), mode="auto")
        self.tfidf_matrix = self.tfidf_vectorizer.index
        print("✓ TF-IDF postings on device")

    def _rows_to_results(self, rows, scores):
        """(row id, score) pairs -> the reference's list[(chunk_dict, score)], dropping ids outside
        [0, len(chunks)) as src/retrieval.py:106 does."""
        n = len(self.chunks)
        return [(self.chunks[int(r)], s) for r, s in zip(rows, scores) if 0 <= int(r) < n]

    # ---------------------------------------------------------------- encoder hand-off
    def _embed(self, texts: List[str]):
        """Query embeddings the way the engine wants them: a CUDA tensor whenever the encoder can give one
        (`encode_device` of pooling.FusedPoolingEncoder, or SentenceTransformer's `convert_to_tensor=True`),
        so that the reference's device -> host -> device round trip at src/retrieval.py:98-102 disappears;
        otherwise whatever `encode` returns (numpy for the stock call)."""
        enc = self.embedding_model
        if hasattr(enc, "encode_device"):
            return enc.encode_device(list(texts))
        try:
            return enc.encode(list(texts), device=self.device, convert_to_tensor=True)
        except TypeError:                                   # an encoder with the bare reference signature
            return enc.encode(list(texts), device=self.device)

    @staticmethod
    def _on_device(x) -> bool:
        return hasattr(x, "is_cuda") and bool(x.is_cuda)

    def _dense_search(self, texts: List[str], k: int, want_device: bool = False):
        """One encoder call and ONE corpus scan for all `texts`.  Embeddings that are already on the GPU go
        straight into the scan kernel; only the [nq, k] results come back (numpy), unless `want_device`."""
        emb = self._embed(texts)
        if self._on_device(emb) or want_device:
            import torch
            if not self._on_device(emb):
                if hasattr(emb, "detach"):
                    emb = emb.detach().cpu().numpy()
                emb = torch.from_numpy(np.ascontiguousarray(np.asarray(emb), dtype=np.float32)).to(f"cuda:{self.faiss_index.device}")
            if emb.dim() == 1:
                emb = emb[None, :]
            if emb.dtype not in (torch.float32, torch.float16, torch.bfloat16):
                emb = emb.float()
            if not want_device and hasattr(self.faiss_index, "search_to_host"):
                return self.faiss_index.search_to_host(emb, k)      # results land in page-locked host memory, no D2H copies
            D, I = self.faiss_index.search(emb, k)
            return (D, I) if want_device else (D.cpu().numpy(), I.cpu().numpy())
        if hasattr(emb, "detach"):
            emb = emb.detach().cpu().numpy()
        return self.faiss_index.search(np.asarray(emb).astype("float32"), k)

    # ---------------------------------------------------------------- dense, src/retrieval.py:92-115
    @_guarded("dense retrieval", list)
    def retrieve_dense(self, query: str, top_k: int = 10) -> List[Tuple[Dict, float]]:
        if not self.embedding_model or not self.faiss_index:
            return []
        sq_l2, rows = self._dense_search([query], top_k)
        return self._rows_to_results(rows[0], 1 / (1 + sq_l2[0]))      # score = 1/(1+d), :108

    @_guarded("batched dense retrieval", list)
    def retrieve_dense_batch(self, queries: List[str], top_k: int = 10) -> List[List[Tuple[Dict, float]]]:
        """Extension: one encoder call and ONE corpus scan for the whole list of
        queries -- the reference loops `retrieve` per question (src/evaluation.py:273-299)."""
        if not self.embedding_model or not self.faiss_index or not queries:
            return [[] for _ in queries]
        sq_l2, rows = self._dense_search(list(queries), top_k)
        return [self._rows_to_results(rows[i], 1 / (1 + sq_l2[i])) for i in range(len(queries))]

    # ---------------------------------------------------------------- sparse, src/retrieval.py:117-172
    @_guarded("BM25 retrieval", list)
    def retrieve_bm25(self, query: str, top_k: int = 10) -> List[Tuple[Dict, float]]:
        if not self.bm25_index:
            return []
        scores, rows = self.bm25_index.get_top_k(query.split(), top_k)
        return self._rows_to_results(rows, scores)

    @_guarded("batched BM25 retrieval", list)
    def retrieve_bm25_batch(self, queries: List[str], top_k: int = 10) -> List[List[Tuple[Dict, float]]]:
        """Extension: every query of the batch in one scoring launch."""
        if not self.bm25_index or not queries:
            return [[] for _ in queries]
        S, I = self.bm25_index.search([q.split() for q in queries], top_k)
        return [self._rows_to_results(I[i][I[i] >= 0], S[i][I[i] >= 0]) for i in range(len(queries))]

    @_guarded("TF-IDF retrieval", list)
    def retrieve_tfidf(self, query: str, top_k: int = 10) -> List[Tuple[Dict, float]]:
        if not self.tfidf_vectorizer or self.tfidf_matrix is None:
            return []
        scores, rows = self.tfidf_vectorizer.get_top_k(query, top_k)
        return self._rows_to_results(rows, scores)

    @_guarded("batched TF-IDF retrieval", list)
    def retrieve_tfidf_batch(self, queries: List[str], top_k: int = 10) -> List[List[Tuple[Dict, float]]]:
        if not self.tfidf_vectorizer or self.tfidf_matrix is None or not queries:
            return [[] for _ in queries]
        S, I = self.tfidf_vectorizer.search(list(queries), top_k)
        return [self._rows_to_results(I[i][I[i] >= 0], S[i][I[i] >= 0]) for i in range(len(queries))]

    # ---------------------------------------------------------------- hybrid, src/retrieval.py:174-220
    @_guarded("hybrid retrieval", list)
    def retrieve_hybrid(self, query: str, top_k: int = 10, dense_weight: float = 0.6,
                        bm25_weight: float = 0.4) -> List[Tuple[Dict, float]]:
        out = self.retrieve_hybrid_batch([query], top_k, dense_weight, bm25_weight)
        return out[0] if out else []

    @_guarded("batched hybrid retrieval", list)
    def retrieve_hybrid_batch(self, queries: List[str], top_k: int = 10, dense_weight: float = 0.6,
                              bm25_weight: float = 0.4) -> List[List[Tuple[Dict, float]]]:
        """The reference's fusion (top-2k of each retriever, each list divided by its own maximum, 0.6 / 0.4
        weighted sum keyed by chunk, stable descending sort, src/retrieval.py:181-216) for a whole batch:
        both top-2k lists stay in HBM and one kernel (csrc/hybrid.cu) fuses every query; only the final
        [nq, k] lists are copied back.  A retriever that is missing or fails contributes an empty list,
        exactly like the reference's guarded sub-calls."""
        import torch
        from .hybrid import hybrid_fuse
        if not queries:
            return []
        nq, k2 = len(queries), 2 * int(top_k)
        dev = torch.device("cuda", self.faiss_index.device if self.faiss_index else self.bm25_index.index.device)
        D = torch.empty((nq, 0), dtype=torch.float32, device=dev)
        Id = torch.empty((nq, 0), dtype=torch.int64, device=dev)
        S = torch.empty((nq, 0), dtype=torch.float64, device=dev)
        Is = torch.empty((nq, 0), dtype=torch.int64, device=dev)
        if self.embedding_model and self.faiss_index:
            try:
                D, Id = self._dense_search(list(queries), k2, want_device=True)
            except Exception as exc:                        # noqa: BLE001 - retrieve_dense's own net (:113-115)
                print(f"Error in dense retrieval: {exc}")
        if self.bm25_index:
            try:
                S, Is = self.bm25_index.search_device([q.split() for q in queries], k2)
            except Exception as exc:                        # noqa: BLE001 - retrieve_bm25's own net (:141-143)
                print(f"Error in BM25 retrieval: {exc}")
        fused, rows = hybrid_fuse(D, Id, S, Is, len(self.chunks), top_k, dense_weight, bm25_weight)
        fused, rows = fused.cpu().numpy(), rows.cpu().numpy()
        return [[(self.chunks[int(r)], float(v)) for r, v in zip(rows[i], fused[i]) if r >= 0] for i in range(nq)]

    # ---------------------------------------------------------------- dispatcher, src/retrieval.py:222-238
    def retrieve(self, query: str, top_k: int = 10) -> List[Tuple[Dict, float]]:
        if not self.is_ready:
            print("Retrieval system is not ready. Please load chunks and index first.")
            return []
        if self.method not in _METHODS:
            print(f"Unknown retrieval method: {self.method}")
            return []
        return getattr(self, f"retrieve_{self.method}")(query, top_k)

    def retrieve_batch(self, queries: List[str], top_k: int = 10) -> List[List[Tuple[Dict, float]]]:
        """Extension: `retrieve` for a list of queries with ONE pass of the engine (one corpus scan / one
        scoring launch for all of them).  Element i equals `retrieve(queries[i], top_k)`."""
        queries = list(queries)
        if not self.is_ready:
            print("Retrieval system is not ready. Please load chunks and index first.")
            return [[] for _ in queries]
        if self.method not in _METHODS:
            print(f"Unknown retrieval method: {self.method}")
            return [[] for _ in queries]
        out = getattr(self, f"retrieve_{self.method}_batch")(queries, top_k)
        return out if len(out) == len(queries) else [[] for _ in queries]     # a failed batch == every query failed

    # ---------------------------------------------------------------- RAG contexts, src/retrieval.py:240-272
    @staticmethod
    def _pack_contexts(retrieved, max_context_length: int) -> Tuple[List[str], List[Dict]]:
        contexts, metadata, budget = [], [], max_context_length
        for chunk, score in retrieved:
            text = chunk["text"]
            if len(text) > budget:                 # does not fit: keep a truncated tail only if > 100 chars remain
                if budget <= 100:
                    break
                text = text[:budget] + "..."
            contexts.append(text)
            metadata.append(dict(chunk_id=chunk["id"], score=score,
                                 chunk_type=chunk.get("chunk_type", "unknown"), length=len(text)))
            budget -= len(text)
            if budget <= 0:
                break
        return contexts, metadata

    def get_contexts_for_rag(self, query: str, top_k: int = 5,
                             max_context_length: int = 2000) -> Tuple[List[str], List[Dict]]:
        return self._pack_contexts(self.retrieve(query, top_k), max_context_length)

    def get_contexts_for_rag_batch(self, queries: List[str], top_k: int = 5,
                                   max_context_length: int = 2000) -> List[Tuple[List[str], List[Dict]]]:
        """Extension: the evaluator's per-question loop (src/evaluation.py:273-299) as one
        batched retrieval; element i equals `get_contexts_for_rag(queries[i], ...)`."""
        return [self._pack_contexts(hits, max_context_length) for hits in self.retrieve_batch(queries, top_k)]

    # ---------------------------------------------------------------- Hit@K / MRR, src/retrieval.py:274-323
    def evaluate_retrieval_quality(self, test_queries: List[Dict],
                                   relevant_chunks: Dict[str, List[str]]) -> Dict[str, float]:
        cuts = (1, 3, 5)
        hit_flags = {c: [] for c in cuts}
        reciprocal_ranks = []
        # queries without labels are skipped (:291-292); all the others are retrieved in ONE batch
        labelled = [(item, relevant_chunks.get(item.get("id", str(pos)), [])) for pos, item in enumerate(test_queries)]
        labelled = [(item, wanted) for item, wanted in labelled if wanted]
        batch = self.retrieve_batch([item["question"] for item, _ in labelled], top_k=10) if labelled else []
        for (item, wanted), hits in zip(labelled, batch):
            got = [chunk["id"] for chunk, _ in hits]
            for c in cuts:
                hit_flags[c].append(any(cid in wanted for cid in got[:c]))
            first = next((rank for rank, cid in enumerate(got, 1) if cid in wanted), None)
            reciprocal_ranks.append(1.0 / first if first else 0.0)
        report = {f"hit_at_{c}": (np.mean(hit_flags[c]) if hit_flags[c] else 0.0) for c in cuts}
        report["mrr"] = np.mean(reciprocal_ranks) if reciprocal_ranks else 0.0
        report["total_queries"] = len(test_queries)
        print("✓ Retrieval evaluation: " + ", ".join(f"Hit@{c}={report[f'hit_at_{c}']:.3f}" for c in cuts)
              + f", MRR={report['mrr']:.3f}")
        return report

    # ---------------------------------------------------------------- cleanup, src/retrieval.py:325-336
    def cleanup(self):
        """Drop the encoder, the device index and the chunks (frees the HBM copies)."""
        for attr in ("embedding_model", "faiss_index", "bm25_index", "tfidf_vectorizer", "tfidf_matrix", "chunks"):
            self.__dict__.pop(attr, None)
        self.is_ready = False
        gc.collect()
        if _cuda_available():
            import torch
            torch.cuda.empty_cache()


class MultiModelRetrieval:
    """Several dense retrievers side by side (src/retrieval.py:339-389)."""

    def __init__(self, model_paths: List[str], device=None, encoders: Optional[Dict[str, Any]] = None):
        self.model_paths = model_paths
        self.device = device or ("cuda" if _cuda_available() else "cpu")
        self.retrievers: Dict[str, RetrievalSystem] = {}
        self._encoders = encoders or {}

    def setup_retrievers(self, chunk_file: str, faiss_indices: Dict[str, str]):
        print("Setting up retrievers for all models...")
        for model_path in self.model_paths:
            model_name = os.path.basename(model_path)
            print(f"\n--- Setting up retriever for {model_name} ---")
            try:
                retriever = RetrievalSystem(method="dense", model_path=model_path, device=self.device,
                                            encoder=self._encoders.get(model_name))
                if retriever.load_chunks_and_index(chunk_file, faiss_indices.get(model_name)):
                    self.retrievers[model_name] = retriever
                    print(f"✓ {model_name} retriever ready")
                else:
                    print(f"✗ Failed to setup {model_name} retriever")
            except Exception as e:
                print(f"Error setting up {model_name}: {e}")

    def compare_retrieval_performance(self, test_queries: List[Dict],
                                      relevant_chunks: Dict[str, List[str]]) -> Dict[str, Dict]:
        results = {}
        for model_name, retriever in self.retrievers.items():
            print(f"\n=== Evaluating {model_name} ===")
            results[model_name] = retriever.evaluate_retrieval_quality(test_queries, relevant_chunks)
        return results

    def cleanup_all(self):
        for retriever in self.retrievers.values():
            retriever.cleanup()
        self.retrievers.clear()
        gc.collect()
