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


class RetrievalSystem:
    def __init__(self, method="dense", model_path=None, device=None, encoder=None, storage="fp32"):
        """method: "dense" | "bm25" | "tfidf" | "hybrid" (src/retrieval.py:13-21).

        `encoder` (extension): any object with `.encode(list[str], device=...) -> ndarray`, used
        instead of loading `SentenceTransformer(model_path)`.  `storage` (extension): how the
        dense corpus is held in HBM: "fp32" (reference parity), "fp16", "bf16"."""
        self.method = method
        self.device = device or ("cuda" if _cuda_available() else "cpu")
        self.storage = storage
        if encoder is not None:
            self.embedding_model = encoder
        elif method in ["dense", "hybrid"] and model_path:
            print(f"Loading embedding model: {model_path}")
            from sentence_transformers import SentenceTransformer   # unchanged encoder (north-star)
            self.embedding_model = SentenceTransformer(model_path, device=self.device)
        else:
            self.embedding_model = None
        self.chunks = None
        self.faiss_index = None
        self.bm25_index = None
        self.tfidf_vectorizer = None
        self.tfidf_matrix = None
        self.is_ready = False

    # ------------------------------------------------------------------ load (src/retrieval.py:38-90)
    def load_chunks_and_index(self, chunk_file: str, faiss_index_file: str = None):
        print(f"Loading chunks from {chunk_file}...")
        try:
            self.chunks = _load_chunks_csv(chunk_file)
            print(f"✓ Loaded {len(self.chunks)} chunks")
        except Exception as e:
            print(f"Error loading chunks: {e}")
            return False
        return self._build_indices(faiss_index_file)

    def load_chunks(self, chunks: List[Dict[str, Any]], faiss_index_file: str = None):
        """Extension: same as load_chunks_and_index with the chunk records already in memory."""
        self.chunks = list(chunks)
        return self._build_indices(faiss_index_file)

    def _build_indices(self, faiss_index_file):
        if self.method in ["dense", "hybrid"] and faiss_index_file and os.path.exists(faiss_index_file):
            try:
                print(f"Loading FAISS index from {faiss_index_file}...")
                self.faiss_index = read_index(faiss_index_file, storage=self.storage)
                print(f"✓ Loaded FAISS index with {self.faiss_index.ntotal} vectors")
            except Exception as e:
                print(f"Error loading FAISS index: {e}")
                return False
        if self.method in ["bm25", "hybrid"]:
            print("Building BM25 index...")
            try:
                tokenized_chunks = [chunk["text"].split() for chunk in self.chunks]
                self.bm25_index = BM25Index(tokenized_chunks)
                print("✓ BM25 index built successfully")
            except Exception as e:
                print(f"Error building BM25 index: {e}")
                return False
        if self.method in ["tfidf", "hybrid"]:
            print("Building TF-IDF index...")
            try:
                chunk_texts = [chunk["text"] for chunk in self.chunks]
                self.tfidf_vectorizer = TfidfIndex(chunk_texts, max_features=10000, ngram_range=(1, 2))
                self.tfidf_matrix = self.tfidf_vectorizer.index
                print("✓ TF-IDF index built successfully")
            except Exception as e:
                print(f"Error building TF-IDF index: {e}")
                return False
        self.is_ready = True
        return True

    # ------------------------------------------------------------------ dense (src/retrieval.py:92-115)
    def retrieve_dense(self, query: str, top_k: int = 10) -> List[Tuple[Dict, float]]:
        if not self.embedding_model or not self.faiss_index:
            return []
        try:
            query_embedding = self.embedding_model.encode([query], device=self.device)
            query_embedding = np.asarray(query_embedding).astype("float32")
            distances, indices = self.faiss_index.search(query_embedding, top_k)
            results = []
            for distance, idx in zip(distances[0], indices[0]):
                if idx >= 0 and idx < len(self.chunks):
                    similarity = 1 / (1 + distance)       # squared-L2 -> score, src/retrieval.py:108
                    results.append((self.chunks[idx], similarity))
            return results
        except Exception as e:
            print(f"Error in dense retrieval: {e}")
            return []

    # ------------------------------------------------------------------ bm25 (src/retrieval.py:117-143)
    def retrieve_bm25(self, query: str, top_k: int = 10) -> List[Tuple[Dict, float]]:
        if not self.bm25_index:
            return []
        try:
            scores, top_indices = self.bm25_index.get_top_k(query.split(), top_k)
            results = []
            for idx, score in zip(top_indices, scores):
                if idx < len(self.chunks):
                    results.append((self.chunks[idx], score))
            return results
        except Exception as e:
            print(f"Error in BM25 retrieval: {e}")
            return []

    # ------------------------------------------------------------------ tfidf (src/retrieval.py:145-172)
    def retrieve_tfidf(self, query: str, top_k: int = 10) -> List[Tuple[Dict, float]]:
        if not self.tfidf_vectorizer or self.tfidf_matrix is None:
            return []
        try:
            scores, top_indices = self.tfidf_vectorizer.get_top_k(query, top_k)
            results = []
            for idx, score in zip(top_indices, scores):
                if idx < len(self.chunks):
                    results.append((self.chunks[idx], score))
            return results
        except Exception as e:
            print(f"Error in TF-IDF retrieval: {e}")
            return []

    # ------------------------------------------------------------------ hybrid (src/retrieval.py:174-220)
    def retrieve_hybrid(self, query: str, top_k: int = 10, dense_weight: float = 0.6,
                        bm25_weight: float = 0.4) -> List[Tuple[Dict, float]]:
        try:
            dense_results = self.retrieve_dense(query, top_k * 2)
            bm25_results = self.retrieve_bm25(query, top_k * 2)
            fused: Dict[Any, Dict[str, Any]] = {}          # insertion order: dense hits first
            for results, mine, weight in ((dense_results, "dense_score", dense_weight),
                                          (bm25_results, "bm25_score", bm25_weight)):
                if not results:
                    continue
                top = max(score for _, score in results)
                for chunk, score in results:
                    part = (score / top if top > 0 else 0) * weight
                    slot = fused.get(chunk["id"])
                    if slot is None:
                        slot = fused[chunk["id"]] = {"chunk": chunk, "dense_score": 0, "bm25_score": 0}
                    slot[mine] = part
            final_results = [(v["chunk"], v["dense_score"] + v["bm25_score"]) for v in fused.values()]
            final_results.sort(key=lambda pair: pair[1], reverse=True)   # stable, like the reference
            return final_results[:top_k]
        except Exception as e:
            print(f"Error in hybrid retrieval: {e}")
            return []

    # ------------------------------------------------------------------ dispatcher (src/retrieval.py:222-238)
    def retrieve(self, query: str, top_k: int = 10) -> List[Tuple[Dict, float]]:
        if not self.is_ready:
            print("Retrieval system is not ready. Please load chunks and index first.")
            return []
        handler = {"dense": self.retrieve_dense, "bm25": self.retrieve_bm25,
                   "tfidf": self.retrieve_tfidf, "hybrid": self.retrieve_hybrid}.get(self.method)
        if handler is None:
            print(f"Unknown retrieval method: {self.method}")
            return []
        return handler(query, top_k)

    # ------------------------------------------------------------------ RAG contexts (src/retrieval.py:240-272)
    def get_contexts_for_rag(self, query: str, top_k: int = 5,
                             max_context_length: int = 2000) -> Tuple[List[str], List[Dict]]:
        contexts: List[str] = []
        metadata: List[Dict] = []
        used = 0
        for chunk, score in self.retrieve(query, top_k):
            text = chunk["text"]
            if used + len(text) > max_context_length:
                room = max_context_length - used
                if room <= 100:
                    break
                text = text[:room] + "..."
            contexts.append(text)
            metadata.append({"chunk_id": chunk["id"], "score": score,
                             "chunk_type": chunk.get("chunk_type", "unknown"), "length": len(text)})
            used += len(text)
            if used >= max_context_length:
                break
        return contexts, metadata

    # ------------------------------------------------------------------ Hit@K / MRR (src/retrieval.py:274-323)
    def evaluate_retrieval_quality(self, test_queries: List[Dict],
                                   relevant_chunks: Dict[str, List[str]]) -> Dict[str, float]:
        print(f"Evaluating retrieval quality on {len(test_queries)} queries...")
        hits = {1: [], 3: [], 5: []}
        mrr_scores = []
        for i, query_data in enumerate(test_queries):
            if i % 50 == 0:
                print(f"  Processing query {i+1}/{len(test_queries)}")
            relevant = relevant_chunks.get(query_data.get("id", str(i)), [])
            if not relevant:
                continue
            retrieved_ids = [chunk["id"] for chunk, _ in self.retrieve(query_data["question"], top_k=10)]
            for cut in hits:
                hits[cut].append(any(cid in relevant for cid in retrieved_ids[:cut]))
            mrr = 0.0
            for rank, cid in enumerate(retrieved_ids, 1):
                if cid in relevant:
                    mrr = 1.0 / rank
                    break
            mrr_scores.append(mrr)
        results = {
            "hit_at_1": np.mean(hits[1]) if hits[1] else 0.0,
            "hit_at_3": np.mean(hits[3]) if hits[3] else 0.0,
            "hit_at_5": np.mean(hits[5]) if hits[5] else 0.0,
            "mrr": np.mean(mrr_scores) if mrr_scores else 0.0,
            "total_queries": len(test_queries),
        }
        print("✓ Retrieval evaluation completed")
        print(f"  Hit@1: {results['hit_at_1']:.3f}")
        print(f"  Hit@3: {results['hit_at_3']:.3f}")
        print(f"  Hit@5: {results['hit_at_5']:.3f}")
        print(f"  MRR: {results['mrr']:.3f}")
        return results

    # ------------------------------------------------------------------ cleanup (src/retrieval.py:325-336)
    def cleanup(self):
        for name in ("embedding_model", "faiss_index", "chunks"):
            if hasattr(self, name):
                delattr(self, name)
        gc.collect()
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        except Exception:
            pass


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
