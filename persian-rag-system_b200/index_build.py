"""Index build on the B200 engine -- the §8 a-2 row.

Mirrors the two index builders of the reference (same names, arguments, return values, file
naming and skip-if-exists behaviour), with the faiss calls replaced by the flat index of this
package:

  create_model_embeddings(model_path, chunk_file, chunk_type) -> bool   src/create_embeddings.py:54-153
      results/faiss/{basename(model_path)}_drugs_{chunk_type}_chunks.index   (:62), skipped when it exists (:64-66)
  setup_faiss_index(embeddings, index_type="flat") -> index                scripts/phase3_pdf_chunking.py:39-71

Differences, all additive: an `encoder` can be passed instead of loading SentenceTransformer (the
encoders are unchanged by this project); `storage` selects fp32 (byte-exact faiss file, default)
or fp16 / bf16 (T64 layout, tcgen05 scan); encoder batches that arrive as CUDA tensors are added
without a host round trip.  The reference's IVF branch (n >= 1000: IndexIVFFlat, nlist = min(100, max(10, n//20)),
trained on the first 10 000 rows, nprobe = 1) is reproduced by ivf.IndexIVFFlat on the same flat-scan kernels;
`exact=True` keeps the exact index instead.
"""
from __future__ import annotations

import os
import time
from typing import Any, Iterable, Optional

import numpy as np

from .flat import FlatIndex, IndexFlatL2, write_index

FAISS_DIR = "results/faiss"                 # src/create_embeddings.py:60
ENCODE_BATCH = 32                           # src/create_embeddings.py:88


def index_path_for(model_path: str, chunk_type: str, faiss_dir: str = FAISS_DIR) -> str:
    """results/faiss/{model}_drugs_{word|sentence}_chunks.index (src/create_embeddings.py:62)."""
    return f"{faiss_dir}/{os.path.basename(model_path)}_drugs_{chunk_type}_chunks.index"


def _batches(items, size: int) -> Iterable[list]:
    for start in range(0, len(items), size):
        yield items[start:start + size]


def _add_rows(index: FlatIndex, rows) -> None:
    """One encoder batch -> index.  CUDA tensors go straight in; anything else becomes float32 numpy
    exactly like `np.array(embeddings).astype('float32')` (src/create_embeddings.py:122)."""
    if hasattr(rows, "is_cuda") and rows.is_cuda:
        index.add(rows)
    else:
        if hasattr(rows, "detach"):
            rows = rows.detach().cpu().numpy()
        index.add(np.asarray(rows).astype("float32"))


def setup_faiss_index(embeddings, index_type: str = "flat", storage: str = "fp32", device: Optional[int] = None, exact: bool = False):
    """The reference's builder (scripts/phase3_pdf_chunking.py:39-71): a flat squared-L2 index for
    `index_type == "flat"` or fewer than 1000 embeddings, otherwise IVF-Flat with
    `nlist = min(100, max(10, n // 20))` trained on the first 10 000 rows (faiss's default nprobe = 1 at search
    time); rows are added in 1000-row batches (:59-64).  `exact=True` keeps the exact flat index for every size
    (a superset in recall of the approximate branch the reference silently switches to)."""
    n, d = int(embeddings.shape[0]), int(embeddings.shape[1])
    if index_type == "flat" or n < 1000 or exact:
        print(f"Setting up flat L2 index for {n} embeddings ({storage} rows in HBM)...")
        index = IndexFlatL2(d, storage=storage, device=device)
        index.reserve(n)
    else:
        from .ivf import IndexIVFFlat
        nlist = min(100, max(10, n // 20))
        print(f"Setting up IVF-Flat index for {n} embeddings (nlist={nlist}, nprobe=1, {storage} lists in HBM)...")
        index = IndexIVFFlat(d, nlist, storage=storage, device=device)
        print("  Training FAISS index...")
        training = embeddings[: min(10000, n)]
        if hasattr(training, "detach"):
            training = training.detach().float().cpu().numpy()
        index.train(np.asarray(training).astype("float32"))
    for start in range(0, n, 1000):
        _add_rows(index, embeddings[start:start + 1000])
    print("✓ index resident on the device")
    return index


def create_model_embeddings(model_path: str, chunk_file: str, chunk_type: str, encoder: Any = None, storage: str = "fp32",
                            faiss_dir: str = FAISS_DIR, device: Optional[int] = None) -> bool:
    """Encode the chunks of `chunk_file` with the model and write the flat index file.  True when the
    file exists afterwards (including "already there"), False on any failure -- the reference's
    contract (src/create_embeddings.py:64-66, 68-70, 151-153)."""
    model_name = os.path.basename(model_path)
    index_file = index_path_for(model_path, chunk_type, faiss_dir)
    os.makedirs(faiss_dir, exist_ok=True)
    if os.path.exists(index_file):
        print(f"✓ Index already exists: {index_file}")
        return True
    if not os.path.exists(chunk_file):
        print(f"✗ Chunk file not found: {chunk_file}")
        return False
    try:
        started = time.time()
        if encoder is None:
            from sentence_transformers import SentenceTransformer
            from .retrieval import _cuda_available
            encoder = SentenceTransformer(model_path, device="cuda" if _cuda_available() else "cpu")
        import pandas as pd
        texts = pd.read_csv(chunk_file, encoding="utf-8")["text"].tolist()
        index = None
        for batch in _batches(texts, ENCODE_BATCH):
            rows = encoder.encode(batch, show_progress_bar=False, convert_to_numpy=True)
            if index is None:
                index = IndexFlatL2(int(rows.shape[1]), storage=storage, device=device)
                index.reserve(len(texts))
            _add_rows(index, rows)
        if index is None:
            raise ValueError("no chunks to embed")
        write_index(index, index_file)
        print(f"✓ wrote {index_file}: {index.ntotal} vectors x {index.d} in {time.time() - started:.2f}s")
        return True
    except Exception as exc:                                     # noqa: BLE001 - same net as the reference
        print(f"✗ Error creating embeddings for {model_name} ({chunk_type}): {exc}")
        return False
