"""B200-native retrieval engine behind the retriever surface of
alirezafarzipour/persian-rag-system (`src/retrieval.py`).

Hot path only: exact flat dense search (faiss IndexFlatL2 / IndexFlatIP), BM25 / TF-IDF scoring
with top-k, the encoder pooling epilogue, and the row-sharded top-k merge -- hand-written
sm_100a CUDA in `csrc/`, reached through the C ABI in `include/prs.h` (ctypes).  There is no
CPU compute path in this package.
"""
from ._lib import (BF16, F16, F32, F64, MAX_K, METRIC_INNER_PRODUCT, METRIC_L2, PrsError, build, lib)
METRIC_IP = METRIC_INNER_PRODUCT      # short alias (faiss spells it METRIC_INNER_PRODUCT)
from .flat import FlatIndex, IndexFlatIP, IndexFlatL2, read_index, write_index
from .index_build import create_model_embeddings, index_path_for, setup_faiss_index
from .sparse import BM25Index, SparseIndex, TfidfIndex
from .pooling import mean_pool_normalize
from .hybrid import hybrid_fuse
from .ivf import IndexIVFFlat
from .retrieval import MultiModelRetrieval, RetrievalSystem
from .sharded import ShardedFlatIndex, ShardedSparseIndex
from .container import read_sharded, write_sharded

__all__ = [
    "FlatIndex", "IndexFlatL2", "IndexFlatIP", "IndexIVFFlat", "read_index", "write_index",
    "SparseIndex", "BM25Index", "TfidfIndex", "mean_pool_normalize", "hybrid_fuse",
    "RetrievalSystem", "MultiModelRetrieval", "ShardedFlatIndex", "ShardedSparseIndex", "read_sharded", "write_sharded", "create_model_embeddings", "setup_faiss_index", "index_path_for",
    "METRIC_L2", "METRIC_INNER_PRODUCT", "METRIC_IP", "F32", "F16", "BF16", "F64", "MAX_K", "PrsError", "build", "lib",
]
