"""oracle/make_golden.py -- TEST INFRASTRUCTURE.

Generates the committed fixtures under tests/golden/ from the reference tree.  Run ONCE in the
build container (where /root/reference exists); the GPU box only sees the committed outputs.

    python oracle/make_golden.py

Outputs
  tests/golden/indices/*.index       3 of the reference's 14 shipped faiss IndexFlatL2 files
                                     (results/faiss/, one per embedding width 384/512/768), verbatim
  tests/golden/index_manifest.json   size / sha256 / d / ntotal / fourcc of ALL 14 shipped indices
  tests/golden/phase4_records.json   the (id, distance, similarity_score) records the reference
                                     recorded in results/phase4_rag_evaluation_results.json
                                     + the 29 distinct chunk texts and 10 questions they contain
  tests/golden/flat_golden.npz       seeded queries + C-oracle top-k for the 3 committed indices
  tests/golden/tfidf_golden.npz/json sklearn (the real reference implementation) TF-IDF scores
  tests/golden/bm25_golden.npz       rank_bm25 restatement scores on the same texts (unpinned)
  tests/golden/pool_golden.npz       torch mean-pool / normalize outputs
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
COMMITTED = [
    "paraphrase-multilingual-MiniLM-L12-v2_finetuned_drugs_word_chunks.index",   # d=384, n=125
    "drugs_sentence_chunks.index",                                                # d=512, n=121
    "multilingual-e5-base_drugs_word_chunks.index",                               # d=768, n=125 (unit norm)
]


def make_queries(x: np.ndarray, nq: int, seed: int) -> np.ndarray:
    """seeded Gaussian perturbations (sigma = 0.1*||row||/sqrt(d)) of cyclic rows."""
    rng = np.random.default_rng(seed)
    n, d = x.shape
    rows = x[np.arange(nq) % n]
    sigma = 0.1 * np.linalg.norm(rows, axis=1, keepdims=True) / np.sqrt(d)
    return (rows + sigma * rng.standard_normal((nq, d))).astype(np.float32)


def main():
    os.makedirs(os.path.join(GOLD, "indices"), exist_ok=True)
    faiss_dir = os.path.join(REF, "results", "faiss")
    manifest = {}
    for f in sorted(os.listdir(faiss_dir)):
        b = open(os.path.join(faiss_dir, f), "rb").read()
        x, metric = O.read_faiss_flat(os.path.join(faiss_dir, f))
        manifest[f] = {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest(), "fourcc": b[:4].decode(),
                       "d": int(x.shape[1]), "ntotal": int(x.shape[0]), "metric": int(metric),
                       "row_norm_mean": float(np.linalg.norm(x, axis=1).mean())}
    json.dump(manifest, open(os.path.join(GOLD, "index_manifest.json"), "w"), indent=1)
    for f in COMMITTED:
        shutil.copyfile(os.path.join(faiss_dir, f), os.path.join(GOLD, "indices", f))
        os.chmod(os.path.join(GOLD, "indices", f), 0o644)

    # ---- flat search goldens (C oracle, fp32, both metrics) ----
    out = {}
    for t, f in enumerate(COMMITTED):
        x, _ = O.read_faiss_flat(os.path.join(GOLD, "indices", f))
        q = make_queries(x, 64, seed=100 + t)
        out[f"q_{t}"] = q
        for k in (1, 5, 10, 20):
            D, I = O.flat_search_c(x, q, k, O.METRIC_L2, form=1)
            out[f"l2_D_{t}_{k}"], out[f"l2_I_{t}_{k}"] = D, I
            D, I = O.flat_search_c(x, q, k, O.METRIC_IP)
            out[f"ip_D_{t}_{k}"], out[f"ip_I_{t}_{k}"] = D, I
    out["files"] = np.array(COMMITTED)
    np.savez_compressed(os.path.join(GOLD, "flat_golden.npz"), **out)

    # ---- recorded retrievals of the reference's own runs ----
    j = json.load(open(os.path.join(REF, "results", "phase4_rag_evaluation_results.json")))
    records, texts, questions = [], {}, []
    for cfg, blob in j.items():
        for item in blob["retrieval_metrics"].get("detailed_retrievals", []):
            if item["question"] not in questions:
                questions.append(item["question"])
            recs = []
            for r in item["retrieved"]:
                texts.setdefault(r["id"], r["text"])
                recs.append({"id": r["id"], "distance": float(r["distance"]),
                             "similarity_score": float(r["similarity_score"])})
            records.append({"config": cfg, "question": item["question"], "retrieved": recs})
    ids = sorted(texts, key=lambda s: (s.split("_")[0], int(s.rsplit("_", 1)[1])))
    json.dump({"records": records, "questions": questions,
               "chunks": [{"id": i, "text": texts[i]} for i in ids]},
              open(os.path.join(GOLD, "phase4_records.json"), "w"), ensure_ascii=False, indent=1)

    # ---- TF-IDF goldens from scikit-learn (the reference's actual implementation) ----
    chunk_texts = [texts[i] for i in ids]
    # queries: the 10 recorded questions (share ~no tokens with the reversed-glyph chunk text,
    # finding 6) plus 10 queries cut from the chunk texts themselves so that scores are non-trivial
    rng = np.random.default_rng(7)
    cut = []
    for t in range(10):
        words = chunk_texts[(3 * t) % len(chunk_texts)].split()
        s = int(rng.integers(0, max(1, len(words) - 8)))
        cut.append(" ".join(words[s:s + 6] + words[s + 2:s + 4]))     # repeated tokens on purpose
    queries = questions + cut
    vec, mat = O.tfidf_fit(chunk_texts)
    S = np.stack([O.tfidf_scores(vec, mat, q) for q in queries])
    feats = vec.get_feature_names_out().tolist()
    np.savez_compressed(os.path.join(GOLD, "tfidf_golden.npz"), scores=S,
                        idf=vec.idf_, data=mat.data, indices=mat.indices, indptr=mat.indptr,
                        shape=np.array(mat.shape))
    json.dump({"queries": queries, "features": feats}, open(os.path.join(GOLD, "tfidf_golden.json"), "w"),
              ensure_ascii=False)

    # ---- BM25 (restatement; unpinned) ----
    bm = O.BM25OkapiOracle([t.split() for t in chunk_texts])
    Sb = np.stack([bm.get_scores(q.split()) for q in queries])
    np.savez_compressed(os.path.join(GOLD, "bm25_golden.npz"), scores=Sb)

    # ---- pooling goldens from torch ----
    import torch
    g = torch.Generator().manual_seed(5)
    B, T, H = 5, 37, 384
    h = torch.randn(B, T, H, generator=g)
    lens = torch.tensor([37, 1, 20, 0, 13])
    mask = (torch.arange(T)[None, :] < lens[:, None]).to(torch.int64)
    me = mask.unsqueeze(-1).expand(h.size()).float()
    pooled = torch.sum(h * me, 1) / torch.clamp(me.sum(1), min=1e-9)
    normed = torch.nn.functional.normalize(pooled, p=2, dim=1)
    np.savez_compressed(os.path.join(GOLD, "pool_golden.npz"), hidden=h.numpy(), mask=mask.numpy(),
                        pooled=pooled.numpy(), normalized=normed.numpy())
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
