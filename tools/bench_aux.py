#!/usr/bin/env python
"""Measurement of the two smaller §8 rows beside their CPU paths (GPU box, 1 GPU):

  sparse : BM25 scoring + top-k over a synthetic Persian-vocabulary CSR matrix of the C4 shape
           (V = 200k terms, ~100 distinct terms per doc, Zipf(1.07) term ids, tf ~ 1+Geom(0.7),
           queries of 1+Poisson(6) Zipf tokens, k = 10).  HBM roofline: 8 bytes per posting touched.
  pool   : masked mean-pool + L2-normalise epilogue (B x T x H fp16).  HBM roofline: bytes of `hidden`.

Each prints ONE JSON line (metric, value, roofline, cpu_baseline, parity).  Not the driver's bench
contract (that is bench.py); results are copied into profiles/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return (float(json.load(open(p))["hbm_gbs"]), "measured") if os.path.exists(p) else (6650.0, "fallback")


def gen_sparse(n_docs, n_terms, seed, dev):
    """C4-shaped doc-by-term BM25 weight matrix, generated on the device, returned as host CSR."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    # doc length ~ max(1, round(LogNormal(ln 100 - 0.125, 0.5))) distinct terms
    ln = torch.empty(n_docs, device=dev).normal_(mean=float(np.log(100.0) - 0.125), std=0.5, generator=g).exp().round().clamp_(min=1).long()
    # Zipf(1.07) over the vocabulary by inverse CDF
    ranks = torch.arange(1, n_terms + 1, device=dev, dtype=torch.float64)
    cdf = torch.cumsum(ranks.pow(-1.07), 0)
    cdf /= cdf[-1].clone()
    tot = int(ln.sum().item())
    doc = torch.repeat_interleave(torch.arange(n_docs, device=dev), ln)
    keys = torch.empty(tot, dtype=torch.int64, device=dev)
    step = 1 << 26
    for a in range(0, tot, step):
        b = min(tot, a + step)
        u = torch.rand(b - a, generator=g, device=dev, dtype=torch.float64)
        term = torch.searchsorted(cdf, u).clamp_(max=n_terms - 1)
        keys[a:b] = doc[a:b] * n_terms + term
    del doc
    keys = torch.unique(keys)                                   # sorted; duplicates inside a doc collapse ("without replacement")
    doc = torch.div(keys, n_terms, rounding_mode="floor")
    term = (keys - doc * n_terms).to(torch.int32)
    del keys
    nnz = int(doc.numel())
    tf = (1 + torch.empty(nnz, device=dev).geometric_(0.7, generator=g) - 1).to(torch.float64)      # 1 + Geometric(0.7) in {1,2,..}
    counts = torch.bincount(doc, minlength=n_docs)
    indptr = torch.zeros(n_docs + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(counts, 0)
    dl = torch.zeros(n_docs, dtype=torch.float64, device=dev).index_add_(0, doc, tf)
    avgdl = dl.mean()
    df = torch.bincount(term.long(), minlength=n_terms).to(torch.float64)
    # rank_bm25 0.2.2: idf = ln(N - df + 0.5) - ln(df + 0.5); negatives -> eps * mean(idf)   (src/retrieval.py:67 uses the defaults)
    idf = torch.log(n_docs - df + 0.5) - torch.log(df + 0.5)
    present = df > 0
    avg_idf = idf[present].sum() / present.sum()
    idf = torch.where(idf < 0, 0.25 * avg_idf, idf)
    k1, bb = 1.5, 0.75
    w = idf[term.long()] * tf * (k1 + 1) / (tf + k1 * (1 - bb + bb * dl[doc] / avgdl))
    return (indptr.cpu().numpy(), term.cpu().numpy(), w.to(torch.float32).cpu().numpy(), cdf, df.cpu().numpy())


def run_sparse(a):
    import torch
    import scipy.sparse as sp
    import persian_rag_system_b200 as P
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    t0 = time.time()
    indptr, indices, values, cdf, df = gen_sparse(a.docs, a.terms, 7, dev)
    nnz = int(indices.shape[0])
    t_gen = time.time() - t0
    print(f"[sparse] generated nnz={nnz} in {t_gen:.1f}s", file=sys.stderr, flush=True)
    t0 = time.time()
    idx = P.SparseIndex(indptr, indices, values, a.terms)
    t_build = time.time() - t0
    print(f"[sparse] index built in {t_build:.1f}s", file=sys.stderr, flush=True)
    # queries: 1 + Poisson(6) Zipf tokens (stop-word-like heads included, the reference removes none)
    rng = np.random.default_rng(11)
    qlen = 1 + rng.poisson(6, size=a.queries)
    q_indptr = np.zeros(a.queries + 1, np.int64)
    q_indptr[1:] = np.cumsum(qlen)
    u = torch.from_numpy(rng.random(int(q_indptr[-1]))).to(dev)
    q_terms = torch.searchsorted(cdf, u).clamp_(max=a.terms - 1).to(torch.int32).cpu().numpy()
    q_w = np.ones(q_terms.shape[0], np.float64)
    torch.cuda.synchronize()
    idx.search(q_indptr[:9], q_terms[: q_indptr[8]], q_w[: q_indptr[8]], a.k)             # warm-up
    print("[sparse] warm-up search done", file=sys.stderr, flush=True)
    d_ip, d_qt, d_qw = (torch.from_numpy(v).to(dev) for v in (q_indptr, q_terms, q_w))
    modes = {}
    for mode in a.modes.split(","):
        t0 = time.time()
        idx.set_mode(mode)
        torch.cuda.synchronize()
        t_prep = time.time() - t0
        idx.search_device(d_ip, d_qt, d_qw, a.k)
        torch.cuda.synchronize()
        ev = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            Sd, Id = idx.search_device(d_ip, d_qt, d_qw, a.k)
            e1.record()
            torch.cuda.synchronize()
            ev.append(e0.elapsed_time(e1) * 1e-3)
        reps = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            S, I = idx.search(q_indptr, q_terms, q_w, a.k)
            reps.append(time.perf_counter() - t0)
        assert np.array_equal(S, Sd.cpu().numpy()) and np.array_equal(I, Id.cpu().numpy())
        modes[mode] = {"device_ms": float(np.median(ev)) * 1e3, "host_call_ms": float(np.median(reps)) * 1e3, "prepare_s": t_prep, "S": S, "I": I}
        print(f"[sparse] mode {mode}: device {modes[mode]['device_ms']:.2f} ms, host call {modes[mode]['host_call_ms']:.2f} ms", file=sys.stderr, flush=True)
    head = a.modes.split(",")[-1]
    S, I = modes[head]["S"], modes[head]["I"]
    t_gpu = modes[head]["device_ms"] * 1e-3
    postings = idx.last_postings
    peak, src = hbm_peak()
    gbs = 8.0 * postings / t_gpu / 1e9
    # CPU path beside it: vectorised BM25 over a scipy CSC matrix + argpartition/stable sort (oracle "port")
    nc = min(a.cpu_queries, a.queries)
    M = sp.csr_matrix((values.astype(np.float64), indices, indptr), shape=(a.docs, a.terms)).tocsc()
    t0 = time.perf_counter()
    bad = 0
    for qi in range(nc):
        terms = q_terms[q_indptr[qi]:q_indptr[qi + 1]]
        sc = np.zeros(a.docs, np.float64)
        for t in terms:                                             # repeated tokens add repeatedly, in order
            lo, hi = M.indptr[t], M.indptr[t + 1]
            sc[M.indices[lo:hi]] += M.data[lo:hi]
        top = O.argsort_topk_canonical(sc, a.k)
        if qi < a.check:
            try:
                O.check_topk_against_scores(I[qi], S[qi], sc, a.k, True, rtol=1e-5, atol=1e-9, what=f"bm25 q{qi}")
                if not np.array_equal(I[qi], top):
                    pass                                            # differences are ties within tolerance (checked above)
            except AssertionError as e:
                bad += 1
                print("PARITY", str(e)[:300], file=sys.stderr)
    t_cpu = (time.perf_counter() - t0) / nc
    out = {"metric": f"BM25 QPS @k={a.k}, {a.docs} docs x {a.terms} terms CSR (~{nnz / a.docs:.0f} nnz/doc)", "value": a.queries / t_gpu,
           "unit": "queries/s", "n_gpus": 1, "ms_per_batch": t_gpu * 1e3, "dtype": "f32 weights; selection " + head + ", exact f64 re-score", "data": "synthetic",
           "modes": {m: {k: v for k, v in d.items() if k not in ("S", "I")} | {"qps_device": a.queries / (d["device_ms"] * 1e-3), "qps_host_call": a.queries / (d["host_call_ms"] * 1e-3),
                                                                            "algorithmic_gbs": 8.0 * postings / (d["device_ms"] * 1e-3) / 1e9} for m, d in modes.items()},
           "config": {"workload": "configs[3] shape (scaled docs)", "docs": a.docs, "terms": a.terms, "nnz": nnz, "queries": a.queries, "k": a.k,
                      "postings_touched": int(postings), "avg_postings_per_query": postings / a.queries},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                        "bytes_per_batch": 8.0 * postings, "peak_source": src,
                        "note": "prs_sparse_search_device (query CSR and results on the device), CUDA events around the call: score kernel + merge + exact re-score"},
           "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "queries/s", "cores": 1, "kind": "port",
                            "sample": f"first {nc} queries, scipy CSC column adds + top-k (rank_bm25 itself is absent; this is faster than its pure-Python loop)"},
           "parity": {"queries_checked": min(a.check, nc), "mismatch": bad},
           "build": {"generate_s": t_gen, "index_build_s": t_build}}
    print(json.dumps(out))


def run_pool(a):
    import torch
    import persian_rag_system_b200 as P
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    res = []
    peak, src = hbm_peak()
    for (B, T, H, dt) in [(32, 128, 384, torch.float16), (32, 512, 768, torch.float16), (256, 512, 768, torch.float16), (256, 512, 768, torch.float32)]:
        g = torch.Generator(device=dev).manual_seed(B + T)
        hid = torch.randn(B, T, H, generator=g, device=dev).to(dt)
        lens = torch.randint(1, T + 1, (B,), generator=g, device=dev)
        mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).to(torch.int64)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(3):
            out = P.mean_pool_normalize(hid, mask, True)
        ts = []
        for i in range(20):
            flush.fill_(i)                                   # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = P.mean_pool_normalize(hid, mask, True); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        nbytes = B * T * H * hid.element_size() + B * T * 8 + B * H * 4
        # torch eager beside it (same device): what sentence-transformers runs
        def eager():
            m = mask.unsqueeze(-1).to(hid.dtype)
            s = (hid * m).sum(1) / m.sum(1).clamp(min=1e-9)
            return torch.nn.functional.normalize(s.float(), p=2, dim=1)
        for _ in range(3):
            eager()
        te = []
        for i in range(10):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eager(); e1.record(); torch.cuda.synchronize(); te.append(e0.elapsed_time(e1))
        hn, mn = hid.float().cpu().numpy(), mask.cpu().numpy()
        t0 = time.perf_counter(); ref = O.mean_pool_normalize(hn, mn, True); t_cpu = time.perf_counter() - t0
        err = float(np.abs(out.cpu().numpy() - ref).max())
        res.append({"B": B, "T": T, "H": H, "dtype": str(dt).replace("torch.", ""), "ms": ms, "GBs": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / peak,
                    "torch_eager_ms": float(np.median(te)), "cpu_numpy_ms": t_cpu * 1e3, "max_abs_err_vs_oracle": err})
    print(json.dumps({"metric": "masked mean-pool + L2-normalise epilogue", "unit": "ms", "peak_GBs": peak, "peak_source": src, "cases": res,
                      "cpu_baseline": {"kind": "port", "cores": 1, "sample": "oracle numpy restatement of sentence-transformers Pooling+Normalize on the same tensors"}}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["sparse", "pool"])
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--terms", type=int, default=200_000)
    ap.add_argument("--queries", type=int, default=256)
    ap.add_argument("--cpu-queries", type=int, default=64)
    ap.add_argument("--check", type=int, default=32)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--modes", default="exact,throughput", help="comma list; the LAST one is the headline")
    a = ap.parse_args()
    run_sparse(a) if a.what == "sparse" else run_pool(a)
