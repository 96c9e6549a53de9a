"""Sharded on-disk container for large flat corpora (second half).

The reference persists ONE faiss file per model (`faiss.write_index`, src/create_embeddings.py:136; read back at
src/retrieval.py:55).  That format (kept byte-exact by `flat.write_index` / `read_index`) is fp32, row-major and
loaded through one host thread -- fine for 125 rows, not for the north-star's 400 M x 384 fp16 corpus (307 GB).
This container is what that corpus is stored in:

    <dir>/manifest.json          format, d, metric, storage, ntotal, [{file, rows, id_offset}, ...]
    <dir>/shard_00000.prst ...   one file per shard = the HBM image itself (T64 blocks for 16-bit storage) followed by
                                 the float32 squared norms, page aligned (mmap-able), see csrc/flat_index.cu

Every rank writes / loads its own shard file(s) in parallel; a load is a straight double-buffered copy through
page-locked staging memory (no conversion kernel, no norm recomputation)."""
from __future__ import annotations

import ctypes
import json
import os
from typing import List

from . import _lib
from ._lib import check
from .flat import FlatIndex, _default_device

MANIFEST = "manifest.json"
FORMAT = "prs-sharded-flat/1"


def _shard_name(i: int) -> str:
    return f"shard_{i:05d}.prst"


def write_shard(index: FlatIndex, path: str) -> None:
    check(index._L.prs_index_write_shard(index._h, os.fsencode(path)))


def read_shard(path: str, device: int | None = None, into: FlatIndex | None = None) -> FlatIndex:
    """Load one shard file; with `into`, append it to that index (consecutive shards of one rank)."""
    L = _lib.lib()
    dev = (_default_device() if device is None else int(device)) if into is None else into.device
    h = ctypes.c_void_p() if into is None else into._h
    check(L.prs_index_read_shard(os.fsencode(path), dev, ctypes.byref(h)))
    return FlatIndex(0, _handle=h, device=dev) if into is None else into


def write_sharded(index, dirpath: str) -> None:
    """`index`: a FlatIndex (one shard) or a ShardedFlatIndex (every rank writes its block; rank 0 writes the manifest)."""
    os.makedirs(dirpath, exist_ok=True)
    local = getattr(index, "local", index)
    rank, world = getattr(index, "rank", 0), getattr(index, "world", 1)
    offset = int(getattr(index, "offset", 0))
    write_shard(local, os.path.join(dirpath, _shard_name(rank)))
    rows: List[int] = [local.ntotal]
    offsets: List[int] = [offset]
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, (local.ntotal, offset), group=index.group)
        rows, offsets = [g[0] for g in gathered], [g[1] for g in gathered]
    if rank == 0:
        man = {"format": FORMAT, "d": local.d, "metric": local.metric_type, "storage": local.storage, "ntotal": int(sum(rows)),
               "layout": "T64 (64-row blocks, k-block-major, 128B-swizzled)" if local.storage != "fp32" else "row-major, pitch = d rounded up to 64",
               "shards": [{"file": _shard_name(i), "rows": int(rows[i]), "id_offset": int(offsets[i])} for i in range(world)]}
        tmp = os.path.join(dirpath, MANIFEST + ".tmp")
        with open(tmp, "w") as f:
            json.dump(man, f, indent=1)
        os.replace(tmp, os.path.join(dirpath, MANIFEST))
    if world > 1:
        import torch.distributed as dist
        dist.barrier(group=index.group)


def read_manifest(dirpath: str) -> dict:
    with open(os.path.join(dirpath, MANIFEST)) as f:
        man = json.load(f)
    if man.get("format") != FORMAT:
        raise _lib.PrsError(_lib.EIO, f"{dirpath}: not a {FORMAT} container")
    return man


def shards_for_rank(n_shards: int, world: int, rank: int):
    """Contiguous block of shard indices for `rank` (same rule as sharded.shard_bounds over shard files)."""
    per = (n_shards + world - 1) // world
    lo = min(n_shards, rank * per)
    return range(lo, min(n_shards, lo + per))


def read_sharded(dirpath: str, device: int | None = None, rank: int = 0, world: int = 1) -> FlatIndex:
    """Load the shard file(s) of `rank` out of `world` into one FlatIndex on `device` (world = 1: the whole
    corpus).  Global ids are preserved through the index's id offset."""
    man = read_manifest(dirpath)
    mine = list(shards_for_rank(len(man["shards"]), world, rank))
    if not mine:
        raise _lib.PrsError(_lib.EINVAL, f"{dirpath}: {len(man['shards'])} shard files cannot feed rank {rank} of {world}")
    idx = None
    for i in mine:
        sh = man["shards"][i]
        if idx is not None and idx.ntotal + int(man["shards"][mine[0]]["id_offset"]) != int(sh["id_offset"]):
            raise _lib.PrsError(_lib.EINVAL, f"{dirpath}: shards {mine} are not contiguous in global id")
        idx = read_shard(os.path.join(dirpath, sh["file"]), device, into=idx)
    idx.set_id_offset(int(man["shards"][mine[0]]["id_offset"]))
    if idx.d != int(man["d"]) or idx.storage != man["storage"] or idx.metric_type != int(man["metric"]):
        raise _lib.PrsError(_lib.EIO, f"{dirpath}: shard files do not match the manifest (d / storage / metric)")
    return idx
