"""On-disk formats of the reference's corpus hand-off, read and written by the B200 store.

The reference persists an encoded corpus as (`DRT/trainer/trainer.py`):
  * per rank  `{encode_corpus_dir}/{ep}.{rank}.npy`   fp32 [n_rank, d] embeddings        (:210-211)
  * per rank  `{encode_corpus_dir}/{ep}.{rank}.json`  JSON lines `{"id": [batch of ids]}`  (:212-216)
  * rank 0    `{index_file}{ep}`                      faiss.write_index of all shards      (:245)
  * rank 0    `{index_order_dir}/{ep}.docid.txt`      one JSON object `{"id": [...]}`      (:246-248)
and rebuilds the index by adding the `.npy` files in `os.listdir` order (:222-241).

With the device-resident store none of these files is needed between encoding and search, but
corpora that were already encoded by the reference can be ingested as they are
(`load_reference_corpus`), and a store can be written back in the same layout
(`save_reference_shard`, `save_docid_order`), so the two code bases can exchange corpora.
The faiss index file itself is handled by `faiss_compat.write_index` / `read_index`.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_CHUNK_ROWS = 1 << 18


def shard_paths(encode_corpus_dir: str, ep, rank) -> Tuple[str, str]:
    base = os.path.join(encode_corpus_dir, f"{ep}.{rank}")
    return base + ".npy", base + ".json"


def save_reference_shard(index, encode_corpus_dir: str, ep, rank, id_batches: Iterable[Sequence]) -> Tuple[str, str]:
    """Write one rank's rows + ids the way `_encoding_corpus` does (trainer.py:210-216).
    `index` needs `.ntotal`, `.d` and `.reconstruct_n(i0, n)` (IndexFlatIP or one store shard);
    rows are streamed through a memory-mapped .npy so the shard never has to fit in host RAM
    twice."""
    os.makedirs(encode_corpus_dir, exist_ok=True)
    npy, js = shard_paths(encode_corpus_dir, ep, rank)
    n, d = int(index.ntotal), int(index.d)
    out = np.lib.format.open_memmap(npy, mode="w+", dtype=np.float32, shape=(n, d))
    for r0 in range(0, n, _CHUNK_ROWS):
        rows = min(_CHUNK_ROWS, n - r0)
        out[r0:r0 + rows] = index.reconstruct_n(r0, rows)
    out.flush()
    del out
    total = 0
    with open(js, "w", encoding="utf-8") as f:
        for batch in id_batches:
            batch = [x.item() if hasattr(x, "item") else x for x in batch]
            total += len(batch)
            json.dump({"id": batch}, f, ensure_ascii=False)
            f.write("\n")
    if total != n:
        raise ValueError(f"{total} ids for {n} rows in shard {ep}.{rank}")
    return npy, js


def read_id_batches(json_path: str) -> List:
    ids: List = []
    with open(json_path, "r", encoding="utf-8") as f:
        for line in f:
            if line.strip():
                ids.extend(json.loads(line)["id"])          # trainer.py:237-240
    return ids


def list_reference_shards(encode_corpus_dir: str, ep) -> List[str]:
    """The `.json` files of epoch `ep` in the order `_index_corpus` visits them
    (os.listdir order, trainer.py:222-227)."""
    prefix = f"{ep}."
    return [f for f in os.listdir(encode_corpus_dir) if f.startswith(prefix) and f.endswith("json")]


def load_reference_corpus(encode_corpus_dir: str, ep, index=None, index_factory=None, order: Optional[Sequence[str]] = None):
    """Rebuild what `_index_corpus` builds (trainer.py:220-241): every `{ep}.*.npy` added to one
    index in listing order, plus the concatenated doc-id list `idx` (faiss row -> doc id).
    Returns (index, idx).  `order` overrides the file order (e.g. sorted) when given."""
    files = list(order) if order is not None else list_reference_shards(encode_corpus_dir, ep)
    if not files:
        raise FileNotFoundError(f"no '{ep}.*.json' shard files in {encode_corpus_dir}")
    idx: List = []
    for name in files:
        js = os.path.join(encode_corpus_dir, name)
        npy = js[:-4] + "npy"                                   # trainer.py:233
        arr = np.load(npy, mmap_mode="r")
        if arr.ndim != 2:
            raise ValueError(f"{npy}: expected a 2-D array, got shape {arr.shape}")
        if index is None:
            if index_factory is None:
                from .faiss_compat import IndexFlatIP

                index_factory = IndexFlatIP
            index = index_factory(int(arr.shape[1]))
        for r0 in range(0, arr.shape[0], _CHUNK_ROWS):
            index.add(np.ascontiguousarray(arr[r0:r0 + _CHUNK_ROWS], dtype=np.float32))
        ids = read_id_batches(js)
        if len(ids) != arr.shape[0]:
            raise ValueError(f"{js}: {len(ids)} ids for {arr.shape[0]} rows")
        idx.extend(ids)
    return index, idx


def save_docid_order(index_order_dir: str, ep, idx: Sequence) -> str:
    """`{ep}.docid.txt` (trainer.py:246-248)."""
    os.makedirs(index_order_dir, exist_ok=True)
    path = os.path.join(index_order_dir, f"{ep}.docid.txt")
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"id": [x.item() if hasattr(x, "item") else x for x in idx]}, f, ensure_ascii=False)
    return path


def load_docid_order(index_order_dir: str, ep) -> List:
    """Inverse of save_docid_order (trainer.py:258-260)."""
    idx: List = []
    with open(os.path.join(index_order_dir, f"{ep}.docid.txt"), "r", encoding="utf-8") as f:
        for line in f:
            if line.strip():
                idx = json.loads(line)["id"]
    return idx
