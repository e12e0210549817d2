"""CPU: the reference's on-disk corpus formats (.npy + id JSONL per rank, docid.txt) written the
way DRT/trainer/trainer.py:210-216,246-248 writes them and read back the way :222-241,258-260
reads them; the index is the CPU oracle (injected), the product default is the CUDA store."""
import json
import os

import numpy as np
import pytest

from denseretrievaltoolkits_b200 import store_io
from oracle import flat_ip


def _reference_style_write(dirname, ep, rank, reps, id_batches):
    """What Trainer._encoding_corpus does, verbatim in effect (np.save + json.dump per batch)."""
    np.save(os.path.join(dirname, f"{ep}.{rank}.npy"), reps)
    with open(os.path.join(dirname, f"{ep}.{rank}.json"), "w", encoding="utf-8") as f:
        for b in id_batches:
            json.dump({"id": b}, f, ensure_ascii=False)
            f.write("\n")


def test_load_corpus_written_by_the_reference(tmp_path):
    rng = np.random.default_rng(0)
    d = 64
    all_rows, all_ids = {}, {}
    for rank, n in enumerate([50, 33, 41]):
        reps = rng.standard_normal((n, d)).astype(np.float32)
        ids = [f"doc{rank}_{i}" for i in range(n)]
        batches = [ids[i:i + 16] for i in range(0, n, 16)]
        _reference_style_write(str(tmp_path), 3, rank, reps, batches)
        all_rows[f"3.{rank}.json"], all_ids[f"3.{rank}.json"] = reps, ids
    _reference_style_write(str(tmp_path), 4, 0, rng.standard_normal((5, d)).astype(np.float32), [["x"] * 5])   # other epoch
    order = store_io.list_reference_shards(str(tmp_path), 3)
    assert sorted(order) == ["3.0.json", "3.1.json", "3.2.json"]
    index, idx = store_io.load_reference_corpus(str(tmp_path), 3, index_factory=flat_ip.IndexFlatIP)
    want_rows = np.concatenate([all_rows[f] for f in order])       # listing order = faiss row order
    want_ids = sum((all_ids[f] for f in order), [])
    assert index.ntotal == 124 and idx == want_ids
    np.testing.assert_array_equal(index.reconstruct_n(0, 124), want_rows)
    q = rng.standard_normal((4, d)).astype(np.float32)
    D, I = index.search(q, 5)
    Dr, Ir = flat_ip.flat_ip_search(want_rows, q, 5)
    np.testing.assert_array_equal(I, Ir)


def test_save_shard_roundtrip_and_docid_order(tmp_path):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((70, 64)).astype(np.float32)
    index = flat_ip.IndexFlatIP(64)
    index.add(x)
    ids = list(range(1000, 1070))
    npy, js = store_io.save_reference_shard(index, str(tmp_path / "enc"), 7, 2, [ids[:32], ids[32:64], ids[64:]])
    np.testing.assert_array_equal(np.load(npy), x)                                  # what np.load(npy_file) sees (trainer.py:235)
    assert [json.loads(l)["id"] for l in open(js)] == [ids[:32], ids[32:64], ids[64:]]
    index2, idx2 = store_io.load_reference_corpus(str(tmp_path / "enc"), 7, index_factory=flat_ip.IndexFlatIP)
    assert idx2 == ids and index2.ntotal == 70
    with pytest.raises(ValueError):
        store_io.save_reference_shard(index, str(tmp_path / "bad"), 7, 0, [ids[:10]])
    p = store_io.save_docid_order(str(tmp_path / "order"), 7, np.array(ids))
    assert json.load(open(p)) == {"id": ids}                                        # trainer.py:246-248
    assert store_io.load_docid_order(str(tmp_path / "order"), 7) == ids


def test_missing_epoch_raises(tmp_path):
    with pytest.raises(FileNotFoundError):
        store_io.load_reference_corpus(str(tmp_path), 1, index_factory=flat_ip.IndexFlatIP)
