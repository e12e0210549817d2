"""CPU: host-side logic of the drop-in layer (file formats, id mapping, chunk planning) with the
compute injected from the oracle (tests may use the oracle; the product never does)."""
import ctypes
import os
import pickle

import numpy as np
import pytest

from denseretrievaltoolkits_b200 import _lib, retrieval
from oracle import flat_ip


class OracleRetriever:
    """Same surface as denseretrievaltoolkits_b200.index.BaseFaissIPRetriever, CPU oracle inside."""

    def __init__(self, init_reps):
        self.index = flat_ip.IndexFlatIP(init_reps.shape[1])

    def add(self, p_reps):
        self.index.add(p_reps)

    def search_with_scores(self, q, k=1000):
        return self.index.search(q, k)

    def batch_search_with_scores(self, q, k, batch_size, quiet=False):
        parts = [self.index.search(q[s:s + batch_size], k) for s in range(0, q.shape[0], batch_size)]
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


@pytest.fixture
def shards(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((300, 64)).astype(np.float32)
    q = rng.standard_normal((5, 64)).astype(np.float32)
    lookups = []
    for i, part in enumerate(np.array_split(x, 3)):
        lk = [f"p{i}_{j}" for j in range(part.shape[0])]
        lookups += lk
        pickle.dump((part, lk), open(tmp_path / f"corpus.{i}.pkl", "wb"))
    pickle.dump((q, [f"q{j}" for j in range(5)]), open(tmp_path / "queries.pkl", "wb"))
    return tmp_path, x, q, lookups


def test_retrieval_cli_text_ranking(shards):
    tmp, x, q, lookups = shards
    out = tmp / "rank.tsv"
    retrieval.main(["--query_reps", str(tmp / "queries.pkl"), "--passage_reps", str(tmp / "corpus.*.pkl"),
                    "--depth", "7", "--batch_size", "2", "--save_ranking_to", str(out), "--save_text", "--quiet"],
                   retriever_cls=OracleRetriever)
    D, I = flat_ip.flat_ip_search(x, q, 7)
    lines = [l.rstrip("\n").split("\t") for l in open(out)]
    assert len(lines) == 5 * 7
    for qi in range(5):
        rows = lines[qi * 7:(qi + 1) * 7]
        assert [r[0] for r in rows] == [f"q{qi}"] * 7
        assert [r[1] for r in rows] == [lookups[i] for i in I[qi]]
        np.testing.assert_allclose([float(r[2]) for r in rows], D[qi], rtol=1e-6)


def test_retrieval_cli_pickle_and_single_call(shards):
    tmp, x, q, lookups = shards
    out = tmp / "rank.pkl"
    retrieval.main(["--query_reps", str(tmp / "queries.pkl"), "--passage_reps", str(tmp / "corpus.*.pkl"),
                    "--depth", "400", "--batch_size", "0", "--save_ranking_to", str(out), "--quiet"],
                   retriever_cls=OracleRetriever)
    scores, psg = pickle.load(open(out, "rb"))
    assert scores.shape == (5, 400) and psg.shape == (5, 400)
    assert (psg[:, 300:] == "").all()              # depth > corpus: padding ids map to ""
    D, I = flat_ip.flat_ip_search(x, q, 400)
    assert [lookups[i] for i in I[0, :300]] == list(psg[0, :300])


def test_parser_defaults_match_reference():
    a = retrieval.build_parser().parse_args(["--query_reps", "q", "--passage_reps", "p", "--save_ranking_to", "o"])
    assert (a.batch_size, a.depth, a.save_text, a.quiet) == (128, 1000, False, False)   # retrieval.py:60-63


def _plan(ntotal, seg_rows, k, attempt):
    lib = _lib.load()
    buf = (ctypes.c_int64 * (3 * 100000))()
    n = lib.drt_plan_chunks(ntotal, seg_rows, k, attempt, buf, 100000)
    assert n >= 0
    kp, cap = ctypes.c_int(), ctypes.c_int()
    assert lib.drt_plan_params(k, attempt, ctypes.byref(kp), ctypes.byref(cap)) == 0
    return [tuple(buf[3 * i:3 * i + 3]) for i in range(n)], kp.value, cap.value


@pytest.mark.parametrize("ntotal,seg_rows,k", [(8_800_000, 1 << 20, 100), (8_800_000, 1 << 20, 1000), (21_000_000, 1 << 20, 100),
                                               (5, 256, 8), (100_000, 4096, 2048), (1 << 20, 1 << 20, 1), (300_001, 8192, 200)])
@pytest.mark.parametrize("attempt", [0, 1, 2])
def test_chunk_plan_covers_corpus_exactly_once(ntotal, seg_rows, k, attempt):
    chunks, kp, cap = _plan(ntotal, seg_rows, k, attempt)
    assert kp > k and cap >= 2 * kp and cap & (cap - 1) == 0
    pos = 0
    for seg, r0, r1 in chunks:
        assert seg * seg_rows + r0 == pos and r0 < r1 <= seg_rows
        assert r0 % 256 == 0                              # tile aligned
        pos = seg * seg_rows + r1
    assert pos == ntotal
    assert chunks[0][2] - chunks[0][1] <= cap // 2        # first chunk: every row is admitted
    if attempt == 2:                                      # fixed chunks can never overflow the buffer
        assert all(r1 - r0 <= cap - kp for _, r0, r1 in chunks[1:])


def test_kprime_margin():
    for k, want_min in [(1, 29), (10, 38), (100, 128), (1000, 1125), (2048, 2304)]:
        _, kp, cap = _plan(1000, 256, k, 0)
        assert kp >= want_min and kp % 4 == 0


def test_faiss_compat_index_file_layout(tmp_path):
    """write_index/read_index byte layout (faiss IndexFlat format) checked without a device by
    driving the module-level writer with a stand-in index object."""
    import struct

    from denseretrievaltoolkits_b200 import faiss_compat

    class Fake:
        d, ntotal = 64, 10
        rows = np.arange(640, dtype=np.float32).reshape(10, 64)

        def reconstruct_n(self, r0, n):
            return self.rows[r0:r0 + n]

    p = tmp_path / "index.faiss"
    faiss_compat.write_index(Fake(), str(p))
    raw = open(p, "rb").read()
    assert raw[:4] == b"IxFI"
    d, n, _, _, trained, metric = struct.unpack("<iqqqBi", raw[4:4 + 33])
    assert (d, n, trained, metric) == (64, 10, 1, 0)
    (count,) = struct.unpack("<Q", raw[37:45])
    assert count == 640 and len(raw) == 45 + 640 * 4
    np.testing.assert_array_equal(np.frombuffer(raw[45:], np.float32), Fake.rows.ravel())


def test_shard_depth_rule_and_peer_buffer_layout():
    """Host logic of the sharded store: mean + 6 sigma per-shard depth from the largest shard's row
    share, and the symmetric-buffer layout of the peer-memory exchange (256-byte aligned regions)."""
    from denseretrievaltoolkits_b200.store import ShardedCorpusStore, _PeerExchange

    class _Fake:
        device = None

        def __init__(self):
            self.ntotal = 0

        def add(self, x):
            self.ntotal += len(x)

    def store(sizes):
        st = ShardedCorpusStore(8, num_virtual_shards=len(sizes), _test_index_factory=_Fake)
        for g, n in enumerate(sizes):
            st.add(np.zeros((n, 8), np.float32), shard=g)
        st.finalize()
        return st

    even8 = store([1000] * 8)
    assert [even8.local_depth(k) for k in (10, 100, 200, 1000, 2048)] == [10, 40, 64, 192, 352]
    assert store([500, 500]).local_depth(100) == 88 and store([250] * 4).local_depth(100) == 56
    assert store([1000]).local_depth(100) == 100                      # one shard: full depth
    skew = store([5700, 4300])                                        # the larger shard owns 57 %
    assert skew.local_depth(200) == 168 and skew.local_depth(20) == 20
    for k in (1, 7, 100, 1000):                                       # always enough entries for a global top-k
        for st in (even8, skew):
            kl = st.local_depth(k)
            assert kl <= k and kl * st.world >= k
    even8._reduce_depth = False
    assert even8.local_depth(100) == 100
    offs, total = _PeerExchange._layout(6980, 40, 100)
    sizes = [6980 * 40 * 4, 6980 * 40 * 8, 6980 * 100 * 4, 6980 * 100 * 8, 6980]
    assert offs[0] == 0 and all(o % 256 == 0 for o in offs) and total % 256 == 0
    assert all(offs[i] + sizes[i] <= offs[i + 1] for i in range(4)) and offs[4] + sizes[4] <= total


def test_deferred_search_returns_per_step_slices_in_order():
    """DeferredSearch (batching Trainer.evaluate's per-step searches, trainer.py:287-297): same
    rows as per-step calls, in submission order, one search per `max_queries`."""
    from denseretrievaltoolkits_b200.deferred import DeferredSearch

    rng = np.random.default_rng(3)
    x = rng.standard_normal((500, 16)).astype(np.float32)

    class Brute:
        calls = 0

        def search(self, q, k):
            Brute.calls += 1
            s = q @ x.T
            I = np.argsort(-s, axis=1, kind="stable")[:, :k]
            return np.take_along_axis(s, I, axis=1), I

    steps = [rng.standard_normal((n, 16)).astype(np.float32) for n in (16, 16, 7, 16, 16, 3)]
    ds = DeferredSearch(Brute(), k=5, max_queries=40)
    got = list(ds.results((i, q) for i, q in enumerate(steps)))
    assert [t for t, _ in got] == list(range(len(steps)))
    assert Brute.calls == 2 and ds.searches == 2 and len(ds) == 0          # 55 queued -> flush at >= 40, then the tail
    for (tag, (D, I)), q in zip(got, steps):
        Dr, Ir = Brute().search(q, 5)
        np.testing.assert_array_equal(I, Ir)
        np.testing.assert_array_equal(D, Dr)
    assert ds.flush() == []
    with pytest.raises(RuntimeError):
        ds.add(np.zeros(16, np.float32))


def test_write_negatives_jsonl_matches_the_reference_file_layout(tmp_path, golden_dir):
    """f3: the mined-negatives file.  Golden = the bytes the reference's own BM25Negatives.save
    wrote for these samples (sampler.py:89-99; its load_passages cache branch, sampler.py:57-66,
    read them back in the generator).  `write_negatives_jsonl` must produce the same bytes, and the
    reference's read loop (json.loads per line) must give back the records QPCollator consumes."""
    import json

    from denseretrievaltoolkits_b200.mining import write_negatives_jsonl

    g = json.load(open(os.path.join(golden_dir, "mining_samples.json"), encoding="utf-8"))
    out = tmp_path / "bm25negatives"
    write_negatives_jsonl(str(out), g["samples"], np.asarray(g["neg_ids"], dtype=np.int64), g["passages"])
    want = open(os.path.join(golden_dir, "bm25negatives.jsonl"), "rb").read()
    assert out.read_bytes() == want
    data = []
    with open(out, "r", encoding="utf-8") as f:                      # sampler.py:61-64
        for line in f.readlines():
            data.append(json.loads(line))
    assert len(data) == len(g["samples"])
    for rec, smp, row in zip(data, g["samples"], g["neg_ids"]):
        assert rec["query"] == smp["query"] and rec["positives"] == smp["positives"]
        assert rec["negatives"] == [g["passages"][j] for j in row if j >= 0]
