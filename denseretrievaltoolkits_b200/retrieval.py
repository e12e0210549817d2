"""Offline dense-retrieval CLI on the B200 store — the drop-in for `DRT/evaluator/retrieval.py`.

Contract kept from the reference (`retrieval.py:56-92`): the flags `--query_reps`,
`--passage_reps` (a glob of pickled `(reps, lookup)` shards), `--batch_size` (default 128;
<= 0 = one call), `--depth` (default 1000), `--save_ranking_to`, `--save_text`, `--quiet`; the
output is either `qid\\tpid\\tscore` text with every query's rows in descending score order
(`retrieval.py:36-42`) or a pickle of `(scores, psg_indices)`; and the module-level names
`search_queries(retriever, q_reps, p_lookup, args)` (`retrieval.py:20`), `write_ranking`,
`pickle_load`, `pickle_save`, `main`.

Upstream the script cannot run: a second `def search_queries` (retrieval.py:31) shadows the
four-argument one `main` calls (retrieval.py:86), and both unpack two values from calls that
return one (retrieval.py:22,24,32; index.py:40).  What is implemented is the evident intent:
search all queries (in `--batch_size` chunks when positive) and map row ids through the
concatenated passage lookup.

Layout here: passage shards are streamed one at a time onto the device (`load_corpus`), so the
host never holds more than one shard; ids of missing results (-1, when depth > corpus size) map
to the empty string.
"""
from __future__ import annotations

import glob
import logging
import pickle
from argparse import ArgumentParser
from typing import Iterator, List, Sequence, Tuple

import numpy as np

from .index import BaseFaissIPRetriever

logger = logging.getLogger(__name__)

# (flag, argparse keywords) — the reference's option table, retrieval.py:57-64
_OPTIONS = (
    ("--query_reps", dict(required=True)),
    ("--passage_reps", dict(required=True)),
    ("--batch_size", dict(type=int, default=128)),
    ("--depth", dict(type=int, default=1000)),
    ("--save_ranking_to", dict(required=True)),
    ("--save_text", dict(action="store_true")),
    ("--quiet", dict(action="store_true")),
)


def build_parser() -> ArgumentParser:
    parser = ArgumentParser(description=__doc__.splitlines()[0])
    for flag, kw in _OPTIONS:
        parser.add_argument(flag, **kw)
    return parser


def pickle_load(path):
    """One `(reps, lookup)` shard as (float array, id list)."""
    with open(path, "rb") as fh:
        pair = pickle.load(fh)
    return np.array(pair[0]), pair[1]


def pickle_save(obj, path):
    with open(path, "wb") as fh:
        pickle.dump(obj, fh)


def iter_passage_shards(pattern: str) -> Iterator[Tuple[np.ndarray, Sequence]]:
    files = sorted(glob.glob(pattern))
    if not files:
        raise FileNotFoundError(f"no passage shard matches {pattern!r}")
    logger.info("%d passage shard file(s) match %s", len(files), pattern)
    for path in files:
        yield pickle_load(path)


def load_corpus(pattern: str, retriever_cls=BaseFaissIPRetriever, quiet: bool = True):
    """Stream every shard into one retriever; returns (retriever, passage ids in row order)."""
    retriever, passage_ids = None, []
    shards = iter_passage_shards(pattern)
    if not quiet:
        from tqdm import tqdm

        shards = tqdm(shards, desc="Loading shards into index")
    for reps, lookup in shards:
        if retriever is None:
            retriever = retriever_cls(reps)        # dimension from the first shard; adds nothing yet
        retriever.add(reps)
        passage_ids.extend(lookup)
    return retriever, passage_ids


def search_queries(retriever, q_reps, p_lookup, args):
    """-> (scores [Q, depth], passage ids [Q, depth] as strings)."""
    if args.batch_size > 0:
        scores, rows = retriever.batch_search_with_scores(q_reps, args.depth, args.batch_size, args.quiet)
    else:
        scores, rows = retriever.search_with_scores(q_reps, args.depth)
    table = np.asarray([str(pid) for pid in p_lookup] + [""], dtype=object)    # row -1 (padding) -> ""
    return np.asarray(scores), table[np.asarray(rows)]


def write_ranking(corpus_indices, corpus_scores, q_lookup, ranking_save_file):
    """`qid<TAB>pid<TAB>score`, one line per hit, each query's hits by descending score."""
    with open(ranking_save_file, "w") as out:
        for qid, row_scores, row_pids in zip(q_lookup, corpus_scores, corpus_indices):
            row_scores = np.asarray(row_scores)
            for j in np.argsort(-row_scores, kind="stable"):
                out.write(f"{qid}\t{row_pids[j]}\t{row_scores[j]}\n")


def main(argv=None, retriever_cls=BaseFaissIPRetriever):
    args = build_parser().parse_args(argv)
    retriever, passage_ids = load_corpus(args.passage_reps, retriever_cls, quiet=args.quiet)
    q_reps, q_lookup = pickle_load(args.query_reps)
    logger.info("searching %d queries to depth %d", len(q_lookup), args.depth)
    all_scores, psg_indices = search_queries(retriever, q_reps, passage_ids, args)
    logger.info("search finished")
    if args.save_text:
        write_ranking(psg_indices, all_scores, q_lookup, args.save_ranking_to)
    else:
        pickle_save((all_scores, psg_indices), args.save_ranking_to)
    return all_scores, psg_indices


if __name__ == "__main__":
    logging.basicConfig(format="%(asctime)s %(levelname)s %(name)s: %(message)s", level=logging.INFO)
    main()
