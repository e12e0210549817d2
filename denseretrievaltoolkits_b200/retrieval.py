"""Mirror of `DRT/evaluator/retrieval.py` (the offline search CLI) on the B200 store.

Same flags and file formats as the reference (`retrieval.py:56-92`): `--passage_reps` is a glob
of pickled `(reps, lookup)` shards, `--query_reps` one such pickle, `--depth` (default 1000),
`--batch_size` (default 128; <= 0 = one call), output either `qid\\tpid\\tscore` text
(`--save_text`, rows sorted by descending score, retrieval.py:36-42) or a pickle of
`(scores, psg_indices)`.

Upstream this script cannot run: a second `def search_queries` (retrieval.py:31) shadows the
four-argument one its `main` calls (retrieval.py:86), and both unpack two values from calls
that return one (retrieval.py:22,24,32; index.py:40).  The intended contract — search all
queries (optionally in `--batch_size` chunks), map faiss row ids through `p_lookup` — is what
is implemented here, with the call signature `search_queries(retriever, q_reps, p_lookup, args)`
of retrieval.py:20 unchanged.
"""
from __future__ import annotations

import glob
import logging
import pickle
from argparse import ArgumentParser
from itertools import chain

import numpy as np

from .index import BaseFaissIPRetriever

logger = logging.getLogger(__name__)


def search_queries(retriever, q_reps, p_lookup, args):
    if args.batch_size > 0:
        all_scores, all_indices = retriever.batch_search_with_scores(q_reps, args.depth, args.batch_size, args.quiet)
    else:
        all_scores, all_indices = retriever.search_with_scores(q_reps, args.depth)
    all_scores = np.asarray(all_scores)
    all_indices = np.asarray(all_indices)
    lookup = np.asarray([str(x) for x in p_lookup] + [""], dtype=object)   # id -1 (padding) -> ""
    psg_indices = lookup[all_indices]
    return all_scores, psg_indices


def write_ranking(corpus_indices, corpus_scores, q_lookup, ranking_save_file):
    with open(ranking_save_file, "w") as f:
        for qid, q_doc_scores, q_doc_indices in zip(q_lookup, corpus_scores, corpus_indices):
            order = np.argsort(-np.asarray(q_doc_scores), kind="stable")
            for j in order:
                f.write(f"{qid}\t{q_doc_indices[j]}\t{q_doc_scores[j]}\n")


def pickle_load(path):
    with open(path, "rb") as f:
        reps, lookup = pickle.load(f)
    return np.array(reps), lookup


def pickle_save(obj, path):
    with open(path, "wb") as f:
        pickle.dump(obj, f)


def build_parser() -> ArgumentParser:
    parser = ArgumentParser()
    parser.add_argument("--query_reps", required=True)
    parser.add_argument("--passage_reps", required=True)
    parser.add_argument("--batch_size", type=int, default=128)
    parser.add_argument("--depth", type=int, default=1000)
    parser.add_argument("--save_ranking_to", required=True)
    parser.add_argument("--save_text", action="store_true")
    parser.add_argument("--quiet", action="store_true")
    return parser


def main(argv=None, retriever_cls=BaseFaissIPRetriever):
    args = build_parser().parse_args(argv)
    index_files = sorted(glob.glob(args.passage_reps))
    if not index_files:
        raise FileNotFoundError(f"no passage shard matches {args.passage_reps!r}")
    logger.info(f"Pattern match found {len(index_files)} files; loading them into index.")

    p_reps_0, p_lookup_0 = pickle_load(index_files[0])
    retriever = retriever_cls(p_reps_0)
    shards = chain([(p_reps_0, p_lookup_0)], map(pickle_load, index_files[1:]))
    look_up = []
    for p_reps, p_lookup in shards:
        retriever.add(p_reps)
        look_up += list(p_lookup)

    q_reps, q_lookup = pickle_load(args.query_reps)
    logger.info("Index Search Start")
    all_scores, psg_indices = search_queries(retriever, q_reps, look_up, args)
    logger.info("Index Search Finished")

    if args.save_text:
        write_ranking(psg_indices, all_scores, q_lookup, args.save_ranking_to)
    else:
        pickle_save((all_scores, psg_indices), args.save_ranking_to)
    return all_scores, psg_indices


if __name__ == "__main__":
    logging.basicConfig(format="%(asctime)s - %(levelname)s - %(name)s -   %(message)s",
                        datefmt="%m/%d/%Y %H:%M:%S", level=logging.INFO)
    main()
