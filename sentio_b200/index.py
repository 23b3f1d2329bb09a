"""Host-side index builders: text / token streams -> the arrays the C ABI (include/sentio_b200.h) consumes.

* ``Bm25IndexData``  -- vocabulary, term-major CSR postings, doc lengths, idf table (replaces the state that
  ``BM25Retriever.index`` builds through rank_bm25: reference src/core/retrievers/sparse.py:70-100).
  Term ids are assigned in FIRST-OCCURRENCE order so that rank_bm25's insertion-ordered ``idf_sum`` (and therefore the
  epsilon floor of negative idfs) is reproduced bit for bit.
* ``hash_tokenize_pairs`` -- deterministic word-piece stand-in for the cross-encoder inputs (SURVEY.md Appendix C;
  no tokenizer vocabulary exists offline).
"""
from __future__ import annotations

import math
import os
import pickle
import zlib
from dataclasses import dataclass, field
from typing import Iterable, Sequence

import numpy as np

__all__ = ["Bm25IndexData", "build_bm25_from_texts", "build_bm25_from_token_ids", "finish_gpu_built_index",
           "tokenize", "tokenize_texts", "hash_tokenize_pairs"]


def tokenize(text: str) -> list[str]:
    """The reference tokeniser: ``text.lower().split()`` (sparse.py:88,174)."""
    return text.lower().split()


@dataclass
class Bm25IndexData:
    variant: str
    k1: float
    b: float
    epsilon: float
    delta: float
    n_docs: int
    avgdl: float
    indptr: np.ndarray      # int64 [V+1]
    post_doc: np.ndarray    # int32 [nnz]
    post_tf: np.ndarray     # uint16 [nnz]
    doc_len: np.ndarray     # int32 [N]
    idf: np.ndarray         # float64 [V]
    vocab: dict | None = None           # token string -> term id (None for pre-tokenised integer corpora)
    token_id_map: np.ndarray | None = None  # for integer corpora: raw token id -> term id (-1 = unseen)
    average_idf: float = 0.0
    extras: dict = field(default_factory=dict)

    @property
    def n_terms(self) -> int:
        return int(len(self.idf))

    def term_ids(self, tokens: Sequence) -> np.ndarray:
        """Query tokens -> term ids (-1 for unknown), duplicates kept (they are scored twice, like the reference)."""
        if self.vocab is not None:
            return np.fromiter((self.vocab.get(t, -1) for t in tokens), dtype=np.int32, count=len(tokens))
        raw = np.asarray(tokens, dtype=np.int64)
        out = np.full(len(raw), -1, dtype=np.int32)
        ok = (raw >= 0) & (raw < len(self.token_id_map))
        out[ok] = self.token_id_map[raw[ok]]
        return out

    def shard(self, lo: int, hi: int) -> "Bm25IndexData":
        """Postings restricted to docs [lo, hi) with shard-local doc indices; idf / avgdl stay corpus-global."""
        sel = (self.post_doc >= lo) & (self.post_doc < hi)
        term_of = np.repeat(np.arange(self.n_terms, dtype=np.int64), np.diff(self.indptr))
        counts = np.bincount(term_of[sel], minlength=self.n_terms).astype(np.int64)
        indptr = np.zeros(self.n_terms + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])
        return Bm25IndexData(
            variant=self.variant, k1=self.k1, b=self.b, epsilon=self.epsilon, delta=self.delta, n_docs=hi - lo,
            avgdl=self.avgdl, indptr=indptr, post_doc=(self.post_doc[sel] - lo).astype(np.int32),
            post_tf=self.post_tf[sel], doc_len=self.doc_len[lo:hi].copy(), idf=self.idf, vocab=self.vocab,
            token_id_map=self.token_id_map, average_idf=self.average_idf)

    # persistence: the reference pickles its rank_bm25 object (sparse.py:102-157); we persist the CSR arrays
    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            pickle.dump({"format": "sentio_b200.bm25.v1", **self.__dict__}, f, protocol=pickle.HIGHEST_PROTOCOL)

    @staticmethod
    def load(path: str) -> "Bm25IndexData":
        with open(path, "rb") as f:
            d = pickle.load(f)
        if d.pop("format", None) != "sentio_b200.bm25.v1":
            raise ValueError(f"{path} is not a sentio_b200 BM25 index")
        return Bm25IndexData(**d)


def _idf_table(df: np.ndarray, n_docs: int, variant: str, epsilon: float) -> tuple[np.ndarray, float]:
    """rank_bm25 0.2.2 ``_calc_idf`` (math.log on Python numbers, sequential idf_sum in vocabulary order)."""
    log = math.log
    V = len(df)
    if variant == "plus":
        ln1 = log(n_docs + 1)
        idf = np.fromiter((ln1 - log(int(f)) for f in df), dtype=np.float64, count=V)
        return idf, 0.0
    vals = [log(n_docs - int(f) + 0.5) - log(int(f) + 0.5) for f in df]
    idf_sum = 0
    for v in vals:  # plain left-to-right float accumulation, like the reference loop
        idf_sum += v
    average_idf = idf_sum / V if V else 0.0
    eps = epsilon * average_idf
    idf = np.array([eps if v < 0 else v for v in vals], dtype=np.float64)
    return idf, average_idf


def global_bm25_stats(shard_term_tokens, shard_dfs, shard_n_docs, shard_n_tokens, variant: str, epsilon: float):
    """Corpus-global BM25 statistics from per-shard GPU builds (contiguous doc ranges, shards in corpus order).

    ``shard_term_tokens[r]`` = raw token of every local term of shard r in LOCAL first-occurrence order, ``shard_dfs[r]`` the
    matching document frequencies.  The global vocabulary order (which fixes the summation order of rank_bm25's average
    idf) is the corpus' first-occurrence order = shard 0's terms, then the terms first seen in shard 1, ...  Returns
    (idf_of_token: dict raw token -> idf, average_idf, n_docs, avgdl): bit-identical to ``_idf_table`` on the whole corpus."""
    order, df_of = [], {}
    for toks, dfs in zip(shard_term_tokens, shard_dfs):
        for t, f in zip(np.asarray(toks).tolist(), np.asarray(dfs).tolist()):
            if t in df_of:
                df_of[t] += f
            else:
                df_of[t] = f
                order.append(t)
    n_docs = int(sum(shard_n_docs))
    df = np.fromiter((df_of[t] for t in order), dtype=np.int64, count=len(order))
    idf, average_idf = _idf_table(df, n_docs, variant, epsilon)
    return dict(zip(order, idf.tolist())), average_idf, n_docs, float(sum(shard_n_tokens)) / n_docs


def finish_gpu_built_index(df: np.ndarray, term_token: np.ndarray, n_docs: int, n_tokens: int, variant: str, k1: float,
                           b: float, epsilon: float, delta: float, csr=None, global_stats=None) -> Bm25IndexData:
    """Host half of the GPU index build (engine.build_bm25_gpu): idf table from the device-computed df (math.log, bit for
    bit like rank_bm25) + raw token -> term id map.  ``csr`` = (indptr, post_doc, post_tf, doc_len) when exported.
    ``global_stats`` = (idf_of_token, average_idf, avgdl) for a corpus SHARD: idf / avgdl stay corpus-global."""
    if global_stats is not None:
        idf_of_token, average_idf, avgdl = global_stats
        idf = np.fromiter((idf_of_token[int(t)] for t in term_token), dtype=np.float64, count=len(term_token))
    else:
        idf, average_idf = _idf_table(df, n_docs, variant, epsilon)
        avgdl = n_tokens / n_docs
    size = int(term_token.max()) + 1 if len(term_token) else 0
    token_id_map = np.full(size, -1, dtype=np.int32)
    token_id_map[term_token] = np.arange(len(term_token), dtype=np.int32)
    empty = (np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.uint16), np.zeros(0, np.int32))
    indptr, post_doc, post_tf, doc_len = csr if csr is not None else empty
    return Bm25IndexData(variant=variant, k1=float(k1), b=float(b), epsilon=float(epsilon), delta=float(delta),
                         n_docs=int(n_docs), avgdl=avgdl, indptr=indptr, post_doc=post_doc, post_tf=post_tf,
                         doc_len=doc_len, idf=idf, vocab=None, token_id_map=token_id_map, average_idf=average_idf,
                         extras={"built_on": "gpu", "postings_on_host": csr is not None})


def build_bm25_from_token_ids(flat_tokens: np.ndarray, doc_offsets: np.ndarray, variant: str = "okapi", k1: float = 1.5,
                              b: float = 0.75, epsilon: float = 0.25, delta: float = 1.0, vocab_tokens=None) -> Bm25IndexData:
    """CSR build from an integer token stream (doc i = flat_tokens[doc_offsets[i]:doc_offsets[i+1]])."""
    variant = variant.lower()
    flat = np.asarray(flat_tokens, dtype=np.int64)
    off = np.asarray(doc_offsets, dtype=np.int64)
    n_docs = len(off) - 1
    if n_docs <= 0:
        raise ValueError("empty corpus")
    doc_len = np.diff(off)
    total = int(off[-1])
    avgdl = total / n_docs  # Python: num_doc / corpus_size (int / int -> correctly rounded float)
    # first-occurrence relabel: term id = rank of the token's first position in the stream
    uniq, first_pos = np.unique(flat, return_index=True)
    order = np.argsort(first_pos, kind="stable")
    rank_of_uniq = np.empty(len(uniq), dtype=np.int64)
    rank_of_uniq[order] = np.arange(len(uniq))
    term = rank_of_uniq[np.searchsorted(uniq, flat)]
    V = len(uniq)
    doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), doc_len)
    key = term * n_docs + doc_of
    key.sort()
    if len(key):
        boundary = np.empty(len(key), dtype=bool)
        boundary[0] = True
        np.not_equal(key[1:], key[:-1], out=boundary[1:])
        starts = np.flatnonzero(boundary)
        ukey = key[starts]
        tf = np.diff(np.append(starts, len(key)))
    else:
        ukey = key
        tf = np.zeros(0, dtype=np.int64)
    if len(tf) and tf.max() > 65535:
        raise ValueError("term frequency above 65535 is not representable in the uint16 postings")
    post_term = ukey // n_docs
    post_doc = (ukey - post_term * n_docs).astype(np.int32)
    df = np.bincount(post_term, minlength=V).astype(np.int64)
    indptr = np.zeros(V + 1, dtype=np.int64)
    np.cumsum(df, out=indptr[1:])
    idf, average_idf = _idf_table(df, n_docs, variant, epsilon)
    token_id_map = None
    vocab = None
    if vocab_tokens is not None:
        vocab = {vocab_tokens[int(u)]: int(r) for u, r in zip(uniq, rank_of_uniq)}
    else:
        size = int(uniq.max()) + 1 if len(uniq) else 0
        token_id_map = np.full(size, -1, dtype=np.int32)
        token_id_map[uniq] = rank_of_uniq.astype(np.int32)
    return Bm25IndexData(variant=variant, k1=float(k1), b=float(b), epsilon=float(epsilon), delta=float(delta),
                         n_docs=n_docs, avgdl=avgdl, indptr=indptr, post_doc=post_doc, post_tf=tf.astype(np.uint16),
                         doc_len=doc_len.astype(np.int32), idf=idf, vocab=vocab, token_id_map=token_id_map,
                         average_idf=average_idf)


def tokenize_texts(texts: Iterable[str]):
    """``text.lower().split()`` (sparse.py:88) over a corpus -> (vocab token->id in first-occurrence order,
    flat int32 token ids, int64 doc offsets): the string half of index building, which stays on the host."""
    vocab: dict[str, int] = {}
    flat: list[int] = []
    offsets = [0]
    get = vocab.get
    for text in texts:
        for tok in text.lower().split():
            tid = get(tok)
            if tid is None:
                tid = len(vocab)
                vocab[tok] = tid
            flat.append(tid)
        offsets.append(len(flat))
    return vocab, np.asarray(flat, dtype=np.int32), np.asarray(offsets, dtype=np.int64)


def build_bm25_from_texts(texts: Iterable[str], variant: str = "okapi", k1: float = 1.5, b: float = 0.75,
                          epsilon: float = 0.25, delta: float = 1.0) -> Bm25IndexData:
    """``text.lower().split()`` tokenisation (sparse.py:88) -> vocabulary in first-occurrence order -> CSR."""
    vocab: dict[str, int] = {}
    flat: list[int] = []
    offsets = [0]
    get = vocab.get
    for text in texts:
        for tok in text.lower().split():
            tid = get(tok)
            if tid is None:
                tid = len(vocab)
                vocab[tok] = tid
            flat.append(tid)
        offsets.append(len(flat))
    data = build_bm25_from_token_ids(np.asarray(flat, dtype=np.int64), np.asarray(offsets, dtype=np.int64), variant, k1,
                                     b, epsilon, delta)
    # ids were already first-occurrence ordered, so the relabel is the identity; keep the string vocabulary
    data.vocab = vocab
    data.token_id_map = None
    return data


# ----------------------------------------------------------------------------------------------- cross-encoder input
CLS_ID, SEP_ID, PAD_ID = 101, 102, 0


def _hash_token(tok: str) -> int:
    return 1000 + zlib.crc32(tok.encode("utf-8")) % 29522


def hash_vocab_ids(n_vocab: int, prefix: str = "w") -> np.ndarray:
    """Hashed word-piece id of every synthetic vocabulary token ``w0 .. w{n-1}`` (uint16; same hash as below)."""
    return np.fromiter((_hash_token(f"{prefix}{t}") for t in range(n_vocab)), dtype=np.uint16, count=n_vocab)


def doc_token_matrix(flat_tokens: np.ndarray, doc_offsets: np.ndarray, vocab_ids: np.ndarray, ld: int = 120):
    """Pre-tokenised documents for the batched rerank path: (tok uint16 [N, ld], len int32 [N]) from an integer token
    stream (doc i = flat_tokens[off[i]:off[i+1]], truncated to ld).  Blocks of 256 k docs are filled by a small thread
    pool (NumPy's gather / scatter release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    off = np.asarray(doc_offsets, dtype=np.int64)
    flat = np.asarray(flat_tokens)
    n = len(off) - 1
    lens = np.minimum(np.diff(off), ld).astype(np.int32)
    tok = np.zeros((n, ld), dtype=np.uint16)
    col = np.arange(ld)[None, :]

    def block(lo):
        hi = min(n, lo + 262144)
        mask = col < lens[lo:hi, None]
        src = (off[lo:hi, None] + col)[mask]
        tok[lo:hi][mask] = vocab_ids[flat[src]]

    with ThreadPoolExecutor(max_workers=min(16, max(1, (os.cpu_count() or 1) // 2))) as ex:
        list(ex.map(block, range(0, n, 262144)))
    return tok, lens


def hash_tokenize_pairs(query: str, docs: Sequence[str], seq_len: int = 128):
    """``[CLS] q [SEP] d [SEP]`` framing with crc32-hashed word ids, padded/truncated to ``seq_len``.

    Returns (input_ids int32 [P,S], token_type int32 [P,S], lengths int32 [P]).  Query keeps at most seq_len//2 - 2
    tokens; the document fills the rest.
    """
    q_ids = [_hash_token(t) for t in query.lower().split()][: max(1, seq_len // 2 - 2)]
    P = len(docs)
    ids = np.full((P, seq_len), PAD_ID, dtype=np.int32)
    tt = np.zeros((P, seq_len), dtype=np.int32)
    lens = np.zeros(P, dtype=np.int32)
    head = [CLS_ID, *q_ids, SEP_ID]
    room = seq_len - len(head) - 1
    for i, text in enumerate(docs):
        d_ids = [_hash_token(t) for t in (text or "").lower().split()][: max(0, room)]
        row = head + d_ids + [SEP_ID]
        n = len(row)
        ids[i, :n] = row
        tt[i, len(head):n] = 1
        lens[i] = n
    return ids, tt, lens
