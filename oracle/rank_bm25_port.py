"""Restatement of rank-bm25 == 0.2.2 (BM25Okapi / BM25Plus), the un-vendored dependency behind
reference src/core/retrievers/sparse.py:16,92,96,177 (pinned in poetry.lock:4846-4847, requirements.txt:29).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Two layers:

* ``BM25Okapi`` / ``BM25Plus``  -- a faithful restatement of the published classes (list-of-dicts doc_freqs, per-term
  Python list comprehension + NumPy vector arithmetic).  This is what gets installed as ``sys.modules['rank_bm25']``
  when the reference's own sparse.py is imported, and it is the timed CPU baseline ("port").
* ``FastBM25``  -- the same arithmetic over a CSR inverted index (what a CPU implementation would do if it cared);
  tests assert it is BIT-IDENTICAL to the faithful classes, then use it where the faithful one is too slow.
"""
from __future__ import annotations

import math

import numpy as np


class BM25:
    def __init__(self, corpus, tokenizer=None):
        self.corpus_size = 0
        self.avgdl = 0
        self.doc_freqs = []
        self.idf = {}
        self.doc_len = []
        self.tokenizer = tokenizer
        if tokenizer:
            corpus = [tokenizer(doc) for doc in corpus]
        nd = self._initialize(corpus)
        self._calc_idf(nd)

    def _initialize(self, corpus):
        nd = {}  # word -> number of documents containing it (insertion order = first occurrence)
        num_doc = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_doc += len(document)
            frequencies = {}
            for word in document:
                frequencies[word] = frequencies.get(word, 0) + 1
            self.doc_freqs.append(frequencies)
            for word in frequencies:
                nd[word] = nd.get(word, 0) + 1
            self.corpus_size += 1
        self.avgdl = num_doc / self.corpus_size
        return nd

    def _calc_idf(self, nd):
        raise NotImplementedError

    def get_scores(self, query):
        raise NotImplementedError

    def get_top_n(self, query, documents, n=5):
        assert self.corpus_size == len(documents)
        scores = self.get_scores(query)
        top_n = np.argsort(scores)[::-1][:n]
        return [documents[i] for i in top_n]


class BM25Okapi(BM25):
    def __init__(self, corpus, tokenizer=None, k1=1.5, b=0.75, epsilon=0.25):
        self.k1 = k1
        self.b = b
        self.epsilon = epsilon
        super().__init__(corpus, tokenizer)

    def _calc_idf(self, nd):
        idf_sum = 0
        negative_idfs = []
        for word, freq in nd.items():
            idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = idf
            idf_sum += idf
            if idf < 0:
                negative_idfs.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative_idfs:
            self.idf[word] = eps

    def get_scores(self, query):
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (q_freq * (self.k1 + 1) /
                                               (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


class BM25Plus(BM25):
    def __init__(self, corpus, tokenizer=None, k1=1.5, b=0.75, delta=1):
        self.k1 = k1
        self.b = b
        self.delta = delta
        super().__init__(corpus, tokenizer)

    def _calc_idf(self, nd):
        for word, freq in nd.items():
            idf = math.log(self.corpus_size + 1) - math.log(freq)
            self.idf[word] = idf

    def get_scores(self, query):
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (self.delta + (q_freq * (self.k1 + 1)) /
                                               (self.k1 * (1 - self.b + self.b * doc_len / self.avgdl) + q_freq))
        return score


class FastBM25:
    """Same arithmetic as the classes above over a CSR inverted index (term-major).

    ``indptr[V+1]``, ``post_doc[nnz]`` (ascending per term), ``post_tf[nnz]``, ``doc_len[N]``, ``idf[V]``, ``avgdl``.
    get_scores(term_ids) reproduces ``score += idf * (...)`` term by term, in query order, touching only postings
    (documents without the term add exactly +0.0 / idf*delta in the dense formulation).
    """

    def __init__(self, indptr, post_doc, post_tf, doc_len, idf, avgdl, variant="okapi", k1=1.5, b=0.75, delta=1):
        self.indptr = np.asarray(indptr, dtype=np.int64)
        self.post_doc = np.asarray(post_doc, dtype=np.int64)
        self.post_tf = np.asarray(post_tf, dtype=np.int64)
        self.doc_len = np.asarray(doc_len, dtype=np.int64)
        self.idf = np.asarray(idf, dtype=np.float64)
        self.avgdl = avgdl
        self.variant = variant
        self.k1, self.b, self.delta = k1, b, delta
        self.n = len(self.doc_len)
        # k1 * (1 - b + b * doc_len / avgdl)  -- identical expression / evaluation order to get_scores above
        self.dnorm = self.k1 * (1 - self.b + self.b * self.doc_len / self.avgdl)

    def get_scores(self, term_ids):
        score = np.zeros(self.n)
        for t in term_ids:
            if t is None or t < 0:
                continue  # unknown token: (self.idf.get(q) or 0) == 0 -> adds zeros
            idf = float(self.idf[t])
            if idf == 0.0:
                continue
            lo, hi = self.indptr[t], self.indptr[t + 1]
            docs = self.post_doc[lo:hi]
            tf = self.post_tf[lo:hi]
            if self.variant == "plus":
                contrib = np.full(self.n, idf * (self.delta + 0.0))
                contrib[docs] = idf * (self.delta + (tf * (self.k1 + 1)) / (self.dnorm[docs] + tf))
                score += contrib
            else:
                score[docs] += idf * (tf * (self.k1 + 1) / (tf + self.dnorm[docs]))
        return score
