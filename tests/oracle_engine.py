"""OracleEngine -- an oracle-backed stand-in for B200Engine used ONLY by the CPU (`not gpu`) tests to exercise the host
logic (retriever classes, pipeline sharding, gloo all-gather) without a GPU.  It is test infrastructure: it imports
oracle/ and must never be reachable from the product."""
from __future__ import annotations

import numpy as np

from oracle import dense as dense_oracle
from oracle import fusion as fusion_oracle
from oracle import scorers as scorers_oracle
from oracle.rank_bm25_port import FastBM25


class OracleEngine:
    device = None

    def __init__(self):
        self.rows = {}
        self.id_base = {}
        self.dense_dim = {}
        self.bm25 = None
        self.fast = None
        self.bm25_id_base = 0

    def close(self):
        pass

    # K1
    def load_dense(self, vecs, id_base=0, slot=0):
        self.rows[slot] = dense_oracle.stored_rows(np.asarray(vecs))
        self.id_base[slot] = id_base
        self.dense_dim[slot] = self.rows[slot].shape[1]

    def dense_topk(self, q, k, slot=0):
        q = np.atleast_2d(np.asarray(q, dtype=np.float32))
        B = q.shape[0]
        ids = np.full((B, k), -1, np.int64)
        sc = np.zeros((B, k))
        cnt = np.zeros(B, np.int32)
        for b in range(B):
            i, s = dense_oracle.dense_topk(self.rows[slot], q[b], k)
            ids[b, :len(i)] = i + self.id_base[slot]
            sc[b, :len(i)] = s
            cnt[b] = len(i)
        return ids, sc, cnt

    def dense_fetch(self, ids, slot=0):
        return self.rows[slot][np.asarray(ids) - self.id_base[slot]].astype(np.float32)

    # K2
    def load_bm25(self, data, id_base=0):
        self.bm25 = data
        self.bm25_id_base = id_base
        self.fast = FastBM25(data.indptr, data.post_doc, data.post_tf, data.doc_len, data.idf, data.avgdl,
                             data.variant, data.k1, data.b, data.delta)

    def build_bm25_gpu(self, flat_tokens, doc_offsets, variant="okapi", k1=1.5, b=0.75, epsilon=0.25, delta=1.0,
                       id_base=0, export=False, stats_hook=None):
        """Double of B200Engine.build_bm25_gpu: the host builder stands in for the device build (``stats_hook``: the
        shard's idf / avgdl are replaced by the corpus-global ones, like the product does)."""
        from sentio_b200.index import build_bm25_from_token_ids

        data = build_bm25_from_token_ids(flat_tokens, doc_offsets, variant, k1, b, epsilon, delta)
        if stats_hook is not None:
            term_token = np.full(data.n_terms, -1, np.int64)
            known = np.nonzero(data.token_id_map >= 0)[0]
            term_token[data.token_id_map[known]] = known
            idf_of, avg_idf, avgdl = stats_hook(term_token, np.diff(data.indptr), data.n_docs, int(data.doc_len.sum()))
            data.idf = np.asarray([idf_of[int(t)] for t in term_token], dtype=np.float64)
            data.average_idf, data.avgdl = avg_idf, avgdl
        self.load_bm25(data, id_base)
        return data

    def bm25_scores(self, term_ids):
        return self.fast.get_scores(list(term_ids))

    def bm25_topk(self, term_id_lists, k):
        B = len(term_id_lists)
        ids = np.full((B, k), -1, np.int64)
        sc = np.zeros((B, k))
        cnt = np.zeros(B, np.int32)
        for b, terms in enumerate(term_id_lists):
            s = self.fast.get_scores(list(terms))
            order = np.argsort(-s, kind="stable")[:k]
            order = [i for i in order if s[i] > 0]
            ids[b, :len(order)] = np.asarray(order, dtype=np.int64) + self.bm25_id_base
            sc[b, :len(order)] = s[order]
            cnt[b] = len(order)
        return ids, sc, cnt

    # K3
    def fuse(self, method, rrf_k, w_dense, w_sparse, k, dense=None, sparse=None, plugin=None, extra=None):
        def rows(lst, b):
            if lst is None:
                return []
            i, s, c = lst
            return [(int(i[b, j]), float(s[b, j])) for j in range(int(c[b]))]

        B = next(x for x in (dense, sparse, plugin) if x is not None)[0].shape[0]
        ids = np.full((B, k), -1, np.int64)
        sc = np.zeros((B, k))
        src = np.zeros((B, k), np.int32)
        cnt = np.zeros(B, np.int32)
        if int(rrf_k) == rrf_k:
            rrf_k = int(rrf_k)
        for b in range(B):
            d, s, p = rows(dense, b), rows(sparse, b), rows(plugin, b)
            ex = [list(extra[b, e]) for e in range(extra.shape[1])] if extra is not None else None
            out = fusion_oracle.fuse(method, rrf_k, w_dense, w_sparse, d, s, p, k, ex)
            d_set, s_set = {i for i, _ in d}, {i for i, _ in s}
            for j, (doc_id, score, _has) in enumerate(out):
                ids[b, j], sc[b, j] = doc_id, score
                src[b, j] = (1 if doc_id in d_set else 0) | (2 if doc_id in s_set else 0)
            cnt[b] = len(out)
        return ids, sc, src, cnt

    # K4
    def semantic_mmr(self, q, cand=None, cand_ids=None, w_sem=0.7, lambda_=0.7, w_mmr=0.5, want_sem=True,
                     want_mmr=True, slot=0):
        if cand is None:
            cand = self.dense_fetch(cand_ids, slot)
        q64 = np.asarray(q, dtype=np.float32).astype(np.float64)
        c64 = [np.asarray(c, dtype=np.float32).astype(np.float64) for c in cand]
        sem = np.asarray(scorers_oracle.semantic(q64, c64, w_sem)) if want_sem else None
        mmr = np.asarray(scorers_oracle.mmr(q64, c64, lambda_, w_mmr)) if want_mmr else None
        return sem, mmr


    # K5 (NumPy restatement of the forward pass, pinned to HuggingFace BERT in tests/test_oracle_golden.py)
    def ce_load(self, blob, cfg):
        from sentio_b200.cross_encoder import CrossEncoderWeights

        w = CrossEncoderWeights(dict(cfg), {})
        at = 0
        for name, shape in CrossEncoderWeights.tensor_order(cfg):
            n = int(np.prod(shape))
            w.tensors[name] = np.asarray(blob[at:at + n], dtype=np.float32).reshape(shape)
            at += n
        self.ce_weights = w

    def ce_score(self, ids, tt, lens):
        from oracle import cross_encoder as ce_oracle

        logits, sig = ce_oracle.numpy_forward(self.ce_weights, ids, tt, lens, dtype=np.float64)
        return logits.astype(np.float32), sig.astype(np.float32)


class OracleEngineTorch(OracleEngine):
    """Adds the `*_dev` methods on CPU torch tensors so HybridPipeline's sharded control flow (record packing, the single
    all-gather, shard merge, fusion) can run under gloo with world_size 2 on a box without GPUs."""

    def _fill(self, out, arrays):
        import torch

        for t, a in zip(out, arrays):
            t.copy_(torch.from_numpy(np.ascontiguousarray(a)).view(t.dtype).reshape(t.shape))
        return out

    def dense_topk_dev(self, q_t, k, slot=0, out=None):
        return self._fill(out, self.dense_topk(q_t.numpy(), k, slot))

    def bm25_topk_dev(self, terms_t, off_t, B, n_terms, max_len, k, out=None):
        terms, off = terms_t.numpy(), off_t.numpy()
        lists = [terms[off[b]:off[b + 1]] for b in range(B)]
        return self._fill(out, self.bm25_topk(lists, k))

    def fuse_dev(self, method, rrf_k, w_dense, w_sparse, k, dense, sparse, out=None):
        d = tuple(x.numpy() for x in dense)
        s = tuple(x.numpy() for x in sparse)
        return self._fill(out, self.fuse(method, rrf_k, w_dense, w_sparse, k, dense=d, sparse=s))

    def merge_shards_dev(self, ids0, scores0, counts0, shard_stride_bytes, G, out=None):
        import torch

        B, k = ids0.shape
        base = ids0.untyped_storage()
        raw = torch.tensor([], dtype=torch.uint8).set_(base)  # whole gathered buffer as bytes
        off_i, off_s, off_c = ids0.storage_offset() * 8, scores0.storage_offset() * 8, counts0.storage_offset() * 4
        ids = np.full((B, k), -1, np.int64)
        sc = np.zeros((B, k))
        cnt = np.zeros(B, np.int32)
        buf = raw.numpy()
        for b in range(B):
            cand = []
            for g in range(G):
                o = g * shard_stride_bytes
                gi = np.frombuffer(buf[o + off_i:o + off_i + B * k * 8].tobytes(), np.int64).reshape(B, k)
                gs = np.frombuffer(buf[o + off_s:o + off_s + B * k * 8].tobytes(), np.float64).reshape(B, k)
                gc = np.frombuffer(buf[o + off_c:o + off_c + B * 4].tobytes(), np.int32)
                cand += [(float(gs[b, j]), int(gi[b, j])) for j in range(int(gc[b]))]
            cand.sort(key=lambda t: (-t[0], t[1]))
            cand = cand[:k]
            ids[b, :len(cand)] = [c[1] for c in cand]
            sc[b, :len(cand)] = [c[0] for c in cand]
            cnt[b] = len(cand)
        return self._fill(out, (ids, sc, cnt))
