"""``Document`` — drop-in for tuotuo/document.py:9-50, backed by a device-resident CSR doc-term matrix.

Same constructor, same attributes (``sample_docs, num_doc_population, sample_docs_list, vocab_doc_count_dict,
vocab_doc_count_array, doc_vocab_count_array, vocab_to_idx, idx_to_vocab, doc_vocab_count_df, V, M``) and the
same three summary prints (tuotuo/utils.py:104-105,120).  Differences, all forced by scale:

* the vocabulary/token-id pass runs once on the host (ids = order of first appearance, tuotuo/utils.py:177-206,
  bit-exact); counting is done on the GPU by ``ttlda_bow_to_csr``;
* the dense matrices / DataFrame are materialised lazily, only when the attribute is read (small corpora);
* extra constructors for corpora that never exist as text: ``from_token_ids``, ``from_csr``, ``from_dense``;
* ``shard(rank, world)`` gives the contiguous document block a rank owns in multi-GPU runs.
"""
from __future__ import annotations

import numpy as np

from .text import text_pipeline, tokenize_corpus


class Document:
    def __init__(self, sample_docs: dict, num_doc_population: int = None) -> None:
        self.sample_docs = sample_docs
        if not num_doc_population:  # document.py:26-29
            self.num_doc_population = len(sample_docs)
        else:
            self.num_doc_population = num_doc_population

        # process_documents(sample=False) (utils.py:89-130) + get_vocab_from_docs / get_np_wct (utils.py:177-206): the
        # text pipeline, tokenisation and first-appearance vocabulary ids for the WHOLE corpus in one vectorised pass
        # (text.tokenize_corpus) instead of per-document Python loops
        ids, offs, words = tokenize_corpus(list(sample_docs.values()))
        length = int(offs[-1])
        print(f"There are {len(offs) - 1} documents in the dataset after processing")
        if len(sample_docs):
            print(f"On average estimated document length is {round(length / len(sample_docs), 1)} "
                  f"words per document after processing")
        vocab = {w: j for j, w in enumerate(words)}
        print(f"There are {len(vocab)} unique vocab in the corpus after processing")

        self._words = words
        self.vocab_to_idx = vocab
        self.idx_to_vocab = {v: k for k, v in vocab.items()}
        self.V = len(vocab)
        self.M = len(offs) - 1
        self._token_ids, self._doc_offsets = ids, offs
        self._host_csr = None
        self._dev = {}

    @property
    def sample_docs_list(self):  # document.py:31 — list of token lists, materialised on demand
        if getattr(self, "_docs_list", None) is None and getattr(self, "_words", None) is not None:
            w = np.asarray(self._words, dtype=object)
            ids, offs = np.asarray(self._token_ids), self._doc_offsets
            self._docs_list = [w[ids[offs[d]:offs[d + 1]]].tolist() for d in range(len(offs) - 1)]
        return getattr(self, "_docs_list", None)

    @sample_docs_list.setter
    def sample_docs_list(self, value):
        self._docs_list = value

    # -- alternative constructors ---------------------------------------------------------------------------------------
    @classmethod
    def _blank(cls, M, V, num_doc_population=None, idx_to_vocab=None):
        self = cls.__new__(cls)
        self.sample_docs = None
        self.sample_docs_list = None
        self.num_doc_population = num_doc_population if num_doc_population else M
        self.M, self.V = int(M), int(V)
        self.idx_to_vocab = idx_to_vocab if idx_to_vocab is not None else _LazyVocab(V)
        self.vocab_to_idx = ({v: k for k, v in idx_to_vocab.items()} if isinstance(idx_to_vocab, dict) else None)
        self._token_ids = self._doc_offsets = None
        self._host_csr = None
        self._dev = {}
        return self

    @classmethod
    def from_token_ids(cls, token_ids, doc_offsets, V, num_doc_population=None, idx_to_vocab=None):
        """token_ids: int32 vocabulary ids (NumPy or CUDA tensor); doc_offsets: int64[M+1]."""
        self = cls._blank(len(doc_offsets) - 1, V, num_doc_population, idx_to_vocab)
        self._token_ids, self._doc_offsets = token_ids, doc_offsets
        return self

    @classmethod
    def from_texts_with_vocab(cls, sample_docs: dict, vocab_to_idx: dict, num_doc_population=None):
        """Raw texts encoded against an EXISTING vocabulary (word -> id) instead of one built by first appearance:
        what online learning needs, where every minibatch must index the same lambda rows.  Same text pipeline as the
        constructor; words outside the vocabulary are dropped (returned count in `.oov_tokens`)."""
        ids, offs, oov = [], [0], 0
        for _, v in sample_docs.items():
            for w in text_pipeline(v).split(" "):
                j = vocab_to_idx.get(w)
                if j is None:
                    oov += 1
                else:
                    ids.append(j)
            offs.append(len(ids))
        self = cls.from_token_ids(np.asarray(ids, dtype=np.int32), np.asarray(offs, dtype=np.int64), len(vocab_to_idx),
                                  num_doc_population, {v: k for k, v in vocab_to_idx.items()})
        self.sample_docs = sample_docs
        self.oov_tokens = oov
        return self

    @classmethod
    def from_csr(cls, rowptr, colidx, counts, V, num_doc_population=None, idx_to_vocab=None):
        """Host CSR (NumPy): rowptr int64[M+1], colidx strictly ascending per row in [0, V), counts >= 0 — what
        np.nonzero(X[d, :]) / X[d, ids] give the reference (tuotuo/lda_model.py:226-227).  Checked here: these arrays
        drive device addressing, where a bad index is an out-of-bounds gather instead of NumPy's IndexError."""
        _validate_host_csr(np.asarray(rowptr), np.asarray(colidx), np.asarray(counts), int(V))
        self = cls._blank(len(rowptr) - 1, V, num_doc_population, idx_to_vocab)
        self._host_csr = (np.ascontiguousarray(rowptr, dtype=np.int64), np.ascontiguousarray(colidx, dtype=np.int32),
                          np.ascontiguousarray(counts, dtype=np.int32))
        return self

    @classmethod
    def from_dense(cls, X, num_doc_population=None, idx_to_vocab=None):
        """Dense M x V counts (the reference's doc_vocab_count_array) -> CSR.  Host-side format conversion."""
        X = np.asarray(X)
        rows, cols = np.nonzero(X)
        rowptr = np.zeros(X.shape[0] + 1, dtype=np.int64)
        np.add.at(rowptr, rows + 1, 1)
        vals = X[rows, cols]
        if not np.all(vals == np.round(vals)):
            raise ValueError("count matrix must hold integers")
        return cls.from_csr(np.cumsum(rowptr), cols, vals.astype(np.int32), X.shape[1], num_doc_population,
                            idx_to_vocab)

    @classmethod
    def from_uci(cls, docword_path, vocab_path=None, num_doc_population=None):
        """UCI "Bag of Words" format (docword.*.txt[.gz]: D, W, NNZ header lines, then `docID wordID count`,
        1-based) straight to CSR — the ingestion path for real corpora such as NYTimes; the
        reference can only ingest raw text through its O(V*M) Python loops.  vocab.*.txt gives idx_to_vocab."""
        import pandas as pd

        with (__import__("gzip").open(docword_path, "rt") if str(docword_path).endswith(".gz")
              else open(docword_path)) as f:
            D, W, NNZ = (int(f.readline()) for _ in range(3))
            t = pd.read_csv(f, sep=r"\s+", header=None, names=["d", "w", "c"], dtype=np.int64, engine="c")
        if len(t) != NNZ:
            raise ValueError(f"{docword_path}: header says {NNZ} entries, file has {len(t)}")
        d, w, c = t["d"].to_numpy() - 1, t["w"].to_numpy() - 1, t["c"].to_numpy()
        if len(t) and (d.min() < 0 or d.max() >= D or w.min() < 0 or w.max() >= W or c.min() < 1):
            raise ValueError(f"{docword_path}: index or count out of range")
        key = d * W + w
        order = np.argsort(key, kind="stable")
        key, c = key[order], c[order]
        uk, start = np.unique(key, return_index=True)  # duplicate (doc, word) lines are summed
        cs = np.add.reduceat(c, start) if len(c) else c
        rowptr = np.zeros(D + 1, dtype=np.int64)
        np.add.at(rowptr, uk // W + 1, 1)
        vocab = None
        if vocab_path is not None:
            with open(vocab_path) as f:
                vocab = {i: line.rstrip("\n") for i, line in enumerate(f)}
            if len(vocab) != W:
                raise ValueError(f"{vocab_path}: {len(vocab)} words, docword header says {W}")
        return cls.from_csr(np.cumsum(rowptr), (uk % W).astype(np.int32), cs.astype(np.int32), W, num_doc_population,
                            vocab)

    # -- device side ---------------------------------------------------------------------------------------------------------
    def device_corpus(self, device):
        """The CSR/CSC doc-term matrix on `device` (built once per device)."""
        import torch

        from .corpus import DeviceCorpus

        dev = torch.device(device)
        key = str(dev)
        if key not in self._dev:
            if self._host_csr is not None:
                self._dev[key] = DeviceCorpus.from_host_csr(*self._host_csr, self.V, dev)
            else:
                tok, offs = self._token_ids, self._doc_offsets
                tok = tok if isinstance(tok, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(tok, dtype=np.int32))
                offs = offs if isinstance(offs, torch.Tensor) else torch.as_tensor(
                    np.ascontiguousarray(offs, dtype=np.int64))
                self._dev[key] = DeviceCorpus.from_token_ids(tok.to(dev), offs.to(dev), self.V)
        return self._dev[key]

    def _default_corpus(self):
        if self._dev:
            return next(iter(self._dev.values()))
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("Document count matrices are built on the GPU (ttlda_bow_to_csr); no CUDA device found")
        return self.device_corpus(torch.device("cuda", torch.cuda.current_device()))

    def host_csr(self):
        if self._host_csr is None:
            self._host_csr = self._default_corpus().host_csr()
        return self._host_csr

    def shard(self, rank: int, world: int) -> "Document":
        """Contiguous document block of `rank` out of `world` (token/nnz balanced)."""
        from .corpus import balanced_doc_partition

        if self._host_csr is not None:
            rp, ci, ct = self._host_csr
            b = balanced_doc_partition(rp, world)
            lo, hi = int(b[rank]), int(b[rank + 1])
            s, e = int(rp[lo]), int(rp[hi])
            sub = Document.from_csr(rp[lo:hi + 1] - s, ci[s:e], ct[s:e], self.V, self.num_doc_population,
                                    self.idx_to_vocab if isinstance(self.idx_to_vocab, dict) else None)
        else:
            import torch

            offs = self._doc_offsets
            offs_np = offs.cpu().numpy() if isinstance(offs, torch.Tensor) else np.asarray(offs)
            b = balanced_doc_partition(offs_np, world)
            lo, hi = int(b[rank]), int(b[rank + 1])
            s, e = int(offs_np[lo]), int(offs_np[hi])
            sub = Document.from_token_ids(self._token_ids[s:e], offs[lo:hi + 1] - s, self.V, self.num_doc_population,
                                          self.idx_to_vocab if isinstance(self.idx_to_vocab, dict) else None)
        sub.doc_range = (lo, hi)
        return sub

    # -- the reference's dense attributes, materialised on demand ---------------------------------------------------------------
    @property
    def doc_vocab_count_array(self) -> np.ndarray:  # document.py:36 (M x V, a transposed view there too)
        return self.vocab_doc_count_array.T

    @property
    def vocab_doc_count_array(self) -> np.ndarray:  # document.py:35 (V x M float64)
        if getattr(self, "_vd", None) is None:
            rp, ci, ct = self.host_csr()
            if self.M * self.V > 2 ** 31:
                raise MemoryError("dense count matrix requested for a large corpus; use host_csr()")
            vd = np.zeros((self.V, self.M), dtype=float)
            vd[ci, np.repeat(np.arange(self.M), np.diff(rp))] = ct
            self._vd = vd
        return self._vd

    @property
    def vocab_doc_count_dict(self) -> dict:  # document.py:34: word -> list of per-document counts
        vd = self.vocab_doc_count_array
        return {self.idx_to_vocab[i]: [int(c) for c in vd[i]] for i in range(self.V)}

    @property
    def doc_vocab_count_df(self):  # document.py:41-44
        import pandas as pd

        return pd.DataFrame(data=self.doc_vocab_count_array, columns=[self.idx_to_vocab[i] for i in range(self.V)])


def _validate_host_csr(rp, ci, ct, V):
    if rp.ndim != 1 or ci.ndim != 1 or ct.ndim != 1 or len(rp) < 1:
        raise ValueError("malformed corpus: rowptr, colidx and counts must be 1-D, rowptr non-empty")
    if len(ci) != len(ct):
        raise ValueError("malformed corpus: colidx and counts must have one entry per nonzero")
    if rp[0] != 0 or rp[-1] != len(ci) or (len(rp) > 1 and bool((rp[1:] < rp[:-1]).any())):
        raise ValueError("malformed corpus: rowptr must start at 0, be non-decreasing and end at nnz")
    if len(ci):
        if ci.min() < 0 or ci.max() >= V:
            raise ValueError("malformed corpus: word ids must lie in [0, V)")
        if ct.min() < 0:
            raise ValueError("malformed corpus: counts must be >= 0")
        if np.asarray(ct).dtype.kind == "f" and not np.all(ct == np.round(ct)):
            raise ValueError("malformed corpus: counts must be integers")
        if len(ci) > 1:
            desc = ci[1:] <= ci[:-1]
            if desc.any():  # allowed only where a new row starts
                starts = np.zeros(len(ci) + 1, dtype=bool)
                starts[np.asarray(rp[1:-1], dtype=np.int64)] = True
                if bool((desc & ~starts[1:-1]).any()):
                    raise ValueError("malformed corpus: word ids must be strictly ascending within a document")


class _LazyVocab:
    """idx -> 'w<idx>' for synthetic corpora without strings (dict-like, no V-sized allocation)."""

    def __init__(self, V):
        self.V = int(V)

    def __getitem__(self, i):
        i = int(i)
        if not 0 <= i < self.V:
            raise KeyError(i)
        return f"w{i}"

    def __len__(self):
        return self.V

    def items(self):
        return ((i, f"w{i}") for i in range(self.V))
