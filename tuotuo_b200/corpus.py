"""Device-resident doc-term matrix: CSR (document-major) plus its CSC mirror (word-major).

Replaces the dense ``Document.doc_vocab_count_array`` (tuotuo/document.py:33-36, built by tuotuo/utils.py:177-206):
an M x V float64 matrix cannot exist at the target sizes (400 GB at 1M docs x 50k words).  The E-step reads the CSR
rows — exactly ``np.nonzero(X[d,:])`` / ``X[d, ids]`` of tuotuo/lda_model.py:226-227 — and the M-step statistics
are accumulated from the CSC side.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nv


class DeviceCorpus:
    """rowptr int64[D+1], colidx int32[nnz] (ascending per row), counts int32[nnz] and the CSC mirror, on one device."""

    L2_WINDOW_BYTES = 50 << 20   # exp(E[log theta]) rows one document block may occupy: well under half of the 126 MB
                                 # L2 (two 63 MB partitions).  Measured at cfg3 with ticket-ordered chunks: statistics
                                 # pass 11.0 ms at 8 blocks (100 MB), 9.1 ms at 12-24 blocks (67-33 MB), DESIGN.md §5
    MAX_BLOCK_BYTES = 12 << 30   # cap on the per-block statistics buffer nblocks * V * ld * 8 (cfg4: 23 x 411 MB; its
                                 # write + fold costs ~3 ms and saves ~23 ms of HBM gathers)
    K_WINDOW = 128               # topic-window width of the windowed statistics pass (ttlda_sstats_csc_windowed_f64):
                                 # 1 KB per gathered row, the (G=8, E=8) shape of the traversal

    def __init__(self, rowptr, colidx, counts, V: int, *, build_csc: bool = False):
        assert rowptr.dtype == torch.int64 and colidx.dtype == torch.int32 and counts.dtype == torch.int32
        self.rowptr, self.colidx, self.counts = rowptr.contiguous(), colidx.contiguous(), counts.contiguous()
        self.device = rowptr.device
        self.D = int(rowptr.numel() - 1)
        self.V = int(V)
        self.nnz = int(colidx.numel())
        self.tokens = int(counts.sum(dtype=torch.int64).item()) if self.nnz else 0
        self.colptr = self.docidx = self.ccounts = self.csr2csc = None
        self.serial_blocks = False
        self.kwindow = 0  # > 0: the statistics pass walks the mirror once per window of this many topics
        self.e_idx = self.e_cnt = self.e_pos = self.hot_words = self.slot_of_word = None  # E-step stream, ensure_estream
        self.nblocks, self.block_docs = 0, 0
        if build_csc:
            self.build_csc()

    # -- constructors ----------------------------------------------------------------------------------------------
    @staticmethod
    def from_host_csr(rowptr: np.ndarray, colidx: np.ndarray, counts: np.ndarray, V: int, device, **kw):
        dev = torch.device(device)
        return DeviceCorpus(torch.as_tensor(np.ascontiguousarray(rowptr, dtype=np.int64)).to(dev),
                            torch.as_tensor(np.ascontiguousarray(colidx, dtype=np.int32)).to(dev),
                            torch.as_tensor(np.ascontiguousarray(counts, dtype=np.int32)).to(dev), V, **kw)

    @staticmethod
    def from_token_ids(token_ids: torch.Tensor, doc_offsets: torch.Tensor, V: int, **kw):
        """Token-id stream (int32, device) + document offsets (int64[D+1], device) -> CSR via ttlda_bow_to_csr."""
        dev = token_ids.device
        T, D = token_ids.numel(), doc_offsets.numel() - 1
        # ids outside [0, V) would corrupt the (doc, word) sort keys and every later gather: refuse them here
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        nv.validate_csr(doc_offsets.contiguous(), token_ids.contiguous() if T else None, None, V, flags, nnz=T)
        bad = int(flags.item())
        if bad:
            raise ValueError(nv.bad_corpus_message(bad).replace("rowptr", "doc_offsets").replace("nnz", "len(token_ids)"))
        ws = torch.empty(nv.bow_to_csr_workspace_bytes(T, D, V), dtype=torch.uint8, device=dev)
        rowptr = torch.empty(D + 1, dtype=torch.int64, device=dev)
        colidx = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
        counts = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
        nnz_out = torch.zeros(1, dtype=torch.int64, device=dev)
        nv.bow_to_csr(token_ids.contiguous(), doc_offsets.contiguous(), V, rowptr, colidx, counts, nnz_out, ws)
        nnz = int(nnz_out.item())
        del ws
        return DeviceCorpus(rowptr, colidx[:nnz].clone(), counts[:nnz].clone(), V, **kw)

    def private_copy(self) -> "DeviceCorpus":
        """A corpus with its own copies of every device array (same blocking, same mirror).  The streaming step
        (LDASmoothed.em_step_from_host) overwrites the CSR / mirror buffers it works in; it does so in such a copy, never
        in the instance `Document.device_corpus()` caches."""
        c = DeviceCorpus.__new__(DeviceCorpus)
        c.__dict__.update(self.__dict__)
        for name in ("rowptr", "colidx", "counts", "colptr", "docidx", "ccounts", "csr2csc", "e_idx", "e_cnt", "e_pos"):
            t = getattr(self, name, None)
            setattr(c, name, t.clone() if t is not None else None)
        if getattr(self, "block_pos", None) is not None:
            c.block_pos = list(self.block_pos)
        return c

    def plan_blocks(self, ld: int, windowed_ok: bool = False):
        """Document blocking of the CSC mirror for K-vector rows of `ld` doubles: blocks whose exp(E[log theta])
        rows fit an L2-sized window, if that needs few enough blocks to keep the per-block statistics small.

        windowed_ok (the statistics will run on the E-step's ratios): when whole-row blocks would need more per-block
        buffer than the cap allows, or rows are long (K > 256), the pass is planned TOPIC-WINDOWED instead:
        one walk of the mirror per K_WINDOW columns, so a block's L2 footprint is block_docs * K_WINDOW * 8 bytes and
        the per-block buffers are nblocks * V * K_WINDOW doubles (cfg5 shard: 25 blocks of 51k documents and 2.6 GB
        instead of 191 blocks and 153 GB).  TTLDA_KWINDOW=<width> forces it, TTLDA_KWINDOW=0 forbids it."""
        import os

        ov = os.environ.get("TTLDA_DOC_BLOCKS")
        kwo = os.environ.get("TTLDA_KWINDOW")
        self.kwindow = 0
        row = ld * 8
        nb = max(1, int(ov)) if ov else max(1, -(-self.D * row // self.L2_WINDOW_BYTES))
        too_big = nb > 1 and nb * self.V * row > self.MAX_BLOCK_BYTES
        long_rows = nb > 1 and ld > 2 * self.K_WINDOW  # measured: cfg4 (K=500) 21.0 -> 19.1 ms, 9.4 -> 0.6 GB of scratch
        if windowed_ok and kwo != "0" and (long_rows or (too_big and not ov) or (kwo and int(kwo) > 0)):
            kw = int(kwo) if kwo and int(kwo) > 0 else self.K_WINDOW
            if kw < ld:
                nbw = max(1, int(ov)) if ov else max(1, -(-self.D * kw * 8 // self.L2_WINDOW_BYTES))
                if nbw * self.V * kw * 8 <= self.MAX_BLOCK_BYTES and nbw * self.V < 2 ** 31:
                    self.kwindow = kw
                    self.serial_blocks = False
                    nbw = min(nbw, max(1, self.D))
                    block_docs = max(1, -(-self.D // nbw))
                    return max(1, -(-self.D // block_docs)), block_docs
        # Per-block partial statistics (one launch over all blocks in ticket order, folded afterwards) when they fit the
        # cap; else unblocked (HBM gather).  TTLDA_SERIAL_BLOCKS=1 selects the buffer-free alternative — one launch per
        # block accumulating into the final table in place (ttlda_sstats_csc_block_f64, accumulate) — which is exact
        # and tested but NOT the default for the shapes that would need it: at cfg5 (K = 1000: 191 blocks of 6.5k
        # documents) a (block, word) segment holds ~13 postings and the per-segment start-up / read-modify-write
        # epilogue outweighs the L2 residency (587 vs 250 ms for the pass).
        self.serial_blocks = os.environ.get("TTLDA_SERIAL_BLOCKS", "") == "1"
        if nb > 1 and not ov and not self.serial_blocks and nb * self.V * row > self.MAX_BLOCK_BYTES:
            nb = 1
        if self.serial_blocks and (nb * self.V >= 2 ** 31 or nb * self.V * 8 > self.MAX_BLOCK_BYTES):
            nb, self.serial_blocks = 1, False
        nb = min(nb, max(1, self.D))
        block_docs = max(1, -(-self.D // nb))
        nb = max(1, -(-self.D // block_docs))
        return nb, block_docs

    def build_csc(self, nblocks: int = 1, block_docs: int | None = None, with_map: bool = True):
        """(Re)build the CSC mirror on the device; document-blocked when nblocks > 1 (ttlda_csr_to_csc).
        with_map=False skips the csr->csc position map (only the E-step's ratio hand-off needs it)."""
        dev = self.device
        nblocks = max(1, int(nblocks))
        block_docs = int(block_docs) if block_docs else max(1, -(-self.D // nblocks))
        if self.colptr is None or self.colptr.numel() != nblocks * self.V + 1:
            self.colptr = torch.empty(nblocks * self.V + 1, dtype=torch.int64, device=dev)
        if self.docidx is None:
            self.docidx = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)
            self.ccounts = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)
            self.csr2csc = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)  # uint32 bit patterns
        # mirror position where every block starts (host copy: the per-block entry points take it as an argument)
        idx = torch.clamp(torch.arange(nblocks + 1, device=dev) * block_docs, max=self.D)
        self.block_pos = [int(x) for x in self.rowptr[idx].cpu().tolist()]
        if nblocks == 1:
            ws = torch.empty(nv.csr_to_csc_workspace_bytes(self.nnz, self.D, self.V, 1), dtype=torch.uint8, device=dev)
            nv.csr_to_csc(self.rowptr, self.colidx, self.counts, self.D, self.V, self.nnz, 1, block_docs,
                          self.colptr, self.docidx, self.ccounts, ws, self.csr2csc if with_map else None)
        else:
            # one block at a time: the same mirror (tested bit for bit) with block-sized sort buffers and block-local
            # random accesses (20 -> 10 ms at cfg3, 4 GB -> 0.3 GB of scratch)
            V = self.V
            big = max(self.block_pos[b + 1] - self.block_pos[b] for b in range(nblocks))
            ws = torch.empty(nv.csr_to_csc_workspace_bytes(big, block_docs, V, 1), dtype=torch.uint8, device=dev)
            for b in range(nblocks):
                d0, d1 = min(b * block_docs, self.D), min((b + 1) * block_docs, self.D)
                s0, s1 = self.block_pos[b], self.block_pos[b + 1]
                nv.csr_to_csc_block(self.rowptr[d0:d1 + 1], self.colidx, self.counts, V, d0, s0, s1 - s0,
                                    self.colptr[b * V:b * V + V + 1], self.docidx, self.ccounts, ws,
                                    self.csr2csc if with_map else None)
        self.nblocks, self.block_docs = nblocks, block_docs
        self.ccounts_valid = True
        self.e_idx = self.e_cnt = self.e_pos = self.hot_words = None  # derived from csr2csc: rebuilt on demand
        del ws

    def ensure_csc(self, ld: int, windowed_ok: bool = False):
        nb, bd = self.plan_blocks(ld, windowed_ok)
        if self.colptr is None or (self.nblocks, self.block_docs) != (nb, bd) or not getattr(self, "ccounts_valid", True):
            self.build_csc(nb, bd)

    def ensure_estream(self, ld: int, with_map: bool = True) -> bool:
        """The E-step stream of ttlda_estep_hot_f64 for rows of `ld` doubles: the H most frequent words (by number of
        documents containing them; H = what the shared-memory cache holds) get a cache slot, and within every document
        the nonzeros of cached words come first.  Derived data like the CSC mirror (12 bytes per nonzero); False when
        this row width has no cached form."""
        import os

        # opt-in (TTLDA_HOT=1): same-box A/B at cfg3 — 15.2 ms with the cache against 14.3 ms without
        H = min(nv.estep_hot_rows(ld), self.V) if os.environ.get("TTLDA_HOT", "0") == "1" else 0
        if ov := os.environ.get("TTLDA_HOT_ROWS"):
            H = min(H, max(0, int(ov)))
        if H <= 0 or self.nnz == 0:
            return False
        if self.e_idx is not None and self.hot_words.numel() == H and (self.e_pos is not None) == with_map:
            return True
        dev = self.device
        df = torch.bincount(self.colidx[:self.nnz], minlength=self.V)  # documents per word (rows hold distinct words)
        top = torch.topk(df, H, sorted=True).indices
        slot = torch.full((self.V,), -1, dtype=torch.int32, device=dev)
        slot[top] = torch.arange(H, dtype=torch.int32, device=dev)
        self.hot_words = top.to(torch.int32).contiguous()
        self.slot_of_word = slot
        self.hot_share = float(df[top].sum().item()) / max(self.nnz, 1)  # fraction of the nonzeros served from the cache
        self.e_idx = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)
        self.e_cnt = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)
        self.e_pos = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev) if with_map else None
        nv.build_estream(self.rowptr, self.colidx, self.counts, self.csr2csc if with_map else None, self.V, slot,
                         self.e_idx, self.e_cnt, self.e_pos)
        return True

    # -- host views ----------------------------------------------------------------------------------------------------
    def host_csr(self):
        return (self.rowptr.cpu().numpy(), self.colidx.cpu().numpy(), self.counts.cpu().numpy())

    def to_dense(self) -> np.ndarray:
        """M x V float64 (the reference's doc_vocab_count_array).  Small corpora only."""
        if self.D * self.V > 2 ** 31:
            raise MemoryError(f"dense {self.D} x {self.V} count matrix requested; use the CSR arrays instead")
        rp, ci, ct = self.host_csr()
        X = np.zeros((self.D, self.V), dtype=np.float64)
        rows = np.repeat(np.arange(self.D), np.diff(rp))
        X[rows, ci] = ct
        return X

    def nbytes(self) -> int:
        n = self.rowptr.numel() * 8 + self.nnz * 8
        if self.colptr is not None:
            n += self.colptr.numel() * 8 + self.nnz * 12
        return n


def balanced_doc_partition(rowptr: np.ndarray, world: int) -> np.ndarray:
    """Contiguous, nnz-balanced document blocks: returns int64[world+1] document boundaries.

    Multi-GPU sharding of SURVEY.md §8(e): documents are independent given lambda, so rank r owns documents
    [b[r], b[r+1]) and the matching gamma rows; only the K x V statistics cross NVLink."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    D = len(rowptr) - 1
    nnz = int(rowptr[-1])
    targets = (np.arange(1, world, dtype=np.float64) * nnz / world)
    cuts = np.searchsorted(rowptr, targets, side="left").astype(np.int64)
    b = np.concatenate([[0], np.clip(cuts, 0, D), [D]]).astype(np.int64)
    return np.maximum.accumulate(b)
