"""``LDASmoothed`` — drop-in for tuotuo/lda_model.py:49-564 whose batch variational-inference hot path runs as
hand-written sm_100a kernels (``libttlda.so``, C ABI in ``include/ttlda.h``).

Same constructor, same ``fit`` signature and control flow (initial perplexity, early ``return`` on a perplexity
delta below ``threshold``, hyper-parameter Newton updates followed by one perplexity with the stale expectations),
same prints, same attributes (``_lambda_`` (K,V), ``_gamma_`` (M,K), ``_alpha_``, ``_eta_``, ``K``, ``M``, ``V``,
``train_doc``).  State lives on the GPU in word-major tables; the NumPy attributes are snapshots copied back on
access.  There is NO CPU path: without a CUDA device / the built library every call raises.

Per EM iteration (tuotuo/lda_model.py:285-316 + :403-409):
    ttlda_estep_f64          gamma (e_step :197-270) + the ELBO token term of the INCOMING state (:125-139): it is the
                             same dot product the normaliser needs, so the perplexity of iteration t is a by-product
                             of E-step t+1; leaves the ratio c/N of every nonzero in mirror order for the statistics
    [48-byte D2H]            perplexity of the incoming state -> the reference's convergence test (:386-390); the
                             E-step wrote into ping-pong buffers, so on "converged" it is simply discarded
    ttlda_sstats_csc_f64     K x V sufficient statistics, word-major, no atomics        (:262)
    [NCCL all_reduce]        documents are sharded across ranks; sum the statistics     
    ttlda_theta_refresh_f64  exp(E[log theta]) of the new gamma + theta-prior term      (:256, :142-148)
    ttlda_mstep_f64          lambda, exp(E[log beta]) refresh, beta-prior term          (:264, m_step :273-283, :157-163)
Beyond the reference: em_step_from_host (one iteration on a corpus streamed from pinned host memory, pipelined per
document block) and partial_fit (online variational Bayes).
After the last iteration one ttlda_elbo_tokens_f64 pass yields the final perplexity.
"""
from __future__ import annotations

import os
import time
import warnings

import numpy as np
import torch
from scipy.special import polygamma, psi

from . import _native as nv
from .corpus import DeviceCorpus
from .document import Document

SEED = 42  # lda_model.py:19
DTYPE = float  # lda_model.py:20


def np_clip_for_exp(x, sub: float = 1e-05):
    """tuotuo/utils.py:20-26."""
    if x > 0:
        x = max(sub, x)
    elif x < 0:
        x = min(sub, x)
    return x


class _State:
    """Device tensors of one fit."""
    pass


class LDASmoothed:
    """Smoothed LDA, batch variational inference (drop-in for tuotuo.lda_model.LDASmoothed)."""

    def __init__(self, num_topics: int, exchangeable_prior: bool = True, verbose: bool = True, *,
                 device=None, sstats_mode: str = "csc", inner_fixed_point: bool = False) -> None:
        self.K = num_topics
        np.random.seed(SEED)  # lda_model.py:71 (global side effect, kept)
        if exchangeable_prior:
            self._alpha_ = 1
        else:
            self._alpha_ = np.random.gamma(shape=100, scale=0.01, size=(1, self.K))
        if verbose:
            print(f"Topic Dirichlet Prior, α")
            print(self._alpha_)
            print()
        self._eta_ = 1
        if verbose:
            print(f"Exchangeable Word Dirichlet Prior, Eta ")
            print(self._eta_)
            print()
        if sstats_mode not in ("csc", "atomic"):
            raise ValueError("sstats_mode must be 'csc' (two-pass, reproducible) or 'atomic' (one pass, red.f64)")
        self._sstats_mode = sstats_mode
        # opt-in: refresh exp(E[log theta_d]) inside the per-document loop (true fixed point) instead of
        # reproducing the reference's frozen-theta loop; NOT parity with the reference by construction
        self._inner_fixed_point = bool(inner_fixed_point)
        self.inner_iterations = 0  # inner passes of the last E-step, summed over documents
        self._fused_refresh = os.environ.get("TTLDA_FUSED_REFRESH", "0") == "1"
        self._ratio_handoff = os.environ.get("TTLDA_RATIO_HANDOFF", "1") == "1"
        self._device = device
        self._st = None
        self._np_cache = {}

    # ------------------------------------------------------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------------------------------------------------------
    def _dev(self) -> torch.device:
        if self._device is None and not torch.cuda.is_available():
            raise nv.TTLDAError("LDASmoothed needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(self._device) if self._device is not None else torch.device("cuda")
        if dev.type == "cuda" and dev.index is None:
            # pin "cuda" to the device that is current NOW: every later call launches there (_native._call)
            dev = torch.device("cuda", torch.cuda.current_device())
        self._device = dev  # a non-CUDA device is refused by nv.check_device in _alloc_state (no CPU path)
        return dev

    @staticmethod
    def _dist():
        import torch.distributed as dist

        return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None

    def _as_document(self, train_doc) -> Document:
        if isinstance(train_doc, Document):
            return train_doc
        if isinstance(train_doc, dict):  # readme.md:42-47 passes the raw dict
            return Document(train_doc)
        if isinstance(train_doc, np.ndarray):
            return Document.from_dense(train_doc)
        if hasattr(train_doc, "doc_vocab_count_array"):  # the reference's own Document, or a duck type
            return Document.from_dense(np.asarray(train_doc.doc_vocab_count_array),
                                       getattr(train_doc, "num_doc_population", None),
                                       getattr(train_doc, "idx_to_vocab", None))
        raise TypeError(f"cannot build a corpus from {type(train_doc)}")

    def _alpha_vec_host(self) -> np.ndarray:
        return np.ascontiguousarray(
            np.broadcast_to(np.asarray(self._alpha_, dtype=np.float64).reshape(-1), (self.K,)))

    def _alpha_sum(self) -> float:
        # np.sum(real_dist_prior) at lda_model.py:43 — a scalar prior sums to itself
        return float(np.sum(self._alpha_))

    def _upload_alpha(self):
        st = self._st
        st.alpha_vec = torch.as_tensor(self._alpha_vec_host().copy()).to(st.dev)

    def _alloc_topic_tables(self, st, sharded_ok: bool):
        """lambda / exp(E[log beta]) / statistics tables [V][ld].  On one GPU: plain tensors.  Document-sharded over N
        ranks (NCCL): the M-step is SHARDED over the vocabulary as well — rank r owns rows [r*Vs, (r+1)*Vs) of lambda —
        so the tables get V_pad = N*Vs rows, and (mode "peer", the default when the ranks share an NVLink domain) the
        statistics and exp(E[log beta]) tables are allocated as symmetric memory: every rank maps every rank's table,
        the reduce-scatter of the statistics becomes plain NVLink loads inside ttlda_mstep_slice_lambda_f64 and the
        all-gather of exp(E[log beta]) plain NVLink stores inside ttlda_mstep_slice_beta_f64.  Mode "nccl" does the two
        exchanges with library collectives around the same kernels; "replicated" is round 1's all-reduce + a full
        M-step on every rank.  TTLDA_MSTEP selects; a failed symmetric-memory rendezvous degrades to "nccl"."""
        dev, V, ld = st.dev, st.V, st.ld
        f64 = torch.float64
        dist = self._dist()
        mode = "local"
        if dist is not None:
            mode = os.environ.get("TTLDA_MSTEP", "peer")
            if mode not in ("peer", "nccl", "replicated"):
                raise ValueError("TTLDA_MSTEP must be peer, nccl or replicated")
            if not sharded_ok or dist.get_backend() != "nccl" or dev.type != "cuda" or dist.get_world_size() > 8:
                mode = "replicated"
        st.mstep_mode, st.lam_synced, st.beta_sharded = mode, True, False
        if mode in ("local", "replicated"):
            st.lam = torch.zeros((V, ld), dtype=f64, device=dev)
            st.eB = torch.zeros((V, ld), dtype=f64, device=dev)
            st.sstats = torch.zeros((V, ld), dtype=f64, device=dev)
            return
        N, r = dist.get_world_size(), dist.get_rank()
        Vs = -(-V // N)
        st.Vs, st.w0 = Vs, r * Vs
        st.wn = max(0, min(Vs, V - st.w0))
        Vp = N * Vs
        st.symm = None
        if mode == "peer":
            try:
                import torch.distributed._symmetric_memory as symm

                group = dist.group.WORLD
                if hasattr(symm, "enable_symm_mem_for_group") and not symm.is_symm_mem_enabled_for_group(group.group_name):
                    symm.enable_symm_mem_for_group(group.group_name)
                st.sstats_pad = symm.empty((Vp, ld), dtype=f64, device=dev)
                st.eB_pad = symm.empty((Vp, ld), dtype=f64, device=dev)
                hs, hb = symm.rendezvous(st.sstats_pad, group), symm.rendezvous(st.eB_pad, group)
                st.S_ptrs, st.eB_ptrs = [int(p) for p in hs.buffer_ptrs], [int(p) for p in hb.buffer_ptrs]
                # NVSwitch multicast objects over the same tables (NVLink SHARP): in-switch sum / replicated store
                use_mc = os.environ.get("TTLDA_MULTICAST", "1") == "1" and hs.has_multicast_support and hb.has_multicast_support
                st.S_mc = int(hs.multicast_ptr) if use_mc else 0
                st.eB_mc = int(hb.multicast_ptr) if use_mc else 0
                if not (st.S_mc and st.eB_mc):
                    st.S_mc = st.eB_mc = 0
                if len(st.S_ptrs) != N or len(st.eB_ptrs) != N:
                    raise RuntimeError("symmetric memory rendezvous returned an unexpected number of peers")
                st.symm = hs
                st.sstats_pad.zero_()
                st.eB_pad.zero_()
                ok = 1
            except Exception as e:  # no peer mapping on this system (no NVLink / fabric handles refused)
                st.peer_error = f"{type(e).__name__}: {e}"
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank takes the same path
            if int(flag.item()) == 0:
                mode = st.mstep_mode = "nccl"
                st.symm = None
        if mode == "nccl":
            st.sstats_pad = torch.zeros((Vp, ld), dtype=f64, device=dev)
            st.eB_pad = torch.zeros((Vp, ld), dtype=f64, device=dev)
            st.S_slice = torch.zeros((Vs, ld), dtype=f64, device=dev)
        st.lam_pad = torch.zeros((Vp, ld), dtype=f64, device=dev)
        st.lam, st.eB, st.sstats = st.lam_pad[:V], st.eB_pad[:V], st.sstats_pad[:V]
        st.colpart_mine = torch.zeros(ld, dtype=f64, device=dev)
        st.colparts = torch.zeros((N, ld), dtype=f64, device=dev)

    def _sync_lambda(self):
        """After a sharded M-step every rank holds only its own rows of lambda: gather them before any full-table use."""
        st = self._st
        if st is not None and not getattr(st, "lam_synced", True):
            self._dist().all_gather_into_tensor(st.lam_pad, st.lam_pad[st.w0:st.w0 + st.Vs])
            st.lam_synced = True

    def _alloc_state(self, corpus: DeviceCorpus, topics_from=None):
        """Device tensors for `corpus`.  topics_from: a previous state whose topic-side tables (lambda, exp(E[log beta]),
        column sums, beta-prior record) are carried over (online learning: the documents change, the topics persist)."""
        dev = corpus.device
        nv.check_device(dev)
        K, V, D = int(self.K), corpus.V, corpus.D
        if not 1 <= K <= nv.MAX_K:
            raise ValueError(f"num_topics must be in [1, {nv.MAX_K}]")
        ld = nv.ld_for(K)
        st = _State()
        st.dev, st.corpus, st.K, st.V, st.D, st.ld = dev, corpus, K, V, D, ld
        f64 = torch.float64
        st.gamma = [torch.zeros((D, ld), dtype=f64, device=dev), torch.zeros((D, ld), dtype=f64, device=dev)]
        st.eT = [torch.zeros((D, ld), dtype=f64, device=dev), torch.zeros((D, ld), dtype=f64, device=dev)]
        st.cur = 0
        self._alloc_topic_tables(st, sharded_ok=topics_from is None)
        st.colsum = torch.zeros(ld, dtype=f64, device=dev)
        st.part = torch.zeros((5, nv.NPART), dtype=f64, device=dev)
        st.iters = torch.zeros(2, dtype=torch.int64, device=dev)
        st.flags = torch.zeros(1, dtype=torch.int32, device=dev)  # TTLDA_BAD_* bits of a streamed corpus
        st.ws = torch.empty(nv.reduce_workspace_bytes(), dtype=torch.uint8, device=dev)
        st.ws_m = torch.empty(nv.mstep_workspace_bytes(V, ld), dtype=torch.uint8, device=dev)
        st.ws_s = st.ratio = None
        if self._sstats_mode == "csc":
            handoff = self._ratio_handoff and not self._inner_fixed_point
            corpus.ensure_csc(ld, windowed_ok=handoff)  # (document-blocked) CSC mirror sized for this K
            st.ws_s = torch.empty(self._sstats_ws_bytes(corpus, corpus.nnz, V, ld), dtype=torch.uint8, device=dev)
            if handoff:
                # c/N per nonzero, written by the E-step in mirror order, consumed by the statistics pass
                st.ratio = torch.empty(max(corpus.nnz, 1), dtype=f64, device=dev)
        # E-step stream with the hot words first (shared-memory row cache); needs the mirror's position map when the
        # ratios are handed over, so it is built after it
        st.hot = (not self._inner_fixed_point and not self._fused_refresh
                  and (st.ratio is not None or self._sstats_mode != "csc" or not self._ratio_handoff)
                  and corpus.ensure_estream(ld, with_map=st.ratio is not None))
        if topics_from is not None:
            if (topics_from.K, topics_from.V, topics_from.dev) != (K, V, dev):
                raise ValueError("the minibatch must have the vocabulary size / device of the model being updated")
            st.lam, st.eB, st.colsum, st.ws_m = topics_from.lam, topics_from.eB, topics_from.colsum, topics_from.ws_m
            st.part[2].copy_(topics_from.part[2])
        st.terms = dict(tok=0.0, theta=0.0, beta=0.0, count=0.0)
        self._st = st
        self._np_cache = {}
        self._upload_alpha()
        return st

    # ------------------------------------------------------------------------------------------------------------------
    # NumPy views of the device state (snapshots)
    # ------------------------------------------------------------------------------------------------------------------
    @property
    def _lambda_(self) -> np.ndarray:
        if self._st is None:
            raise AttributeError("'LDASmoothed' object has no attribute '_lambda_'")
        if "lam" not in self._np_cache:
            st = self._st
            self._sync_lambda()  # collective when the M-step is sharded: call it on every rank
            out = torch.empty((st.K, st.V), dtype=torch.float64, device=st.dev)
            nv.transpose_unpad(st.lam, st.K, out)
            self._np_cache["lam"] = out.cpu().numpy()
        return self._np_cache["lam"]

    @_lambda_.setter
    def _lambda_(self, value):
        st = self._st
        if st is None:
            raise AttributeError("fit() first")
        v = torch.as_tensor(np.ascontiguousarray(value, dtype=np.float64)).to(st.dev)
        assert v.shape == (st.K, st.V)
        nv.transpose_pad(v, st.lam)
        st.lam_synced = True
        self._np_cache.pop("lam", None)

    @property
    def _gamma_(self) -> np.ndarray:
        if self._st is None:
            raise AttributeError("'LDASmoothed' object has no attribute '_gamma_'")
        if "gamma" not in self._np_cache:
            st = self._st
            local = st.gamma[st.cur][:, :st.K].contiguous()
            dist = self._dist()
            if dist is not None:  # gather the document shards in rank order
                sizes = [torch.zeros(1, dtype=torch.int64, device=st.dev) for _ in range(dist.get_world_size())]
                dist.all_gather(sizes, torch.tensor([st.D], dtype=torch.int64, device=st.dev))
                mx = int(max(int(s.item()) for s in sizes))
                pad = torch.zeros((mx, st.K), dtype=torch.float64, device=st.dev)
                pad[:st.D] = local
                parts = [torch.empty_like(pad) for _ in sizes]
                dist.all_gather(parts, pad)
                local = torch.cat([p[:int(s.item())] for p, s in zip(parts, sizes)], 0)
            self._np_cache["gamma"] = local.cpu().numpy()
        return self._np_cache["gamma"]

    @_gamma_.setter
    def _gamma_(self, value):
        st = self._st
        if st is None:
            raise AttributeError("fit() first")
        v = torch.as_tensor(np.ascontiguousarray(value, dtype=np.float64)).to(st.dev)
        assert v.shape == (st.D, st.K)
        st.gamma[st.cur].zero_()
        st.gamma[st.cur][:, :st.K] = v
        self._np_cache.pop("gamma", None)

    def _invalidate(self, *names):
        for n in names or ("lam", "gamma"):
            self._np_cache.pop(n, None)

    # ------------------------------------------------------------------------------------------------------------------
    # device-level steps
    # ------------------------------------------------------------------------------------------------------------------
    def _refresh_theta(self):
        st = self._st
        nv.theta_refresh(st.gamma[st.cur], st.K, st.alpha_vec, self._alpha_sum(), st.eT[st.cur], st.part[1], st.ws)

    def _refresh_theta_term_only(self):
        st = self._st
        nv.theta_refresh(st.gamma[st.cur], st.K, st.alpha_vec, self._alpha_sum(), None, st.part[1], st.ws)

    def _refresh_beta(self):
        st = self._st
        self._sync_lambda()
        nv.beta_refresh(st.lam, st.K, float(self._eta_), st.eB, None, st.colsum, st.part[2], st.ws_m)
        st.beta_sharded = False  # every rank now holds the complete beta-prior term

    def _tokens(self):
        st, c = self._st, self._st.corpus
        st.ev_corpus_free = None  # this pass reads the CSR arrays: a later streamed step must wait for it
        nv.elbo_tokens(c.rowptr, c.colidx, c.counts, st.K, st.eT[st.cur], st.eB, st.part[3], st.ws)

    def _estep_speculative(self, e_num_step: int, e_threshold: float = 1e-05):
        """E-step (lda_model.py:197-270) from the current state into the ping-pong buffers.  By-product: the ELBO
        token term of the CURRENT state (part[0][0]), which completes its perplexity."""
        st, c = self._st, self._st.corpus
        cur, nxt = st.cur, st.cur ^ 1
        st.ev_corpus_free = None  # reads the CSR arrays: a later streamed step must wait for it
        if self._inner_fixed_point:
            nv.estep_fixed_point(c.rowptr, c.colidx, c.counts, st.K, st.V, st.alpha_vec, self._alpha_sum(),
                                 st.eT[cur], st.eB, st.gamma[cur], st.gamma[nxt], st.eT[nxt], e_num_step,
                                 e_threshold, st.iters, st.part[0], st.ws)
        elif self._fused_refresh:
            nv.estep(c.rowptr, c.colidx, c.counts, st.K, st.V, st.alpha_vec, self._alpha_sum(), st.eT[cur], st.eB,
                     st.gamma[cur], st.gamma[nxt], st.eT[nxt], e_num_step, e_threshold, st.iters, st.part[0], st.ws,
                     c.csr2csc if st.ratio is not None else None, st.ratio)
        elif getattr(st, "hot", False):
            # gamma only, with the most frequent words' exp(E[log beta]) rows cached in shared memory
            nv.estep_hot(c.rowptr, c.e_idx, c.e_cnt, c.e_pos if st.ratio is not None else None, c.hot_words, st.K, st.V,
                         st.alpha_vec, st.eT[cur], st.eB, st.gamma[cur], st.gamma[nxt], st.ratio, e_num_step,
                         e_threshold, st.iters, st.part[0], st.ws)
        else:
            # gamma only; the theta refresh (digamma / log-gamma / exp per gamma_dk) runs as its own streaming pass
            nv.estep(c.rowptr, c.colidx, c.counts, st.K, st.V, st.alpha_vec, self._alpha_sum(), st.eT[cur], st.eB,
                     st.gamma[cur], st.gamma[nxt], None, e_num_step, e_threshold, st.iters, st.part[0], st.ws,
                     c.csr2csc if st.ratio is not None else None, st.ratio)

    @staticmethod
    def _sstats_ws_bytes(c, nnz, V, ld) -> int:
        if c.kwindow:
            return max(nv.sstats_windowed_workspace_bytes(nnz, V, c.kwindow, c.nblocks),
                       nv.sstats_workspace_bytes(nnz, V, ld, 1))  # the recomputing form runs unblocked buffers
        return nv.sstats_workspace_bytes(nnz, V, ld, 1 if c.serial_blocks else c.nblocks)

    def _sstats_pass(self, eT, ratio):
        """The word-major statistics pass over the (blocked) mirror into st.sstats."""
        st, c = self._st, self._st.corpus
        if c.kwindow:
            if ratio is None:
                raise nv.TTLDAError("the topic-windowed statistics pass runs on the E-step's ratios")
            nv.sstats_csc_windowed(c.colptr, c.docidx, st.V, st.K, c.nnz, c.nblocks, c.kwindow, eT, ratio, st.sstats,
                                   st.ws_s)
        elif c.serial_blocks and c.nblocks > 1:
            # one launch per document block, accumulating in place: the launches serialise the blocks, so every block's
            # theta window is L2-resident without nblocks * V * ld per-block buffers; same bits as the folded form
            V = st.V
            for b in range(c.nblocks):
                s, e = c.block_pos[b], c.block_pos[b + 1]
                nv.sstats_csc_block(c.colptr[b * V:b * V + V + 1], c.docidx, c.ccounts, V, st.K, s, e - s, eT, st.eB,
                                    st.sstats, st.ws_s, ratio, accumulate=b > 0)
        else:
            nv.sstats_csc(c.colptr, c.docidx, c.ccounts, st.V, st.K, c.nnz, c.nblocks, eT, st.eB, st.sstats, st.ws_s, ratio)

    def _launch_statistics(self, use_ratio: bool = True, have_sstats: bool = False):
        """Statistics pass (lda_model.py:262) with the state the speculative E-step read, [NCCL sum over the document
        shards], and the theta refresh of the E-step's gamma.  Writes st.sstats, the ping-pong exp(E[log theta]) buffer
        and a scratch record only — nothing of the CURRENT state — so it can be queued before the host knows whether the
        iteration will be kept (the perplexity record is still in flight, `_collect_begin`).  Returns the pending
        collective (or None).  have_sstats: st.sstats already holds this rank's statistics (the streaming path)."""
        st, c = self._st, self._st.corpus
        # the statistics use the theta the E-step read — or, in fixed-point mode, the refreshed one it produced
        eT = st.eT[st.cur ^ 1] if self._inner_fixed_point else st.eT[st.cur]
        if have_sstats:
            pass
        elif self._sstats_mode == "atomic":
            nv.sstats_atomic(c.rowptr, c.colidx, c.counts, st.K, st.V, eT, st.eB, st.sstats)
        else:
            if not (use_ratio and st.ratio is not None) and not getattr(c, "ccounts_valid", True):
                c.build_csc(c.nblocks, c.block_docs)  # a streamed step left the mirrored counts stale
            self._sstats_pass(eT, st.ratio if use_ratio else None)
        dist = self._dist()
        work = None
        if dist is not None:
            # sum of the K x V statistics over the document shards; the theta refresh below overlaps it
            if st.mstep_mode == "replicated":
                work = dist.all_reduce(st.sstats, async_op=True)
            elif st.mstep_mode == "nccl":  # every rank receives the sum of ITS rows only
                work = dist.reduce_scatter_tensor(st.S_slice, st.sstats_pad, async_op=True)
            # "peer": nothing to launch — the M-step's first kernel reads the peers' tables itself
            if nv.timing is not None:  # bench.py: from here to the M-step's first kernel
                st.ev_ar = torch.cuda.Event(enable_timing=True)
                st.ev_ar.record()
        if not (self._inner_fixed_point or self._fused_refresh):
            nxt = st.cur ^ 1
            nv.theta_refresh(st.gamma[nxt], st.K, st.alpha_vec, self._alpha_sum(), st.eT[nxt], st.part[4], st.ws)
        return work

    def _finish_iteration(self, work=None, online=None):
        """M-step (lda_model.py:264, m_step :273-283) on the summed statistics, then adopt the E-step's gamma /
        exp(E[log theta]) and its theta-prior ELBO term."""
        st = self._st
        if self._inner_fixed_point or self._fused_refresh:
            st.part[1, 0].copy_(st.part[0, 1])  # the E-step kernel produced gamma's theta-prior term itself
        else:
            st.part[1, 0].copy_(st.part[4, 0])
        if work is not None:
            work.wait()
        mode = getattr(st, "mstep_mode", "local")
        if mode == "peer":
            st.symm.barrier(channel=0)  # every rank's statistics pass has finished: its table may be read over NVLink
        if nv.timing is not None and getattr(st, "ev_ar", None) is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            nv.timing.setdefault("exchange_to_mstep", []).append((st.ev_ar, e1))
            st.ev_ar = None
        if mode in ("peer", "nccl"):
            if online is not None:
                raise nv.TTLDAError("the online M-step runs replicated")
            dist = self._dist()
            eta = float(self._eta_)
            src = st.S_ptrs if mode == "peer" else [st.S_slice.data_ptr()]
            nv.mstep_slice_lambda(src, st.w0 if mode == "peer" else 0, st.w0, st.wn, st.K, eta, st.eB, st.lam,
                                  st.colpart_mine, st.ws_m, multicast=st.S_mc if mode == "peer" else 0)
            # K column sums per rank: a 8K-byte exchange that is also the point after which no rank reads the others'
            # statistics any more (the next statistics pass may overwrite them)
            dist.all_gather_into_tensor(st.colparts, st.colpart_mine)
            dst = st.eB_ptrs if mode == "peer" else [st.eB.data_ptr()]
            nv.mstep_slice_beta(st.colparts, st.w0, st.wn, st.K, eta, st.lam, dst, dist.get_rank() == 0, st.colsum,
                                st.part[2], st.ws_m, multicast=st.eB_mc if mode == "peer" else 0)
            if mode == "peer":
                st.symm.barrier(channel=1)  # all ranks' rows of exp(E[log beta]) have landed in this rank's table
            else:
                dist.all_gather_into_tensor(st.eB_pad, st.eB_pad[st.w0:st.w0 + st.Vs])
            st.lam_synced, st.beta_sharded = False, True
        elif online is None:
            nv.mstep(st.sstats, st.K, float(self._eta_), st.eB, st.lam, None, st.colsum, st.part[2], st.ws_m)
        else:  # (scale, rho): stochastic natural-gradient step on lambda (partial_fit)
            nv.mstep_online(st.sstats, st.K, float(self._eta_), online[0], online[1], st.eB, st.lam, None, st.colsum,
                            st.part[2], st.ws_m)
        st.cur ^= 1
        self._invalidate()

    def _commit_iteration(self, use_ratio: bool = True, have_sstats: bool = False, online=None):
        self._finish_iteration(self._launch_statistics(use_ratio, have_sstats), online)

    def em_iteration(self, perplexity_prev=None, threshold: float = 1e-05, sampling: bool = False,
                     e_num_step: int = 100):
        """ONE pass of `fit`'s loop body (lda_model.py:384-409) on the fitted corpus: E-step from the current state (which
        also completes the perplexity OF the current state), the reference's convergence test against
        `perplexity_prev`, statistics, [all-reduce], M-step.  Returns (perplexity of the incoming state, committed):
        committed is False when `perplexity_prev - perplexity < threshold` (lda_model.py:386-390) — the state is then
        left untouched, as the reference returns before its em_step.  `perplexity_prev=None` skips the test (first
        iteration).  This is the call `bench.py` times as `value`.

        The statistics pass, the all-reduce and the theta refresh are queued BEFORE the host reads the perplexity
        record: they only write scratch / ping-pong buffers, so a converged iteration simply drops them, and the GPU
        never waits for the host's decision."""
        if self._st is None:
            raise RuntimeError("fit() first (em_num_step=0 initialises the state without iterating)")
        self._estep_speculative(e_num_step)
        pending = self._collect_begin(tok_row=0)
        work = self._launch_statistics()
        self.inner_iterations, capped = self._collect_end(pending)
        perplexity = self._perplexity_from_terms(sampling)
        if perplexity_prev is not None and perplexity_prev - perplexity < threshold:
            if work is not None:
                work.wait()  # keep the collective sequence identical on every rank; its result is dropped
            return perplexity, False
        if capped:
            warnings.warn(f"Estep, Maximum iteration reached at step {e_num_step + 1} for {capped} documents")
        self._finish_iteration(work)
        return perplexity, True

    # ------------------------------------------------------------------------------------------------------------------
    # streaming entry point: one EM iteration whose corpus arrives from (pinned) host memory
    # ------------------------------------------------------------------------------------------------------------------
    def em_step_from_host(self, rowptr, colidx, counts, e_num_step: int = 100, chunks: int = 0,
                          sampling: bool = False) -> float:
        """One EM iteration (lda_model.py:285-316 + the perplexity of the incoming state, :403-409) on a CSR corpus
        that lives in host memory (torch CPU tensors, ideally pinned): rowptr int64[D+1], colidx/counts int32[nnz]
        — or compact: colidx uint16/int16 bit patterns when V <= 65536, counts uint8 or uint16, widened on the device —
        same D and V as the fitted corpus.  The upload is pipelined per DOCUMENT BLOCK of the blocked mirror: while
        block b+1 crosses PCIe on a copy stream, block b is mirrored (ttlda_csr_to_csc_block) and E-stepped (leaving
        c/N per nonzero in mirror order); then one statistics pass, theta refresh and M-step.  Bit-identical to uploading the corpus
        and running one iteration of `fit`.  Returns the perplexity of the incoming state on this corpus.  This is
        the per-step path `bench.py` times as `e2e`.  (`chunks` is accepted for compatibility and ignored: the
        pipeline granularity is the mirror's document block.)"""
        st = self._st
        if st is None:
            raise RuntimeError("fit() first (em_num_step=0 initialises the state without iterating)")
        if self._inner_fixed_point or self._sstats_mode != "csc":
            raise NotImplementedError("em_step_from_host supports the default (reference-parity, csc) mode")
        D = rowptr.numel() - 1
        nnz = int(colidx.numel())
        if D != st.D:
            raise ValueError("the streamed corpus must have the fitted number of documents")
        if counts.numel() != nnz:
            raise ValueError("colidx and counts must have one entry per nonzero")
        if not getattr(st, "owns_corpus", False):
            # the streamed arrays overwrite the device CSR / mirror: do that in buffers owned by THIS fit state, not in
            # the DeviceCorpus that Document.device_corpus() caches and hands to later fit() / host_csr() calls
            st.corpus_src = st.corpus
            st.corpus = st.corpus.private_copy()
            st.owns_corpus = True
        c = st.corpus
        if nnz > c.colidx.numel():
            raise ValueError("the streamed corpus has more nonzeros than the device buffers hold")
        dev, V, K, ld = st.dev, st.V, st.K, st.ld
        nb, bd = c.nblocks, c.block_docs
        rp_np = rowptr.numpy()
        dlo = [min(b * bd, D) for b in range(nb + 1)]
        dlo[nb] = D
        plo = [int(rp_np[d]) for d in dlo]
        # The row pointers drive host-side slicing and every launch: they are checked here, before anything is uploaded
        # (~1 ms per million documents, hidden behind the previous step's theta refresh / M-step still on the GPU).
        # Word ids in [0, V) and counts >= 0 are checked AND made harmless on the device as they land
        # (ttlda_widen_i32 / ttlda_validate_csr with sanitize) and reported with the step's one device->host read,
        # before anything is committed.
        if plo[0] != 0 or plo[nb] != nnz or (D and bool((rp_np[1:] < rp_np[:-1]).any())):
            raise ValueError(nv.bad_corpus_message(nv.BAD_ROWPTR))
        max_blk = max((plo[b + 1] - plo[b] for b in range(nb)), default=0)
        # buffers that depend on the streamed corpus' nonzero count
        if c.docidx.numel() < max(nnz, 1):
            c.docidx = torch.empty(nnz, dtype=torch.int32, device=dev)
            c.ccounts = torch.empty(nnz, dtype=torch.int32, device=dev)
            c.csr2csc = torch.empty(nnz, dtype=torch.int32, device=dev)
        hot = bool(getattr(st, "hot", False)) and c.e_idx is not None
        if hot and c.e_idx.numel() < max(nnz, 1):
            c.e_idx = torch.empty(nnz, dtype=torch.int32, device=dev)
            c.e_cnt = torch.empty(nnz, dtype=torch.int32, device=dev)
            c.e_pos = torch.empty(nnz, dtype=torch.int32, device=dev) if c.e_pos is not None else None
        if st.ratio is not None and st.ratio.numel() < max(nnz, 1):
            st.ratio = torch.empty(nnz, dtype=torch.float64, device=dev)
        need = self._sstats_ws_bytes(c, nnz, V, ld)
        if st.ws_s.numel() < need:
            st.ws_s = torch.empty(need, dtype=torch.uint8, device=dev)
        need = nv.csr_to_csc_workspace_bytes(max_blk, bd, V, 1)
        if getattr(st, "ws_c", None) is None or st.ws_c.numel() < need:
            st.ws_c = torch.empty(need, dtype=torch.uint8, device=dev)

        compute = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        # The uploads overwrite the CSR arrays (or their compact staging buffers), which the previous step's widening,
        # mirror and E-step kernels read — but not its statistics pass, theta refresh and M-step (they read the mirror, the
        # ratios and the tables).  After a streamed step the uploads of the next one therefore wait for that step's last
        # E-step only and cross PCIe underneath its statistics pass: with several ranks feeding from one host (8 GPUs:
        # ~150 GB/s for all of them) the transfers, not the GPUs, pace the step, and this keeps the feed continuous
        # across step boundaries.  After anything else the uploads wait for all queued work.
        ev = getattr(st, "ev_corpus_free", None)
        if ev is not None and os.environ.get("TTLDA_UPLOAD_OVERLAP", "1") == "1":
            cs.wait_event(ev)
        else:
            cs.wait_stream(compute)
        st.ev_corpus_free = None
        with torch.cuda.stream(cs):
            c.rowptr.copy_(rowptr, non_blocking=True)
        st.flags.zero_()
        # compact host formats: 2-byte word ids (V <= 65536) and 1- or 2-byte counts, all unsigned, are staged as they
        # are and widened into the int32 arrays on the device (3 instead of 8 bytes per nonzero over PCIe)
        def staged(host, name):
            if host.dtype == torch.int32:
                return host, None
            if host.element_size() not in (1, 2) or host.dtype.is_floating_point:
                raise TypeError(f"{name}: int32, or an unsigned 8/16-bit integer type, expected; got {host.dtype}")
            host = host.view(torch.uint8 if host.element_size() == 1 else torch.int16)
            buf = getattr(st, "stage_" + name, None)
            if buf is None or buf.dtype != host.dtype or buf.numel() < max(nnz, 1):
                buf = torch.empty(max(nnz, 1), dtype=host.dtype, device=dev)
                setattr(st, "stage_" + name, buf)
            return host, buf
        if colidx.element_size() == 2 and V > 65536:
            raise ValueError("16-bit word ids need V <= 65536")
        colidx, stage_ci = staged(colidx, "colidx")
        counts, stage_ct = staged(counts, "counts")
        cur, nxt = st.cur, st.cur ^ 1
        acc = torch.zeros(nv.NPART, dtype=torch.float64, device=dev)
        it_acc = torch.zeros(2, dtype=torch.int64, device=dev)
        use_ratio = st.ratio is not None
        group = int(os.environ.get("TTLDA_ESTEP_GROUP", "0")) or (2 if (stage_ci is not None and stage_ct is not None) else 1)
        last = max((b for b in range(nb) if dlo[b + 1] > dlo[b]), default=-1)
        g0 = 0
        # every block's upload is queued on the copy stream up front (they depend on nothing but the buffers being free)
        arrived = {}
        with torch.cuda.stream(cs):
            for b in range(nb):
                s, e = plo[b], plo[b + 1]
                if dlo[b + 1] <= dlo[b]:
                    continue
                (c.colidx if stage_ci is None else stage_ci)[s:e].copy_(colidx[s:e], non_blocking=True)
                (c.counts if stage_ct is None else stage_ct)[s:e].copy_(counts[s:e], non_blocking=True)
                arrived[b] = torch.cuda.Event()
                arrived[b].record(cs)
        for b in range(nb):
            d0, d1, s, e = dlo[b], dlo[b + 1], plo[b], plo[b + 1]
            if d1 <= d0:
                continue
            compute.wait_event(arrived[b])
            if stage_ci is not None:  # word ids >= V are flagged and zeroed by the widening pass itself
                nv.widen_i32(stage_ci[s:e], c.colidx[s:e], V if V < 65536 else 0, st.flags)
            if stage_ct is not None:
                nv.widen_i32(stage_ct[s:e], c.counts[s:e])
            if stage_ci is None:  # int32 word ids: one validation pass (compact ids are checked while being widened)
                nv.validate_csr(None, c.colidx[s:e], c.counts[s:e] if stage_ct is None else None, V, st.flags,
                                sanitize=True, nnz=e - s)
            colptr_b = c.colptr[b * V:b * V + V + 1]
            nv.csr_to_csc_block(c.rowptr[d0:d1 + 1], c.colidx, c.counts, V, d0, s, e - s, colptr_b, c.docidx,
                                None if use_ratio else c.ccounts, st.ws_c, c.csr2csc if use_ratio else None)
            # E-step over the blocks mirrored since the last launch: every block when the upload paces the step (int32
            # arrays), `group` blocks at a time when the GPU does (compact formats) — fewer launch tails
            if (b + 1 - g0) >= group or b == last:
                e0, e1 = dlo[g0], d1
                if hot:  # same kernel as the resident iteration: hot words (of the FITTED corpus) first, rows cached
                    nv.build_estream(c.rowptr[e0:e1 + 1], c.colidx, c.counts, c.csr2csc if use_ratio else None, V,
                                     c.slot_of_word, c.e_idx, c.e_cnt, c.e_pos if use_ratio else None)
                    nv.estep_hot(c.rowptr[e0:e1 + 1], c.e_idx, c.e_cnt, c.e_pos if use_ratio else None, c.hot_words, K,
                                 V, st.alpha_vec, st.eT[cur][e0:e1], st.eB, st.gamma[cur][e0:e1], st.gamma[nxt][e0:e1],
                                 st.ratio if use_ratio else None, e_num_step, 1e-05, st.iters, st.part[0], st.ws)
                else:
                    nv.estep(c.rowptr[e0:e1 + 1], c.colidx, c.counts, K, V, st.alpha_vec, self._alpha_sum(),
                             st.eT[cur][e0:e1], st.eB, st.gamma[cur][e0:e1], st.gamma[nxt][e0:e1], None, e_num_step,
                             1e-05, st.iters, st.part[0], st.ws, c.csr2csc if use_ratio else None, st.ratio)
                acc += st.part[0]
                it_acc += st.iters
                g0 = b + 1
        if use_ratio:  # (without the ratio hand-off the statistics may rebuild the mirrored counts from the CSR arrays)
            st.ev_corpus_free = torch.cuda.Event()
            st.ev_corpus_free.record(compute)  # nothing queued from here on reads the CSR arrays / staging buffers
        c.nnz = nnz
        c.ccounts_valid = not use_ratio  # the mirrored counts are not rebuilt when the statistics use the ratios
        # one statistics pass over the whole mirror: its ticket order already walks the blocks one L2 window at a time,
        # and the GPU, not PCIe, bounds this step — per-block passes (ttlda_sstats_csc_block_f64) cost 2.3 ms more
        c.block_pos = plo
        st.part[0].copy_(acc)
        st.iters.copy_(it_acc)
        # the record's device->host read is queued behind the last E-step, BEFORE the statistics pass: the host gets the
        # perplexity terms (and the malformed-corpus flags) while that pass runs, and returns to the caller — whose next
        # step's uploads then start underneath it
        pending = self._collect_begin(tok_row=0)
        self._sstats_pass(st.eT[cur], st.ratio)  # writes the statistics table and scratch only: nothing is committed yet
        try:
            self.inner_iterations, capped = self._collect_end(pending)
        except ValueError:
            # a malformed corpus was streamed in (and sanitised on the device): nothing was committed; go back to the
            # corpus the model was fitted on rather than keep the sanitised garbage as the state's corpus
            st.corpus, st.owns_corpus = st.corpus_src, False
            raise
        perplexity = self._perplexity_from_terms(sampling)
        if capped:
            warnings.warn(f"Estep, Maximum iteration reached at step {e_num_step + 1} for {capped} documents")
        self._commit_iteration(have_sstats=True)
        return perplexity

    # ------------------------------------------------------------------------------------------------------------------
    # online learning (the reference's roadmap item, readme.md:131-136)
    # ------------------------------------------------------------------------------------------------------------------
    def partial_fit(self, minibatch, total_docs: int = None, tau0: float = 1.0, kappa: float = 0.7,
                    e_num_step: int = 100, return_perplexity: bool = False, *, gamma_init: float = None,
                    e_threshold: float = 1e-05):
        """One step of online variational Bayes (Hoffman, Blei & Bach 2010, Algorithm 2) on a minibatch of documents.

        E-step: every document's gamma is iterated to its fixed point against the current topics
        (ttlda_estep_fixed_point_f64, gamma_0 = alpha + V/K as in lda_model.py:364); M-step:
            lambda <- (1 - rho_t) lambda + rho_t (eta + (D/|B|) * sstats),   rho_t = (tau0 + t)^(-kappa)
        (ttlda_mstep_online_f64), D = total_docs (default: Document.num_doc_population), t = number of earlier calls.
        The topics persist across calls; lambda_0 is the reference's (seed 42, Gamma(100, .01)).  The reference has no
        such update (its `sampling=True` only rescales the ELBO): this is what gives it one.  With sharded minibatches
        (torch.distributed) the statistics are summed over ranks as in fit().

        return_perplexity: also return the per-word perplexity bound of the minibatch (document part scaled by D/|B| as
        lda_model.py:152-153,187-188 do) under the converged gamma and the topics BEFORE this update — costs one more
        token pass and a device->host read; otherwise the call is fully asynchronous and returns None.

        gamma_init / e_threshold: start value of every gamma_dk (default alpha + V/K, lda_model.py:364) and the L1
        tolerance of the per-document loop (default 1e-5, :202).  With gamma_init=1.0, e_threshold = K * mean_change_tol,
        tau0 = learning_offset + 1 and kappa = learning_decay one call IS one
        sklearn.decomposition.LatentDirichletAllocation._em_step(batch_update=False) without its random gamma start
        (tests/test_sklearn_pin.py)."""
        if not (0.5 < kappa <= 1.0) or tau0 < 0 or tau0 + (getattr(self, "_online_t", None) or 0) <= 0:
            raise ValueError("online LDA needs kappa in (0.5, 1] and tau0 >= 0 (tau0 > 0 for the first step)")
        if isinstance(minibatch, dict) and getattr(getattr(self, "train_doc", None), "vocab_to_idx", None):
            # raw text: every minibatch must index the SAME lambda rows, so it is encoded against the vocabulary of the
            # first corpus this model saw (a fresh first-appearance vocabulary per minibatch would silently permute the
            # word ids whenever the sizes happened to agree)
            doc = Document.from_texts_with_vocab(minibatch, self.train_doc.vocab_to_idx)
            if doc.oov_tokens:
                warnings.warn(f"partial_fit: {doc.oov_tokens} tokens outside the model's vocabulary were dropped")
        else:
            doc = self._as_document(minibatch)
        corpus = doc.device_corpus(self._dev())
        dist = self._dist()
        n_mb = corpus.D
        if dist is not None:
            t = torch.tensor([corpus.D], dtype=torch.int64, device=corpus.device)
            dist.all_reduce(t)
            n_mb = int(t.item())
        if n_mb == 0:
            raise ValueError("empty minibatch")
        total = int(total_docs) if total_docs else max(int(doc.num_doc_population), n_mb)
        if not hasattr(self, "V"):
            self.V = doc.V
        if self.V != doc.V:
            raise ValueError("the minibatch vocabulary size differs from the model's")
        if not hasattr(self, "train_doc"):
            self.train_doc = doc
        if not hasattr(self, "M"):
            self.M = total  # the population the Newton updates (update_alpha) scale with
        # topics already on the device — from earlier partial_fit calls OR from a batch fit() — are what is updated;
        # only a model that has never been fitted starts from the reference's lambda_0
        first = self._st is None
        if getattr(self, "_online_t", None) is None:
            self._online_t = 0
        saved = self._inner_fixed_point
        self._inner_fixed_point = True  # documents are solved to their fixed point, not with the frozen-theta loop
        try:
            if not first:
                self._sync_lambda()  # a sharded batch M-step leaves every rank with its own rows of lambda only
                resharded = getattr(self._st, "beta_sharded", False)
            st = self._alloc_state(corpus, topics_from=None if first else self._st)
            K, V = st.K, st.V
            if not first and resharded:
                self._refresh_beta()  # complete beta-prior term on every rank (the online M-step runs replicated)
            if first:
                np.random.seed(SEED)  # lda_model.py:356-357
                lam0 = np.random.gamma(shape=100, scale=0.01, size=(K, V)).astype(DTYPE, copy=False)
                nv.transpose_pad(torch.as_tensor(lam0).to(st.dev), st.lam)
                self._refresh_beta()
            st.gamma[st.cur].zero_()  # lda_model.py:364
            st.gamma[st.cur][:, :K] = (st.alpha_vec + float(V) / float(K)) if gamma_init is None else float(gamma_init)
            self._refresh_theta()
            self._estep_speculative(e_num_step, e_threshold)
            perplexity = None
            if return_perplexity:
                c, nxt = st.corpus, st.cur ^ 1
                nv.elbo_tokens(c.rowptr, c.colidx, c.counts, K, st.eT[nxt], st.eB, st.part[3], st.ws)
                buf = torch.stack([st.part[3, 0], st.part[0, 1], st.part[3, 2], st.iters[0].to(torch.float64),
                                   st.iters[1].to(torch.float64), st.part[2, 0]])
                if dist is not None:
                    dist.all_reduce(buf[:5])
                h = buf.cpu().numpy()
                self.inner_iterations = int(h[3])
                if int(h[4]):
                    warnings.warn(f"Estep, Maximum iteration reached at step {e_num_step + 1} for {int(h[4])} documents")
                scale = total / n_mb
                elbo = (float(h[0]) + float(h[1])) * scale + float(h[5])
                self._last_elbo = elbo
                perplexity = float(np.exp(np_clip_for_exp(-elbo / (np.float64(h[2]) * scale))))
            rho = float((tau0 + self._online_t) ** (-kappa))
            self._commit_iteration(online=(total / n_mb, rho))
            self._online_t += 1
        finally:
            self._inner_fixed_point = saved
        return perplexity

    def _collect_begin(self, tok_row: int):
        """Queue the device->host read of the terms of the CURRENT state's ELBO (summed over ranks when sharded) and
        return a handle for `_collect_end`.  tok_row = 0: token term / token count from the speculative E-step record;
        3: from ttlda_elbo_tokens_f64.  The theta-prior and beta-prior terms of the current state are kept in
        part[1][0] and part[2][0].

        Nothing here blocks the compute stream: the (48-byte) all-reduce over ranks runs on NCCL's stream behind the
        work queued so far, a side stream waits for it and copies the record into pinned host memory, and the caller
        keeps launching (statistics pass, theta refresh) before it asks for the result."""
        st = self._st
        flags = st.flags.to(torch.float64)
        # [tok, theta, count, passes, capped, flags, beta (this rank's share, sharded M-step) | beta (replicated)]
        beta = st.part[2, 0:1]
        zero = torch.zeros_like(beta)
        sharded = getattr(st, "beta_sharded", False)
        buf = torch.cat([torch.stack([st.part[tok_row, 0], st.part[1, 0], st.part[tok_row, 2],
                                      st.iters[0].to(torch.float64), st.iters[1].to(torch.float64)]),
                         flags, beta if sharded else zero, zero if sharded else beta])
        dist = self._dist()
        work = dist.all_reduce(buf[:7], async_op=True) if dist is not None else None
        if st.dev.type != "cuda":  # only reachable with the test-only CPU stand-ins (tests/cpu_kernels.py, gloo)
            if work is not None:
                work.wait()
            self._host_terms = buf
            return None
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=st.dev)
            self._host_terms = torch.empty(8, dtype=torch.float64).pin_memory()
        main = torch.cuda.current_stream(st.dev)
        side = self._side_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if work is not None:
                work.wait()  # makes the SIDE stream wait for the collective; the compute stream does not
            self._host_terms.copy_(buf, non_blocking=True)
            done = torch.cuda.Event()
            done.record(side)
        buf.record_stream(side)
        return done

    def _collect_end(self, done):
        if done is not None:
            t0 = time.perf_counter()
            done.synchronize()
            self.host_wait_s = getattr(self, "host_wait_s", 0.0) + (time.perf_counter() - t0)
        st = self._st
        h = self._host_terms.numpy().copy()
        if int(h[5]):  # summed over ranks: only "some rank saw a malformed corpus" survives the sum
            st.flags.zero_()
            raise ValueError(nv.bad_corpus_message(int(h[5]) if self._dist() is None else 7)
                             + " — the iteration was not committed")
        st.terms = dict(tok=float(h[0]), theta=float(h[1]), beta=float(h[6]) + float(h[7]), count=float(h[2]))
        return int(h[3]), int(h[4])

    def _collect(self, tok_row: int):
        """Blocking form: queue the read and wait for it."""
        return self._collect_end(self._collect_begin(tok_row))

    def _perplexity_from_terms(self, sampling: bool) -> float:
        t = self._st.terms
        elbo = t["tok"] + t["theta"]  # document-dependent part (lda_model.py:125-148)
        total = np.float64(t["count"])  # np.sum(X) in the reference: a zero count yields inf/nan, not an exception
        if sampling:  # lda_model.py:152-153,187-188
            scale = self.D_population / self._num_doc_global
            elbo = elbo * scale
            total = total * scale
        elbo += t["beta"]  # :157-163
        self._last_elbo = elbo
        return float(np.exp(np_clip_for_exp(-elbo / total)))

    # ------------------------------------------------------------------------------------------------------------------
    # fit  (lda_model.py:318-432)
    # ------------------------------------------------------------------------------------------------------------------
    def fit(self, train_doc, sampling: bool = False, threshold: float = 1e-05, em_num_step: int = 100,
            e_num_step: int = 100, newton_num_step: int = 100, update_hyper_parms: bool = True, batch: bool = True,
            verbose=False, return_perplexities: bool = False):
        train_doc = self._as_document(train_doc)
        corpus = train_doc.device_corpus(self._dev())
        dist = self._dist()
        n_docs = corpus.D
        if dist is not None:
            t = torch.tensor([corpus.D], dtype=torch.int64, device=corpus.device)
            dist.all_reduce(t)
            n_docs = int(t.item())
        if not hasattr(self, "train_doc"):  # lda_model.py:340-348 (cached on first call only)
            self.train_doc = train_doc
        if not hasattr(self, "M"):
            self.M = n_docs
        if not hasattr(self, "V"):
            self.V = train_doc.V
        if self.V != train_doc.V or self.M != n_docs:
            raise ValueError("fit() was first called with a corpus of a different shape (M, V are cached like the "
                             "reference does, lda_model.py:340-348); create a new LDASmoothed")
        self._num_doc_global = n_docs
        if sampling and not hasattr(self, "D_population"):
            # the reference reads self.D_population without ever assigning it (AttributeError);
            # the only population size it knows is Document.num_doc_population — use it.
            self.D_population = train_doc.num_doc_population

        st = self._alloc_state(corpus)
        K, V = st.K, st.V

        np.random.seed(SEED)  # lda_model.py:356-357: legacy MT19937 gamma stream, generated on the host
        lam0 = np.random.gamma(shape=100, scale=0.01, size=(K, V)).astype(DTYPE, copy=False)
        nv.transpose_pad(torch.as_tensor(lam0).to(st.dev), st.lam)
        if verbose:
            print(f"Var Inf - Word Dirichlet prior, Lambda")
            print((K, V))
            print()
        st.gamma[st.cur].zero_()  # lda_model.py:364: alpha + ones * V / K
        st.gamma[st.cur][:, :K] = st.alpha_vec + float(V) / float(K)
        if verbose:
            print(f"Var Inf - Topic Dirichlet prior, Gamma")
            print((self.M, K))
            print()

        perplexities = []
        perplexity = np.inf
        self._refresh_theta()
        self._refresh_beta()
        for it in range(em_num_step):  # lda_model.py:384
            # E-step `it` also completes the perplexity of state `it` (initial perplexity for it == 0, :373-378;
            # perplexity after em_step it-1 otherwise, :403-409)
            perplexity, committed = self.em_iteration(perplexity if it > 0 else None, threshold, sampling, e_num_step)
            perplexities.append(np.float64(perplexity))
            if it == 0 and verbose:
                print(f"Init perplexity = {perplexity}")
            if not committed:
                # :386-390 early return (also skips the hyper-parameter update); the speculative work is dropped
                return perplexities if return_perplexities else None
        # perplexity of the final state (or the initial one when em_num_step == 0)
        self._tokens()
        self._collect(tok_row=3)
        perplexity = self._perplexity_from_terms(sampling)
        perplexities.append(np.float64(perplexity))
        if em_num_step == 0 and verbose:
            print(f"Init perplexity = {perplexity}")

        if update_hyper_parms:  # lda_model.py:414-424
            self.update_alpha()
            self.update_eta()
            # perplexity with the *stale* expectations and the new alpha / eta: only the prior terms change
            # (exp(E[log beta]) is rewritten with identical values: it does not depend on eta)
            self._refresh_theta_term_only()
            self._refresh_beta()
            self._collect(tok_row=3)
            perplexity = self._perplexity_from_terms(sampling)
            perplexities.append(np.float64(perplexity))

        if verbose:
            print(f"End perplexity = {perplexities[-1]}")
        return perplexities if return_perplexities else None

    # ------------------------------------------------------------------------------------------------------------------
    # hyper-parameter Newton updates (lda_model.py:435-554): device reductions with an exact digamma, scalar
    # Newton on the host
    # ------------------------------------------------------------------------------------------------------------------
    def _gamma_psi_stat(self) -> float:
        """sum_dk psi(gamma_dk) - K * sum_d psi(sum_k gamma_dk)   (loop-invariant part of lda_model.py:459-462)."""
        st = self._st
        nv.hyper_stats(st.gamma[st.cur], st.K, st.part[4], st.ws)
        v = st.part[4, :2].clone()
        dist = self._dist()
        if dist is not None:
            dist.all_reduce(v)
        h = v.cpu().numpy()
        return float(h[0] - st.K * h[1])

    def _lambda_psi_stat(self) -> float:
        """sum_kv psi(lambda_kv) - V * sum_k psi(sum_v lambda_kv)   (lda_model.py:522-528)."""
        st = self._st
        self._sync_lambda()
        nv.hyper_stats(st.lam, st.K, st.part[4], st.ws)
        pc = torch.empty(st.K, dtype=torch.float64, device=st.dev)
        nv.psi_exact(st.colsum[:st.K].contiguous(), pc)
        return float(st.part[4, 0].item() - st.V * pc.sum().item())

    def _gamma_psi_colstat(self) -> np.ndarray:
        """(psi(gamma) - psi(rowsum)).sum(axis=0): the K-vector of lda_model.py:452-456 (loop-invariant)."""
        st = self._st
        out = torch.empty(st.K, dtype=torch.float64, device=st.dev)
        nv.hyper_colstats(st.gamma[st.cur], st.K, out)
        dist = self._dist()
        if dist is not None:
            dist.all_reduce(out)
        return out.cpu().numpy()

    def update_alpha(self, step: int = 500, threshold: float = 1e-07, verbose: bool = False) -> None:
        if isinstance(self._alpha_, np.ndarray):
            # Non-exchangeable branch (lda_model.py:452-456,464-476), reproduced AS IS — including `sum_alpha =
            # alpha * K` (:446, an array, where the sum of alpha was meant), which makes the iteration diverge when it
            # is run to its cap: best effort, the first Newton iterations are pinned by a golden trace.
            stat = self._gamma_psi_colstat()
            M, K = self.M, self.K
            it = 0
            while it <= step:
                sum_alpha = self._alpha_ * K  # :446
                g = M * (psi(sum_alpha) - psi(self._alpha_)) + stat  # :452-456
                h = -M * polygamma(1, self._alpha_)  # :467
                z = M * polygamma(1, sum_alpha)  # :470
                c = np.sum(g / h) / ((1. / z) + np.sum(1. / h))  # :473
                alpha_new = self._alpha_ - (g - c) / h  # :476, :485
                delta = np.sum(np.abs(alpha_new - self._alpha_))
                if verbose:
                    print(f"M Step: Iteration {it}, Delta Alpha = {delta}")
                    print(f"Alpha Old:{self._alpha_} -> Alpha New:{alpha_new}")
                self._alpha_ = alpha_new
                it += 1
                if delta < threshold:
                    self._upload_alpha()
                    return
            self._upload_alpha()
            warnings.warn(f"Update alphda Maximum iteration reached at step {it}")
            return
        stat = self._gamma_psi_stat()
        M, K = self.M, self.K
        it = 0
        while it <= step:
            g = M * K * (psi(K * self._alpha_) - psi(self._alpha_)) + stat  # :459-462
            h = M * K * (K * polygamma(1, K * self._alpha_) - polygamma(1, self._alpha_))  # :479
            alpha_new = self._alpha_ - g / h
            delta = np.sum(np.abs(alpha_new - self._alpha_))
            if verbose:
                print(f"M Step: Iteration {it}, Delta Alpha = {delta}")
                print(f"Alpha Old:{self._alpha_} -> Alpha New:{alpha_new}")
            self._alpha_ = alpha_new
            it += 1
            if delta < threshold:
                self._upload_alpha()
                return
        self._upload_alpha()
        warnings.warn(f"Update alphda Maximum iteration reached at step {it}")

    def update_eta(self, step: int = 1000, threshold: float = 1e-09, verbose: bool = False) -> None:
        stat = self._lambda_psi_stat()
        K, V = self.K, self.V
        it = 0
        while it <= step:
            g = K * V * (psi(V * self._eta_) - psi(self._eta_)) + stat  # :525-528
            h = K * V * (V * polygamma(1, V * self._eta_) - polygamma(1, self._eta_))  # :531
            eta_new = self._eta_ - g / h
            if eta_new < 0:
                raise ValueError(f"Dirichlet Parameter become < 0 at iteration {it}")
            delta = np.abs(eta_new - self._eta_)
            if verbose:
                print(f"M Step: delta eta is {delta}")
                print(f"Eta Old {self._eta_} -> Eta New {eta_new}")
            self._eta_ = eta_new
            it += 1
            if delta < threshold:
                return
        warnings.warn(f"Update Eta: Maximum iteration reached at step {it}")

    # ------------------------------------------------------------------------------------------------------------------
    # reference-shaped step API (NumPy in / NumPy out); thin wrappers over the same kernels
    # ------------------------------------------------------------------------------------------------------------------
    def _elog_np(self):
        st = self._st
        self._sync_lambda()
        et = torch.empty_like(st.gamma[st.cur])
        nv.dirichlet_expectation(st.gamma[st.cur], st.K, out_elog=et)
        eb = torch.empty_like(st.lam)
        self_colwise = torch.empty((st.K, st.V), dtype=torch.float64, device=st.dev)
        nv.beta_refresh(st.lam, st.K, float(self._eta_), st.eB.clone(), eb, st.colsum, st.part[2], st.ws_m)
        nv.transpose_unpad(eb, st.K, self_colwise)
        return et[:, :st.K].cpu().numpy(), self_colwise.cpu().numpy()

    def _check_X(self, X):
        if X is not None and self._st is not None:
            shape = getattr(X, "shape", None)
            if shape is not None and tuple(shape) != (self._st.D, self._st.V):
                raise ValueError("X does not match the fitted corpus")

    def _load_elogs(self, expec_log_theta, expec_log_beta):
        st = self._st
        if expec_log_theta is not None:
            e = torch.as_tensor(np.ascontiguousarray(expec_log_theta, dtype=np.float64)).to(st.dev)
            st.eT[st.cur].zero_()
            st.eT[st.cur][:, :st.K] = torch.exp(e)
        else:
            self._refresh_theta()
        if expec_log_beta is not None:
            e = torch.exp(torch.as_tensor(np.ascontiguousarray(expec_log_beta, dtype=np.float64)).to(st.dev))
            nv.transpose_pad(e, st.eB)
        else:
            self._refresh_beta()

    def approx_elbo(self, X=None, sampling: bool = False, expec_log_theta=None, expec_log_beta=None):
        """lda_model.py:89-165.  X must be the fitted corpus (or None).  Explicit expectations are honoured for the
        token term and returned unchanged; the prior terms use DE(gamma) / DE(lambda), which is what every reference
        call site passes."""
        self._check_X(X)
        self._load_elogs(expec_log_theta, expec_log_beta)
        self._refresh_theta_term_only()
        st = self._st
        self._sync_lambda()
        nv.beta_refresh(st.lam, st.K, float(self._eta_), st.eB.clone(), None, st.colsum, st.part[2], st.ws_m)
        st.beta_sharded = False
        self._tokens()
        self._collect(tok_row=3)
        self._perplexity_from_terms(sampling)
        if expec_log_theta is None or expec_log_beta is None:
            et, eb = self._elog_np()
            expec_log_theta = et if expec_log_theta is None else expec_log_theta
            expec_log_beta = eb if expec_log_beta is None else expec_log_beta
        return self._last_elbo, (expec_log_theta, expec_log_beta)

    def approx_perplexity(self, X=None, sampling: bool = False, expec_log_theta=None, expec_log_beta=None):
        """lda_model.py:167-194."""
        _, logs = self.approx_elbo(X, sampling, expec_log_theta, expec_log_beta)
        return self._perplexity_from_terms(sampling), logs

    def e_step(self, X, expec_log_theta, expec_log_beta, step: int = 100, threshold: float = 1e-05,
               batch: bool = True, verbose: bool = False):
        """lda_model.py:197-270: returns (gamma, (lambda_update (K,V), Elogtheta)).  With inner_fixed_point=True the
        per-document loop refreshes exp(E[log theta_d]) every pass (scikit-learn's _update_doc_distribution)."""
        self._check_X(X)
        self._load_elogs(expec_log_theta, expec_log_beta)
        st, c = self._st, self._st.corpus
        nxt = st.cur ^ 1
        self._estep_speculative(step, threshold)
        eT = st.eT[nxt] if self._inner_fixed_point else st.eT[st.cur]
        if self._sstats_mode == "atomic":
            nv.sstats_atomic(c.rowptr, c.colidx, c.counts, st.K, st.V, eT, st.eB, st.sstats)
        else:
            if st.ratio is None and not getattr(c, "ccounts_valid", True):
                c.build_csc(c.nblocks, c.block_docs)
            self._sstats_pass(eT, st.ratio)
        dist = self._dist()
        if dist is not None:
            dist.all_reduce(st.sstats)
        upd = torch.empty((st.K, st.V), dtype=torch.float64, device=st.dev)
        nv.transpose_unpad(st.sstats * st.eB, st.K, upd)  # lda_model.py:264
        st.cur = nxt
        if not (self._inner_fixed_point or self._fused_refresh):
            self._refresh_theta()  # the E-step wrote gamma only
        else:
            st.part[1, 0].copy_(st.part[0, 1])
        self._invalidate("gamma")
        h = st.iters.cpu()
        self.inner_iterations, capped = int(h[0]), int(h[1])
        if capped:
            warnings.warn(f"Estep, Maximum iteration reached at step {step + 1} for {capped} documents")
        et = torch.empty_like(st.gamma[st.cur])
        nv.dirichlet_expectation(st.gamma[st.cur], st.K, out_elog=et)
        return self._gamma_, (upd.cpu().numpy(), et[:, :st.K].cpu().numpy())

    def m_step(self, lambda_update, verbose: bool = False, batch: bool = True):
        """lda_model.py:273-283: lambda = eta + lambda_update; returns (lambda, Elogbeta)."""
        st = self._st
        lam = float(self._eta_) + torch.as_tensor(np.ascontiguousarray(lambda_update, dtype=np.float64)).to(st.dev)
        nv.transpose_pad(lam, st.lam)
        st.lam_synced, st.beta_sharded = True, False
        self._invalidate("lam")
        eb = torch.empty_like(st.lam)
        nv.beta_refresh(st.lam, st.K, float(self._eta_), st.eB, eb, st.colsum, st.part[2], st.ws_m)
        out = torch.empty((st.K, st.V), dtype=torch.float64, device=st.dev)
        nv.transpose_unpad(eb, st.K, out)
        return self._lambda_, out.cpu().numpy()

    def em_step(self, X, expec_log_theta, expec_log_beta, step: int = 100, threshold: float = 1e-05,
                batch: bool = True, verbose: bool = False):
        """lda_model.py:285-316: returns (Elogtheta, Elogbeta) after one E+M step."""
        _, (upd, elogtheta) = self.e_step(X, expec_log_theta, expec_log_beta, step, threshold, verbose)
        _, elogbeta = self.m_step(upd, verbose)
        return elogtheta, elogbeta

    def get_top_k_words(self, k: int):
        """lda_model.py:556-564."""
        lam = self._lambda_
        for topic_index in range(lam.shape[0]):
            topk = np.argsort(lam[topic_index, :], )[-k:]
            print(f"Topic {topic_index}")
            for i, idx in enumerate(topk):
                print(f"Top {i + 1} -> {self.train_doc.idx_to_vocab[idx]}")
            print()
