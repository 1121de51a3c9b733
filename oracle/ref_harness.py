"""oracle/ref_harness.py — run the REAL reference (``/root/reference/tuotuo``) in this process.

TEST INFRASTRUCTURE ONLY, and only usable where ``/root/reference`` exists (the build container, not the
GPU box).  Used by ``tests/golden/make_golden.py`` to produce the committed golden fixtures and by the
optional ``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent).

Nothing is copied: the reference's pure-Python modules are imported from where they lie, with
 * ``tuotuo.cutils``  -> the module compiled by ``oracle/build_ref.py`` into ``oracle/_ref/``;
 * ``spacy`` / ``nltk`` / ``pyro`` (not installed, no network) -> minimal ``sys.modules`` stubs, as
   described in SURVEY.md §8(c).  The numeric path (``lda_model.py``) touches none of them.

Because the repo also ships a drop-in package importable as ``tuotuo``, this harness must run in a process
that has not imported the repo's ``tuotuo`` shim; ``import_reference()`` asserts that.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = "/root/reference"
# on the GPU box /root/reference does not exist: the unmodified modules staged by oracle/build_ref.py are used
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    for root in ((STAGED_ROOT,) if os.environ.get("TTLDA_REF_STAGED_ONLY") else (REF_ROOT, STAGED_ROOT)):
        if os.path.exists(os.path.join(root, "tuotuo", "lda_model.py")):
            return root
    return None


def reference_available() -> bool:
    return reference_root() is not None


def _install_stubs(stopwords):
    import torch

    spacy = types.ModuleType("spacy")
    nltk = types.ModuleType("nltk")
    corpus = types.ModuleType("nltk.corpus")
    tokenize = types.ModuleType("nltk.tokenize")

    class _Words:
        def __init__(self, w):
            self._w = list(w)

        def words(self, *a, **k):
            return list(self._w)

    corpus.stopwords = _Words(stopwords)
    corpus.words = _Words([])

    class ToktokTokenizer:
        def tokenize(self, s):
            return s.split()

    tokenize.ToktokTokenizer = ToktokTokenizer
    nltk.corpus = corpus
    nltk.tokenize = tokenize
    nltk.wordpunct_tokenize = lambda s: s.split()

    pyro = types.ModuleType("pyro")
    pdist = types.ModuleType("pyro.distributions")

    class _Callable:
        def __call__(self, sample_shape=()):
            return self.sample(torch.Size(sample_shape))

    class Dirichlet(_Callable, torch.distributions.Dirichlet):
        pass

    class Categorical(_Callable, torch.distributions.Categorical):
        pass

    pdist.Dirichlet = Dirichlet
    pdist.Categorical = Categorical
    pyro.distributions = pdist
    for name, mod in [("spacy", spacy), ("nltk", nltk), ("nltk.corpus", corpus), ("nltk.tokenize", tokenize),
                      ("pyro", pyro), ("pyro.distributions", pdist)]:
        sys.modules.setdefault(name, mod)


def import_reference(stopwords=None):
    """Returns the reference's (lda_model, document, generator, cutils) modules."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not present at /root/reference nor staged in oracle/_ref/tuotuo")
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    from build_ref import build_ref, load_ref_cutils

    build_ref()
    cutils = load_ref_cutils()
    loaded = sys.modules.get("tuotuo")
    if loaded is not None and not str(getattr(loaded, "__file__", "")).startswith(root):
        raise RuntimeError("the repo's own `tuotuo` shim is already imported in this process")
    if stopwords is None:
        stopwords = ["not", "the", "a", "an", "and", "of"]
    _install_stubs(stopwords)
    sys.modules["tuotuo.cutils"] = cutils
    sys.path.insert(0, root)
    import tuotuo.lda_model as lda_model  # noqa: E402
    import tuotuo.document as document  # noqa: E402
    import tuotuo.generator as generator  # noqa: E402

    assert lda_model.__file__.startswith(root)
    return lda_model, document, generator, cutils


class DenseCorpus:
    """Duck-typed stand-in for the reference's ``Document``: ``fit`` reads only .M, .V and
    .doc_vocab_count_array (tuotuo/lda_model.py:344-351)."""

    def __init__(self, X):
        self.doc_vocab_count_array = X
        self.M, self.V = X.shape
        self.idx_to_vocab = {i: f"w{i}" for i in range(self.V)}
