"""tuotuo_b200 — B200-native (sm_100a) implementation of TuoTuo's batch variational-inference LDA hot path.

Public API mirrors the reference package (``tuotuo.lda_model.LDASmoothed``, ``tuotuo.document.Document``,
``tuotuo.generator.doc_generator``); the kernels live in ``libttlda.so`` behind the C ABI of ``include/ttlda.h``.
Submodules are imported lazily so that host-only pieces (vocabulary building, sharding) work without CUDA.
"""
__version__ = "0.1.0"

_LAZY = {
    "LDASmoothed": ("lda_model", "LDASmoothed"),
    "Document": ("document", "Document"),
    "doc_generator": ("generator", "doc_generator"),
    "DeviceCorpus": ("corpus", "DeviceCorpus"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
