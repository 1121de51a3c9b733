"""Host-side text pipeline for ``Document`` (reference: tuotuo/utils.py:71-86 ``text_pipeline`` and
tuotuo/text_pre_processor.py:23-69).

The reference depends on NLTK (stop-word corpus + ToktokTokenizer), which is not installable here; the
stop-word list below is the public NLTK "english" list (179 entries) minus ``'not'`` (the reference removes
it at text_pre_processor.py:14).  After punctuation has been stripped the Toktok tokenizer degenerates to a
whitespace split, which is what is used here.  On the tokens emitted by ``doc_generator`` and by the
synthetic corpora the pipeline is the identity apart from stop-word removal, so vocab ids are bit-exact with
the reference (tests/test_document.py, golden fixture ``document_vocab.json``).
"""
from __future__ import annotations

import re
import string
import unicodedata

NLTK_ENGLISH_STOPWORDS = [
    "i", "me", "my", "myself", "we", "our", "ours", "ourselves", "you", "you're", "you've", "you'll", "you'd",
    "your", "yours", "yourself", "yourselves", "he", "him", "his", "himself", "she", "she's", "her", "hers",
    "herself", "it", "it's", "its", "itself", "they", "them", "their", "theirs", "themselves", "what", "which",
    "who", "whom", "this", "that", "that'll", "these", "those", "am", "is", "are", "was", "were", "be", "been",
    "being", "have", "has", "had", "having", "do", "does", "did", "doing", "a", "an", "the", "and", "but", "if",
    "or", "because", "as", "until", "while", "of", "at", "by", "for", "with", "about", "against", "between",
    "into", "through", "during", "before", "after", "above", "below", "to", "from", "up", "down", "in", "out",
    "on", "off", "over", "under", "again", "further", "then", "once", "here", "there", "when", "where", "why",
    "how", "all", "any", "both", "each", "few", "more", "most", "other", "some", "such", "no", "nor", "not",
    "only", "own", "same", "so", "than", "too", "very", "s", "t", "can", "will", "just", "don", "don't",
    "should", "should've", "now", "d", "ll", "m", "o", "re", "ve", "y", "ain", "aren", "aren't", "couldn",
    "couldn't", "didn", "didn't", "doesn", "doesn't", "hadn", "hadn't", "hasn", "hasn't", "haven", "haven't",
    "isn", "isn't", "ma", "mightn", "mightn't", "mustn", "mustn't", "needn", "needn't", "shan", "shan't",
    "shouldn", "shouldn't", "wasn", "wasn't", "weren", "weren't", "won", "won't", "wouldn", "wouldn't",
]

# text_pre_processor.py:14 — 'not' is kept as a content word
STOPWORDS = frozenset(w for w in NLTK_ENGLISH_STOPWORDS if w != "not")

_SPECIAL = re.compile(r"[^a-zA-z0-9.,!?/:;\"\'\s]")  # text_pre_processor.py:30 (the A-z range is the reference's)
_SPACES = re.compile(r"^\s*|\s\s*")  # text_pre_processor.py:68
_PUNCT = str.maketrans("", "", string.punctuation)


def remove_accented_chars(text: str) -> str:  # text_pre_processor.py:23-25
    return unicodedata.normalize("NFKD", text).encode("ascii", "ignore").decode("utf-8", "ignore")


def remove_special_characters(text: str) -> str:  # text_pre_processor.py:28-31
    return _SPECIAL.sub(" ", text)


def remove_punctuation(text: str) -> str:  # text_pre_processor.py:40-42
    return text.translate(_PUNCT)


def remove_extra_whitespace_tabs(text: str) -> str:  # text_pre_processor.py:66-69
    return _SPACES.sub(" ", text).strip()


def remove_stopwords(text: str) -> str:  # text_pre_processor.py:56-63
    tokens = [t.strip() for t in text.split()]
    return " ".join(t for t in tokens if t.lower() not in STOPWORDS)


def text_pipeline(s: str) -> str:  # utils.py:71-86 (same order of stages)
    s = remove_accented_chars(s)
    s = remove_special_characters(s)
    s = remove_punctuation(s)
    s = remove_extra_whitespace_tabs(s)
    s = remove_stopwords(s)
    return s


# ----------------------------------------------------------------------------------------------------------------------
# Whole-corpus form of the same pipeline: the reference runs `text_pipeline` once per document in
# Python (utils.py:96-98) and then counts words with dict/list loops (utils.py:177-206), O(V*M).  Here the three
# character-level stages run ONCE over all documents joined by a separator, the split into tokens is one C call, and
# stop words are dropped by a vectorised mask over the token CODES — Python touches each distinct word once, not each
# token.  Same tokens, same order, same first-appearance vocabulary ids as the per-document form
# (tests/test_host_logic.py::test_corpus_tokenizer_equals_per_document_pipeline).
# ----------------------------------------------------------------------------------------------------------------------
_SEP = "\x1d"  # ASCII group separator: whitespace for `\s` (kept by stage 2), not punctuation (kept by stage 3)


def tokenize_corpus(texts):
    """texts: sequence of str.  Returns (codes int32[T], offsets int64[len(texts)+1], vocab list[str]) where
    vocab[codes[offsets[d]:offsets[d+1]]] are the tokens `text_pipeline(texts[d]).split(' ')` yields (an empty document
    yields the single token '' like the reference) and vocabulary ids are assigned by first appearance."""
    import numpy as np
    import pandas as pd

    n = len(texts)
    if n == 0:
        return np.empty(0, dtype=np.int32), np.zeros(1, dtype=np.int64), []
    if any(_SEP in t for t in texts):  # the separator occurs in the data: per-document path
        docs = [text_pipeline(t).split(" ") for t in texts]
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(d) for d in docs], out=offs[1:])
        codes, uniques = pd.factorize(np.asarray([w for d in docs for w in d], dtype=object), sort=False)
        return codes.astype(np.int32), offs, uniques.tolist()
    joined = (" " + _SEP + " ").join(texts)
    joined = remove_punctuation(remove_special_characters(remove_accented_chars(joined)))  # stages 1-3, char level
    toks = np.asarray(joined.split(), dtype=object)  # stage 4 (whitespace runs) is what split() does
    # NB: _SEP is whitespace for str.split() too, so the boundaries are recovered from a second, separator-only pass
    sep_free = joined.split(_SEP)
    if len(sep_free) != n:
        raise AssertionError("document separator lost in the character-level stages")
    lens_raw = np.fromiter((len(x.split()) for x in sep_free), dtype=np.int64, count=n)
    codes, uniques = pd.factorize(toks, sort=False)
    keep_word = np.fromiter((w.lower() not in STOPWORDS for w in uniques), dtype=bool, count=len(uniques))  # stage 5
    keep = keep_word[codes] if len(codes) else np.zeros(0, dtype=bool)
    doc_of = np.repeat(np.arange(n), lens_raw)
    lens = np.bincount(doc_of[keep], minlength=n).astype(np.int64)
    codes = codes[keep]
    empty = lens == 0
    if empty.any() or not keep_word.all():
        # re-number by first appearance among the KEPT tokens, with '' for documents left without any token
        words = uniques.to_numpy(dtype=object) if hasattr(uniques, "to_numpy") else np.asarray(uniques, dtype=object)
        if empty.any():
            blank = len(words)
            words = np.append(words, "")
            lens2 = np.where(empty, 1, lens)
            offs = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(lens2, out=offs[1:])
            out = np.empty(int(offs[-1]), dtype=np.int64)
            fill = np.ones(int(offs[-1]), dtype=bool)
            fill[offs[:-1][empty]] = False
            out[fill] = codes
            out[~fill] = blank
            codes, lens = out, lens2
        new_codes, first = pd.factorize(codes, sort=False)
        vocab = [words[i] for i in first]
        codes = new_codes
    else:
        vocab = uniques.tolist()
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    return codes.astype(np.int32), offs, vocab
