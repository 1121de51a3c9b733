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
