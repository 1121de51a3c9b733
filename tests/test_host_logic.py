"""Host-side logic that needs no GPU: text pipeline, vocabulary ids, Document API surface, sharding, generator."""
import json
import os

import numpy as np
import pytest

from tuotuo_b200.corpus import balanced_doc_partition
from tuotuo_b200.document import Document
from tuotuo_b200.text import STOPWORDS, text_pipeline


def test_vocab_ids_bit_exact_vs_reference_cfg1(golden_dir, capsys):
    j = json.load(open(os.path.join(golden_dir, "cfg1_docs.json")))
    docs = {int(k): v for k, v in j["docs"].items()}
    d = Document(docs)
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "There are 100 documents in the dataset after processing"
    assert out[1] == "On average estimated document length is 20.0 words per document after processing"
    assert out[2] == "There are 40 unique vocab in the corpus after processing"
    assert d.vocab_to_idx == j["vocab_to_idx"] and list(d.vocab_to_idx) == list(j["vocab_to_idx"])
    assert d.sample_docs_list == j["sample_docs_list"]
    assert d.M == 100 and d.V == 40 and d.num_doc_population == 100
    assert d.idx_to_vocab[0] == list(j["vocab_to_idx"])[0]
    g = np.load(os.path.join(golden_dir, "fit_cfg1.npz"))
    # token stream -> per-document sorted unique counts must equal the reference's dense matrix rows
    for i in (0, 17, 99):
        ids = d._token_ids[d._doc_offsets[i]:d._doc_offsets[i + 1]]
        u, c = np.unique(ids, return_counts=True)
        np.testing.assert_array_equal(u, g["colidx"][g["rowptr"][i]:g["rowptr"][i + 1]])
        np.testing.assert_array_equal(c, g["counts"][g["rowptr"][i]:g["rowptr"][i + 1]])


def test_text_pipeline_vs_reference_free_text(golden_dir):
    """Accents, punctuation, case-sensitivity, stop words ('not' is kept), empty document -> ''."""
    j = json.load(open(os.path.join(golden_dir, "document_vocab.json")))
    d = Document(dict(j["texts"]))
    assert d.sample_docs_list == j["sample_docs_list"]
    assert d.vocab_to_idx == j["vocab_to_idx"] and list(d.vocab_to_idx) == list(j["vocab_to_idx"])
    assert d.M == j["M"] and d.V == j["V"] and d.num_doc_population == j["num_doc_population"]
    assert "" in d.vocab_to_idx  # the empty document contributes the token ''
    assert "not" not in STOPWORDS and "the" in STOPWORDS
    assert text_pipeline("  The   fox!!  ") == "fox"


def test_num_doc_population():
    d = Document({0: "a b", 1: "b c"}, num_doc_population=1000)
    assert d.num_doc_population == 1000


def test_from_dense_and_shard_roundtrip():
    rs = np.random.RandomState(0)
    X = rs.poisson(0.3, size=(37, 23)).astype(float)
    X[5] = 0
    d = Document.from_dense(X)
    rp, ci, ct = d._host_csr
    assert d.M == 37 and d.V == 23 and rp[-1] == np.count_nonzero(X)
    np.testing.assert_array_equal(d.doc_vocab_count_array, X)
    assert d.doc_vocab_count_df.shape == (37, 23)
    assert d.vocab_doc_count_dict["w3"] == [int(v) for v in X[:, 3]]
    parts = [d.shard(r, 3) for r in range(3)]
    assert sum(p.M for p in parts) == 37
    np.testing.assert_array_equal(np.vstack([p.doc_vocab_count_array for p in parts]), X)
    assert parts[0].doc_range[0] == 0 and parts[-1].doc_range[1] == 37
    with pytest.raises(ValueError):
        Document.from_dense(X + 0.5)


def test_balanced_partition_properties():
    rs = np.random.RandomState(1)
    lens = rs.poisson(50, size=1000)
    lens[::7] = 0
    rowptr = np.concatenate([[0], np.cumsum(lens)])
    for world in (1, 2, 4, 8):
        b = balanced_doc_partition(rowptr, world)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0) and len(b) == world + 1
        loads = np.diff(rowptr[b])
        assert loads.max() - loads.min() <= 2 * lens.max()
    b = balanced_doc_partition(np.array([0, 0, 0, 0]), 2)  # all-empty corpus
    assert b[0] == 0 and b[-1] == 3


def test_shard_token_stream():
    ids = np.arange(30, dtype=np.int32) % 7
    offs = np.array([0, 5, 5, 12, 20, 30], dtype=np.int64)
    d = Document.from_token_ids(ids, offs, 7)
    p0, p1 = d.shard(0, 2), d.shard(1, 2)
    assert p0.M + p1.M == 5
    np.testing.assert_array_equal(np.concatenate([p0._token_ids, p1._token_ids]), ids)
    assert p0._doc_offsets[0] == 0 and p1._doc_offsets[0] == 0


def test_doc_generator_api():
    import torch

    from tuotuo.generator import doc_generator

    torch.manual_seed(3)
    g = doc_generator(M=7, L=11, topic_prior=torch.tensor([1, 1, 1, 1, 1], dtype=torch.double))
    assert g.K == 5 and g.V == 40 and g.M == 7 and g.L == 11
    assert g.beta.shape == (5, 40) and g.beta.dtype == torch.double
    assert torch.allclose(g.beta.sum(1), torch.ones(5, dtype=torch.double))
    assert g.theta.shape == (7, 5)
    assert g.words[13]["name"] == "Olympic" and g.topics[4] == "law" and g.words[0]["prob"][0] == 0.96
    docs = g.generate_doc()
    assert list(docs) == list(range(7)) and all(len(v.split(" ")) == 11 for v in docs.values())
    assert docs is g.docs
    names = {w["name"] for w in g.words.values()}
    assert all(t in names for v in docs.values() for t in v.split(" "))
    with pytest.raises(AssertionError):
        doc_generator(M=1, L=1, topic_prior=torch.ones(4, dtype=torch.double))


def test_lda_constructor_prints_and_priors(capsys):
    from tuotuo.lda_model import LDASmoothed

    m = LDASmoothed(num_topics=5)
    out = capsys.readouterr().out
    assert "Topic Dirichlet Prior, α\n1\n" in out and "Exchangeable Word Dirichlet Prior, Eta \n1\n" in out
    assert m._alpha_ == 1 and m._eta_ == 1 and m.K == 5
    m2 = LDASmoothed(num_topics=6, exchangeable_prior=False, verbose=False)
    np.random.seed(42)
    np.testing.assert_array_equal(m2._alpha_, np.random.gamma(shape=100, scale=0.01, size=(1, 6)))
    with pytest.raises(AttributeError):
        m._lambda_


def test_uci_bag_of_words_reader(tmp_path):
    """docword/vocab files (UCI format, 1-based, unsorted, with a duplicate line) -> CSR."""
    (tmp_path / "docword.toy.txt").write_text("4\n5\n7\n3 2 4\n1 1 2\n1 5 1\n3 1 1\n4 3 9\n1 1 3\n3 5 2\n")
    (tmp_path / "vocab.toy.txt").write_text("alpha\nbeta\ngamma\ndelta\nepsilon\n")
    d = Document.from_uci(tmp_path / "docword.toy.txt", tmp_path / "vocab.toy.txt")
    rp, ci, ct = d._host_csr
    assert d.M == 4 and d.V == 5
    np.testing.assert_array_equal(rp, [0, 2, 2, 5, 6])  # document 2 is empty
    np.testing.assert_array_equal(ci, [0, 4, 0, 1, 4, 2])
    np.testing.assert_array_equal(ct, [5, 1, 1, 4, 2, 9])  # the duplicated (1,1) line is summed
    assert d.idx_to_vocab[2] == "gamma" and d.vocab_to_idx["epsilon"] == 4
    X = d.doc_vocab_count_array
    assert X[0, 0] == 5 and X[3, 2] == 9 and X.sum() == 22
    (tmp_path / "bad.txt").write_text("1\n2\n2\n1 1 1\n")
    with pytest.raises(ValueError):
        Document.from_uci(tmp_path / "bad.txt")


def test_from_csr_rejects_malformed_corpora():
    """Document.from_csr checks what will drive device addressing: ValueError, as NumPy's IndexError would
    have been in the reference's dense indexing (tuotuo/lda_model.py:226-227)."""
    import numpy as np
    import pytest

    from tuotuo_b200.document import Document

    rp, ci, ct = np.array([0, 2, 2, 5]), np.array([1, 3, 0, 2, 4]), np.array([1, 2, 1, 1, 3])
    d = Document.from_csr(rp, ci, ct, 5)
    assert d.M == 3 and d.V == 5
    for bad in [(np.array([0, 2, 1, 5]), ci, ct), (rp, np.array([1, 3, 0, 2, 5]), ct), (rp, np.array([1, 1, 0, 2, 4]), ct),
                (rp, ci, -ct), (np.array([1, 2, 2, 5]), ci, ct), (rp, np.array([3, 1, 0, 2, 4]), ct), (rp, ci, ct[:4]),
                (rp, np.array([1, 3, 0, 2, 2 ** 32 + 1]), ct)]:
        with pytest.raises(ValueError):
            Document.from_csr(*bad, 5)


def test_corpus_tokenizer_equals_per_document_pipeline():
    """text.tokenize_corpus (one vectorised pass over the whole corpus) yields exactly what the reference's
    per-document `text_pipeline(doc).split(' ')` + first-appearance ids do (tuotuo/utils.py:71-98,177-206): accents,
    special characters, punctuation, whitespace runs, stop words, empty documents ('' gets a vocabulary id)."""
    import random

    import numpy as np

    from tuotuo_b200.document import Document
    from tuotuo_b200.text import text_pipeline, tokenize_corpus

    random.seed(3)
    words = ["the", "Olympic", "FIFA", "not", "café", "naïve", "hello", "world", "don't", "isn't", "a", "I", "x1",
             "42", "re-use", "e.g.", "…", "日本", "tab\there", "new\nline", "", "  ", "^caret", "a_b"]

    def make():
        return " ".join(random.choice(words) + random.choice(["", ",", "!", "  ", "?"]) for _ in range(random.randint(0, 14)))

    texts = [make() for _ in range(1500)] + ["", "the a an", "   ", "!!!"]
    for corpus in (texts, ["only one"], ["", ""], texts[:7] + ["with \x1d separator inside"] + texts[7:20]):
        codes, offs, vocab = tokenize_corpus(corpus)
        docs = [text_pipeline(t).split(" ") for t in corpus]
        seen, ids = {}, []
        for d in docs:
            for w in d:
                ids.append(seen.setdefault(w, len(seen)))
        assert vocab == list(seen)
        assert np.array_equal(codes, np.asarray(ids, dtype=np.int32))
        assert np.array_equal(np.diff(offs), [len(d) for d in docs])
    d = Document({i: t for i, t in enumerate(texts[:50])})
    assert d.sample_docs_list == [text_pipeline(t).split(" ") for t in texts[:50]]
    assert d.M == 50 and d.V == len(d.vocab_to_idx)


def test_synthetic_shards_share_one_model():
    """synthetic_token_stream: `seed` fixes the generative model, `doc_seed` the documents drawn from it.  Weak-scaling
    shards (bench.py, N > 1) are documents of ONE corpus: same topics, different documents, statistically equal work
    per rank; doc_seed=None keeps the single-stream corpus every N = 1 number was measured on."""
    import torch

    from tuotuo_b200.synth import synthetic_token_stream

    D, L, K, V = 3000, 60, 8, 4000
    a = synthetic_token_stream(D, L, K, V, device="cpu", seed=11)
    b = synthetic_token_stream(D, L, K, V, device="cpu", seed=11, doc_seed=None)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    shards = [synthetic_token_stream(D, L, K, V, device="cpu", seed=11, doc_seed=100 + r)[0] for r in range(3)]
    assert not torch.equal(shards[0], shards[1]) and not torch.equal(shards[0], a[0])
    other_model = synthetic_token_stream(D, L, K, V, device="cpu", seed=12, doc_seed=100)[0]

    def profile(tok):  # corpus-wide word frequencies: a fingerprint of the topics, not of the documents
        return torch.bincount(tok.to(torch.int64), minlength=V).to(torch.float64) / tok.numel()

    same = float((profile(shards[0]) - profile(shards[1])).abs().sum())
    diff = float((profile(shards[0]) - profile(other_model)).abs().sum())
    assert same < 0.25 * diff  # shards of one model agree far better than two models do

    def distinct_per_doc(tok):
        t = tok.view(D, L).to(torch.int64)
        return float((torch.sort(t, 1).values.diff(dim=1) != 0).sum(1).add(1).double().mean())

    d = [distinct_per_doc(s) for s in shards]
    assert max(d) / min(d) < 1.01  # equal nonzeros per document = equal work per rank
