"""``doc_generator`` — drop-in for tuotuo/generator.py:10-289 (API surface only; not on the hot path).

Un-smoothed-LDA toy generator over the reference's fixed 5 topics x 40 words table: beta (5 x 40, double,
row-normalised, generator.py:240-250), theta ~ Dirichlet(topic_prior) per document (:252-253), and per token
z ~ Cat(theta_d), w ~ Cat(beta_z) (:260-278).  pyro is not installable here; ``torch.distributions`` gives the
same distributions.  Sampling is vectorised (one categorical draw per token position for all documents) instead
of the reference's M*L scalar draws; like the reference, the RNG is whatever ``torch``'s global generator holds.
"""
from __future__ import annotations

from typing import Dict

import torch as tr

DTYPE = tr.double

TOPICS = {0: "science", 1: "sport", 2: "art", 3: "health", 4: "law"}

# (name, home topic, P(word | topic k) before row normalisation for k = science, sport, art, health, law)
_WORD_TABLE = (
    ("quantum", "science", (0.96, 0.01, 0.01, 0.01, 0.01)),
    ("genetics", "science", (0.5, 0.04, 0.04, 0.4, 0.02)),
    ("research", "science", (0.7, 0.04, 0.03, 0.2, 0.03)),
    ("scientst", "science", (0.9, 0.025, 0.025, 0.025, 0.025)),
    ("energy", "science", (0.4, 0.3, 0.1, 0.1, 0.1)),
    ("electricity", "science", (0.7, 0.1, 0.05, 0.1, 0.05)),
    ("immunology", "science", (0.25, 0.2, 0.2, 0.25, 0.1)),
    ("astrophysics", "science", (0.9, 0.025, 0.025, 0.025, 0.025)),
    ("athletics", "sport", (0.05, 0.8, 0.05, 0.05, 0.05)),
    ("exercise", "sport", (0.04, 0.5, 0.03, 0.4, 0.03)),
    ("physical", "sport", (0.3, 0.5, 0.15, 0.025, 0.025)),
    ("football", "sport", (0.05, 0.9, 0.01, 0.01, 0.03)),
    ("game", "sport", (0.1, 0.4, 0.3, 0.1, 0.1)),
    ("Olympic", "sport", (0.03, 0.9, 0.03, 0.02, 0.02)),
    ("recreation", "sport", (0.01, 0.8, 0.1, 0.04, 0.04)),
    ("FIFA", "sport", (0.01, 0.95, 0.01, 0.01, 0.02)),
    ("Technique", "art", (0.1, 0.2, 0.35, 0.2, 0.15)),
    ("content", "art", (0.1, 0.05, 0.7, 0.05, 0.1)),
    ("Craftsmanship", "art", (0.2, 0.05, 0.7, 0.025, 0.025)),
    ("form", "art", (0.2, 0.1, 0.4, 0.1, 0.2)),
    ("Symmetrical", "art", (0.15, 0.15, 0.6, 0.15, 0.15)),
    ("asymmetrical", "art", (0.15, 0.15, 0.6, 0.15, 0.15)),
    ("picture", "art", (0.035, 0.035, 0.9, 0.015, 0.015)),
    ("concert", "art", (0.04, 0.04, 0.9, 0.01, 0.01)),
    ("allergy", "health", (0.3, 0.04, 0.04, 0.6, 0.02)),
    ("appetite", "health", (0.05, 0.2, 0.01, 0.7, 0.04)),
    ("fever", "health", (0.1, 0.1, 0.1, 0.5, 0.2)),
    ("infection", "health", (0.3, 0.25, 0.05, 0.35, 0.05)),
    ("contagious", "health", (0.15, 0.05, 0.05, 0.7, 0.05)),
    ("bruise", "health", (0.1, 0.2, 0.05, 0.6, 0.05)),
    ("decongestant", "health", (0.025, 0.025, 0.025, 0.9, 0.025)),
    ("injection", "health", (0.03, 0.03, 0.1, 0.8, 0.04)),
    ("contract", "law", (0.025, 0.025, 0.025, 0.025, 0.9)),
    ("bankrupt", "law", (0.01, 0.01, 0.01, 0.01, 0.96)),
    ("evidence", "law", (0.3, 0.03, 0.03, 0.04, 0.6)),
    ("court", "law", (0.025, 0.025, 0.025, 0.025, 0.9)),
    ("attorney", "law", (0.05, 0.04, 0.03, 0.03, 0.85)),
    ("copyright", "law", (0.1, 0.1, 0.1, 0.1, 0.6)),
    ("accuse", "law", (0.025, 0.025, 0.025, 0.025, 0.9)),
    ("divorce", "law", (0.025, 0.025, 0.025, 0.025, 0.9)),
)


class doc_generator:
    """Document generator based on the un-smoothed version of LDA (same attributes as the reference's class)."""

    def __init__(self, M: int, L: int, topic_prior: tr.Tensor, beta=None, eta=None) -> None:
        self.topics = dict(TOPICS)
        self.words = {i: {"name": n, "prob": list(p), "topic": t} for i, (n, t, p) in enumerate(_WORD_TABLE)}
        self.L = L
        self.K = 5
        assert self.K == len(topic_prior)
        self.M = M
        self.V = 40
        cols = tr.tensor([w["prob"] for w in self.words.values()], dtype=DTYPE)  # (V, K)
        self.beta = cols.T.contiguous()
        self.beta = self.beta / self.beta.sum(dim=1, keepdim=True)  # generator.py:247-248
        assert (tr.round(self.beta.sum(dim=1)) == tr.ones(self.K, dtype=DTYPE)).all()
        self.alpha = tr.distributions.Dirichlet(topic_prior)
        self.theta = self.alpha.sample((self.M,))  # generator.py:253

    def generate_doc(self, verbose: bool = False) -> Dict[int, str]:
        z = tr.distributions.Categorical(self.theta).sample((self.L,)).T  # (M, L) topic of every token
        w = tr.distributions.Categorical(self.beta[z]).sample()  # (M, L) word of every token
        docs = {}
        for d in range(self.M):
            names = [self.words[int(i)]["name"] for i in w[d]]
            if verbose:
                for n, (zi, nm) in enumerate(zip(z[d], names)):
                    print(f"Document: {d} | word: {n} -> topic: {self.topics[int(zi)]} -> word: {nm}")
            doc_string = " ".join(names)
            if verbose:
                print(f"Document {d}: {doc_string}")
                print()
            docs[d] = doc_string
        self.docs = docs
        return docs
