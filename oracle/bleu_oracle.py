"""CPU oracle for the caption metric -- TEST INFRASTRUCTURE ONLY (see oracle/git_oracle.py header).

Restates ``calculate_bleu_score_corpus`` (/root/reference/src/metrics.py:42-68).  That function hands
its inputs to ``nltk.translate.bleu_score.corpus_bleu`` (nltk==3.8.1, /root/reference/requirements.txt:21;
nltk is NOT installed here, so its published algorithm is restated below) *without* tokenising them: the
loops at metrics.py:58-65 only rebind loop variables.  ``corpus_bleu`` therefore iterates the Python
``str`` objects element by element, i.e. it computes CHARACTER-level BLEU-4.
The product's ``metrics.calculate_bleu_score_corpus`` must reproduce exactly this, quirk included.

PARITY STATUS: unpinned by the reference (its only test calls a function that does not exist,
/root/reference/tests/test_metrics.py:22).  Anchors: the nltk 3.8.1 algorithm + hand-computed cases in
tests/test_metrics.py (this repo).
"""
from __future__ import annotations

import math
import sys
from collections import Counter
from fractions import Fraction
from typing import List, Sequence


def _ngrams(seq: Sequence, n: int):
    return [tuple(seq[i:i + n]) for i in range(len(seq) - n + 1)]


def modified_precision(references, hypothesis, n):
    """nltk.translate.bleu_score.modified_precision (clipped n-gram counts)."""
    counts = Counter(_ngrams(hypothesis, n)) if len(hypothesis) >= n else Counter()
    max_counts = {}
    for reference in references:
        reference_counts = Counter(_ngrams(reference, n)) if len(reference) >= n else Counter()
        for ngram in counts:
            max_counts[ngram] = max(max_counts.get(ngram, 0), reference_counts[ngram])
    clipped = {ng: min(c, max_counts[ng]) for ng, c in counts.items()}
    return sum(clipped.values()), max(1, sum(counts.values()))


def closest_ref_length(references, hyp_len):
    ref_lens = (len(r) for r in references)
    return min(ref_lens, key=lambda ref_len: (abs(ref_len - hyp_len), ref_len))


def brevity_penalty(closest_ref_len, hyp_len):
    if hyp_len > closest_ref_len:
        return 1
    elif hyp_len == 0:
        return 0
    return math.exp(1 - closest_ref_len / hyp_len)


def corpus_bleu(list_of_references, hypotheses, weights=(0.25, 0.25, 0.25, 0.25)) -> float:
    """nltk 3.8.1 corpus_bleu with the default SmoothingFunction().method0 and no auto_reweigh."""
    p_num, p_den = Counter(), Counter()
    hyp_lengths, ref_lengths = 0, 0
    assert len(list_of_references) == len(hypotheses)
    for references, hypothesis in zip(list_of_references, hypotheses):
        for i in range(1, len(weights) + 1):
            num, den = modified_precision(references, hypothesis, i)
            p_num[i] += num
            p_den[i] += den
        hyp_len = len(hypothesis)
        hyp_lengths += hyp_len
        ref_lengths += closest_ref_length(references, hyp_len)
    bp = brevity_penalty(ref_lengths, hyp_lengths)
    if p_num[1] == 0:
        return 0
    # method0: zero-count precisions are replaced by the smallest positive float
    p_n = [Fraction(p_num[i], p_den[i]) if p_num[i] != 0 else sys.float_info.min for i in range(1, len(weights) + 1)]
    s = (w * math.log(p) for w, p in zip(weights, p_n) if p > 0)
    return bp * math.exp(math.fsum(s))


def calculate_bleu_score_corpus(references: List[List[str]], candidates: List[str]) -> float:
    """metrics.py:42-68.  References / candidates are passed through UNtokenised (character n-grams)."""
    assert len(references) == len(candidates), "The lengths of references and candidates must be the same"
    assert isinstance(references, list), "References must be a list as it is looking at multiple captions"
    assert isinstance(candidates, list), "Candidates must be a list as it is looking at multiple captions"
    return corpus_bleu(references, candidates) * 100
