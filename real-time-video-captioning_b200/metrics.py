"""Caption metric of the reference path: ``calculate_bleu_score_corpus`` (/root/reference/src/metrics.py:42-68).

Host-side string arithmetic (not a GPU kernel).  The reference forwards its arguments to nltk's
``corpus_bleu`` WITHOUT tokenising them -- the loops at metrics.py:58-65 only rebind their loop variables --
so nltk iterates the Python strings character by character and the returned number is a character-level
BLEU-4 (x100).  Exact metric parity requires reproducing that, so this implementation counts n-grams over
whatever sequence elements it is given (characters for str inputs), with nltk 3.8.1's conventions: clipped
counts summed over the corpus, closest reference length with ties to the shorter one, brevity penalty,
uniform 4-gram weights, ``SmoothingFunction().method0`` (zero precisions replaced by the smallest positive
float), and 0 when no unigram matches.
"""
from __future__ import annotations

import math
import sys
from collections import Counter
from typing import List


def _count(seq, n: int) -> Counter:
    return Counter(tuple(seq[i:i + n]) for i in range(len(seq) - n + 1)) if len(seq) >= n else Counter()


def _bleu(refs_per_item, hyps, max_n: int = 4) -> float:
    num = [0] * (max_n + 1)
    den = [0] * (max_n + 1)
    hyp_total = ref_total = 0
    for refs, hyp in zip(refs_per_item, hyps):
        for n in range(1, max_n + 1):
            h = _count(hyp, n)
            best = Counter()
            for r in refs:
                rc = _count(r, n)
                for g in h:
                    best[g] = max(best[g], rc[g])
            num[n] += sum(min(c, best[g]) for g, c in h.items())
            den[n] += max(1, sum(h.values()))
        hyp_total += len(hyp)
        ref_total += min((len(r) for r in refs), key=lambda rl: (abs(rl - len(hyp)), rl))
    if num[1] == 0:
        return 0
    if hyp_total > ref_total:
        bp = 1.0
    elif hyp_total == 0:
        bp = 0.0
    else:
        bp = math.exp(1 - ref_total / hyp_total)
    logs = []
    for n in range(1, max_n + 1):
        p = num[n] / den[n] if num[n] else sys.float_info.min
        logs.append(0.25 * math.log(p))
    return bp * math.exp(math.fsum(logs))


def calculate_bleu_score_corpus(references: List[List[str]], candidates: List[str]) -> float:
    """Same signature, assertions and (character-level) result as the reference function."""
    assert len(references) == len(candidates), "The lengths of references and candidates must be the same"
    assert isinstance(references, list), "References must be a list as it is looking at multiple captions"
    assert isinstance(candidates, list), "Candidates must be a list as it is looking at multiple captions"
    return _bleu(references, candidates) * 100
