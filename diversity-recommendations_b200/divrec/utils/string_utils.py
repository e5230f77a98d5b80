"""Metric-name helper.  Same outputs as the reference's divrec/utils/string_utils.py:1-7 (which, despite
its name, produces snake_case): an underscore goes before every upper-case letter that is followed by
a lower-case one, except at position 0 and at the last character; everything is lower-cased.
  RecallAtKScore -> recall_atk_score,  PRI -> pri,  NDCGScore -> ndcg_score.
(A one-character name comes back doubled, as in the reference; class names never are.)"""
import re

# a word boundary sits before an upper-case letter followed by a lower-case one
_BOUNDARY = re.compile(r"(?=[A-Z][a-z])")


def to_camel_case(s: str) -> str:
    if len(s) == 1:
        return s.lower() * 2
    cuts = [m.start() for m in _BOUNDARY.finditer(s) if 0 < m.start() < len(s) - 1]
    words = [s[a:b] for a, b in zip([0] + cuts, cuts + [len(s)])]
    return "_".join(words).lower()
