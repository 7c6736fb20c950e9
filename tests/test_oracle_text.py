"""The text oracle (oracle/text_ref.{c,py}) pinned against golden outputs of the
unmodified reference functions (tools.py:51-139,326-493)."""
import re

import pytest

from oracle import text_ref as T


def test_whitespace_class_matches_python_re():
    ws = {cp for cp in range(0x110000) if re.match(r"\s", chr(cp))}
    assert ws == T._WS


def test_levenshtein_golden(text_golden):
    for c in text_golden["levenshtein"]:
        assert T.levenshtein(c["a"], c["b"]) == c["d"]
        if len(c["a"]) * len(c["b"]) < 40_000:
            assert T.levenshtein_py(c["a"], c["b"]) == c["d"]
    for c in text_golden["levenshtein_words"]:
        assert T.levenshtein_words(c["a"], c["b"]) == c["d"]


def test_normalize_golden(text_golden):
    for c in text_golden["normalize_text"]:
        assert T.normalize_text(c["in"]) == c["out"]
        assert T.normalize_text(c["in"], True) == c["out_lower"]


def test_compare_versions_golden(text_golden):
    for c in text_golden["compare_versions"]:
        assert T.compare_versions(c["v1"], c["v2"]) == c["out"]


def test_merge_versions_golden(text_golden):
    for c in text_golden["merge_versions"]:
        assert T.merge_versions(c["versions"]) == c["out"]


def test_tier1_golden(text_golden):
    for c in text_golden["tier1_metrics"]:
        assert T.tier1_metrics(c["gt"], c["ocr"], c["lower"]) == c["out"]
