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


def test_oracle_vs_unmodified_reference_random(capsys):
    """Differential check against the reference's own functions (imported unmodified, only `ollama` stubbed) on random
    strings with Unicode whitespace, curly quotes, dashes and non-Latin characters.  Runs where /root/reference exists."""
    import os
    import random
    import sys
    import types
    if not os.path.isdir("/root/reference/ocr_agent"):
        pytest.skip("reference tree not present")
    from oracle import text_ref as T
    saved = {k: v for k, v in sys.modules.items() if k == "ocr_agent" or k.startswith("ocr_agent.") or k == "ollama"}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, "/root/reference")
    sys.modules["ollama"] = types.ModuleType("ollama")
    try:
        from ocr_agent import tools as ref
        rnd = random.Random(5)
        alpha = "abcde fgh\n\t ijk.,;'’“”–—   AB Cé中"

        def rs(n):
            return "".join(rnd.choice(alpha) for _ in range(n))

        def mutate(s):
            s = list(s)
            for _ in range(rnd.randint(0, max(1, len(s) // 6))):
                if not s:
                    break
                i, op = rnd.randrange(len(s)), rnd.random()
                if op < 0.3:
                    del s[i]
                elif op < 0.6:
                    s.insert(i, rnd.choice(alpha))
                else:
                    s[i] = rnd.choice(alpha)
            return "".join(s)

        for it in range(300):
            a = rs(rnd.randint(0, 80))
            b = mutate(a) if it % 3 else rs(rnd.randint(0, 60))
            c = mutate(a)
            for lower in (False, True):
                assert T.normalize_text(a, lower) == ref.normalize_text(a, lower)
                if a.strip():
                    assert T.tier1_metrics(a, b, lower) == ref.tier1_metrics(a, b, lower=lower), it
            assert T.levenshtein(a, b) == ref.levenshtein(a, b)
            assert T.compare_versions(a, b) == ref.compare_versions(a, b), it
            for vs in ([a, b, c], [a, b], [a], [], [b, a, c, a]):
                assert T.merge_versions(list(vs)) == ref.merge_versions(list(vs)), (it, len(vs))
            assert T.evaluate(b, a) == ref.evaluate(b, a)
            assert T.evaluate(b) == ref.evaluate(b) == {}
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "ocr_agent" or k.startswith("ocr_agent.") or k == "ollama"]:
            del sys.modules[k]
        sys.modules.update(saved)
        capsys.readouterr()
