"""The product package must never import, call or load anything under oracle/ (or a CPU fallback):
the oracle is test infrastructure."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_does_not_touch_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "spuq_b200")):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            text = open(os.path.join(base, f), errors="ignore").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "oracle/" in text.replace("# oracle/", ""):
                bad.append(os.path.join(base, f))
    assert not bad, bad
