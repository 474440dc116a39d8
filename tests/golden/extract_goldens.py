#!/usr/bin/env python3
"""Extract the golden vectors held by the reference's own unit tests into JSON.

Run in the build container only (needs /root/reference):
    python tests/golden/extract_goldens.py

It parses the `std::vector<T> name = {...};` initialisers of the test files listed in
SURVEY.md §8(c) and writes tests/golden/reference_goldens.json.  Only *data* (numbers) is
extracted; no reference code is copied.  The JSON travels to the GPU box, the reference
does not.
"""
import json
import os
import re
import sys

REF = os.environ.get("SHAMROCK_REFERENCE", "/root/reference")

FILES = [
    "src/tests/shamtree/MortonCodeSetTests.cpp",
    "src/tests/shamtree/MortonReducedSetTests.cpp",
    "src/tests/shamtree/KarrasRadixTreeTests.cpp",
    "src/tests/shamtree/KarrasRadixTreeAABBTests.cpp",
    "src/tests/shamtree/KarrasRadixTreeFieldTests.cpp",
    "src/tests/shamtree/CLBVHObjectIteratorTests.cpp",
    "src/tests/shammodels/sph/modules/IterateSmoothingLengthDensityTests.cpp",
]

VEC_RE = re.compile(
    r"std::vector<\s*([A-Za-z0-9_:]+)\s*>\s+([A-Za-z0-9_]+)\s*(?:=\s*)?\{(.*?)\}\s*;", re.S
)
TEST_RE = re.compile(r'NEW_TEST\s*\(\s*\w+\s*,\s*"([^"]+)"')


def strip_comments(s):
    s = re.sub(r"//[^\n]*", "", s)
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return s


def parse_scalar(tok):
    tok = tok.strip()
    tok = re.sub(r"_u64$|_u32$|ULL$|UL$|U$|u$", "", tok)
    if tok.startswith("0b"):
        return int(tok[2:], 2)
    if re.fullmatch(r"-?\d+", tok):
        return int(tok)
    return float(tok)


def parse_body(body):
    body = strip_comments(body)
    out = []
    # vectors: Tvec(a, b, c) or Tvec{a,b,c}
    if re.search(r"Tvec\s*[\({]", body):
        for m in re.finditer(r"Tvec\s*[\({]([^\)}]*)[\)}]", body):
            out.append([float(parse_scalar(t)) for t in m.group(1).split(",")])
        return out
    if "{" in body:  # nested {a, b, c} triples
        for m in re.finditer(r"\{([^{}]*)\}", body):
            out.append([float(parse_scalar(t)) for t in m.group(1).split(",")])
        return out
    for tok in body.split(","):
        tok = tok.strip()
        if tok:
            out.append(parse_scalar(tok))
    return out


def main():
    res = {}
    for rel in FILES:
        path = os.path.join(REF, rel)
        src = open(path).read()
        # positions of the tests
        marks = [(m.start(), m.group(1)) for m in TEST_RE.finditer(src)]
        marks.append((len(src), None))
        fres = {}
        # file-level (before first test)
        blocks = [("<file>", src[: marks[0][0]])]
        for i in range(len(marks) - 1):
            blocks.append((marks[i][1], src[marks[i][0] : marks[i + 1][0]]))
        for name, blk in blocks:
            d = {}
            for m in VEC_RE.finditer(blk):
                typ, var, body = m.group(1), m.group(2), m.group(3)
                try:
                    val = parse_body(body)
                except ValueError:
                    continue
                key = var
                k = 1
                while key in d:  # same name twice in a test (scoped blocks): keep order
                    k += 1
                    key = f"{var}#{k}"
                d[key] = {"type": typ, "value": val}
            if d:
                fres[name] = d
        res[rel] = fres
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
    with open(out, "w") as f:
        json.dump(res, f, indent=0, separators=(",", ":"))
    n = sum(len(t) for fr in res.values() for t in fr.values())
    print(f"wrote {out}: {n} vectors from {len(FILES)} files")


if __name__ == "__main__":
    sys.exit(main())
