#!/usr/bin/env python
"""Checks every `file.py:LINE[-LINE]` citation of the reference in the docs, headers and sources: the file must exist under
/root/reference (searched by suffix) and have that many lines.  Usage: python tools/check_citations.py  (exit 1 on a miss)"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PAT = re.compile(r"(?<![\w/.])((?:[\w.-]+/)*[\w.-]+\.(?:py|rst|json|toml|cfg|txt)|Pipfile):(\d+)(?:-(\d+))?")
ALIAS = {"overrides.py": "network_simple_overrides.py"}       # the docs' shorthand
OWN = ("oracle/", "tests/", "itsim_b200/", "tools/", "bench.py", "profiles/", "__graft_entry__.py")


def reference_files():
    out = {}
    for d, _, files in os.walk(REF):
        if "/.git" in d:
            continue
        for f in files:
            p = os.path.join(d, f)
            out[os.path.relpath(p, REF)] = p
    return out


def main():
    if not os.path.isdir(REF):
        print("reference tree absent: nothing checked")
        return 0
    ref = reference_files()
    lines = {}
    scan = []
    for d, _, files in os.walk(ROOT):
        if any(part in d for part in ("/.git", "/gpurun_out", "/tests/golden", "__pycache__", "/variants", "/baseline")):
            continue
        for f in files:
            if f.endswith((".md", ".h", ".cuh", ".cu", ".py", ".cpp", ".sh")) and f not in ("SURVEY.md", "VERDICT.md", "ADVICE.md", "PAPERS.md", "SNIPPETS.md", "BASELINE.md"):
                scan.append(os.path.join(d, f))
    bad = 0
    checked = 0
    for path in sorted(scan):
        text = open(path, errors="replace").read()
        for m in PAT.finditer(text):
            name, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            name = ALIAS.get(name, name)
            if name.startswith(OWN) or os.path.exists(os.path.join(ROOT, name)):
                continue                      # a citation of this repo's own file
            cands = [r for r in ref if r == name or r.endswith("/" + name)]
            if not cands:
                own = [1 for d, _, fs in os.walk(ROOT) if "/.git" not in d for f in fs if f == os.path.basename(name)]
                if own:
                    continue                  # a file of this repo cited by its base name
                print(f"{os.path.relpath(path, ROOT)}: {m.group(0)}: no such file in the reference")
                bad += 1
                continue
            ok = False
            for c in cands:
                if c not in lines:
                    lines[c] = sum(1 for _ in open(ref[c], errors="replace"))
                if max(a, b) <= lines[c]:
                    ok = True
            checked += 1
            if not ok:
                print(f"{os.path.relpath(path, ROOT)}: {m.group(0)}: {cands[0]} has {lines[cands[0]]} lines")
                bad += 1
    print(f"{checked} citations checked, {bad} bad")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
