#!/usr/bin/env python
"""Share of a file's substantive lines that occur verbatim (whitespace-normalised) anywhere in the reference's
.py/.cu sources -- the spread-copy check VERDICT r01 ran by hand.  Build-container tool (needs /root/reference)."""
import os
import re
import sys

REF = os.environ.get("MISO_REFERENCE_ROOT", "/root/reference")


def norm(line):
    return re.sub(r"\s+", "", line.split("#")[0] if not line.strip().startswith('"""') else line)


def substantive(line):
    s = norm(line)
    return len(s) >= 12 and not s.startswith(("import", "from", '"""', "'''", "@", "else:", "return", "pass"))


def main(paths):
    ref = set()
    for root, _, files in os.walk(REF):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                try:
                    for ln in open(os.path.join(root, f), errors="ignore"):
                        if substantive(ln):
                            ref.add(norm(ln))
                except OSError:
                    pass
    for p in paths:
        lines = [ln for ln in open(p, errors="ignore") if substantive(ln)]
        hit = sum(norm(ln) in ref for ln in lines)
        print(f"{p}: {hit}/{len(lines)} = {100.0 * hit / max(len(lines), 1):.1f}%")


if __name__ == "__main__":
    main(sys.argv[1:])
