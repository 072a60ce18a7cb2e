#!/usr/bin/env python
"""SASS of one kernel of libqforest_b200.so (cuobjdump), optionally cut to the lines between two
mnemonic patterns:  python tools/sass_excerpt.py <mangled-name-substring> [first-regex] [last-regex]"""
import re
import subprocess
import sys


def main():
    pat = sys.argv[1]
    first = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    last = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["cuobjdump", "-sass", "scikit-garden_b200/libqforest_b200.so"], capture_output=True, text=True).stdout
    take, lines = False, []
    for ln in out.splitlines():
        if "Function :" in ln:
            take = pat in ln
            if take:
                lines.append(ln.strip())
            continue
        if take and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
    if first:
        idx = [i for i, l in enumerate(lines) if first.search(l)]
        end = [i for i, l in enumerate(lines) if last.search(l)] if last else []
        if idx:
            a = max(1, idx[0] - 12)
            b = (max(e for e in end if e >= idx[0]) + 4) if any(e >= idx[0] for e in end) else min(len(lines), idx[0] + 120)
            lines = lines[:1] + lines[a:b]
    print("\n".join(lines))


if __name__ == "__main__":
    main()
