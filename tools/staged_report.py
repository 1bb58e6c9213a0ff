#!/usr/bin/env python
"""Summarise a hardware run of the staged GPU tests (round-2 helper).

Input: the files tools/round2_first_call.sh writes (gpurun_out/tests_*.txt = tails of `pytest -rA --runxfail` runs, one per test
file).  Output: per test function, how many parametrisations passed / failed, so that the functions whose every case passed on
the B200 can lose their `xfail` marker and the failing ones get looked at first.

    python tools/staged_report.py gpurun_out/tests_*.txt
"""
import re
import sys
from collections import defaultdict

LINE = re.compile(r"^(PASSED|FAILED|ERROR|SKIPPED|XFAIL|XPASS)\s+(\S+?)(?:\s+-\s+(.*))?$")


def main(paths):
    stat = defaultdict(lambda: defaultdict(int))
    why = {}
    for p in paths:
        for ln in open(p, errors="replace"):
            m = LINE.match(ln.strip())
            if not m:
                continue
            outcome, nodeid, msg = m.groups()
            fn = nodeid.split("[")[0]
            stat[fn][outcome] += 1
            if outcome in ("FAILED", "ERROR") and fn not in why and msg:
                why[fn] = msg[:160]
    if not stat:
        print("no pytest -rA result lines found")
        return 1
    clean, dirty = [], []
    for fn, s in sorted(stat.items()):
        bad = s["FAILED"] + s["ERROR"] + s["XFAIL"]
        (dirty if bad else clean).append((fn, dict(s)))
    print(f"{len(clean)} test functions passed in every parametrisation (candidates for removing the xfail marker):")
    for fn, s in clean:
        print(f"  ok   {fn}  {s}")
    print(f"\n{len(dirty)} test functions with failures:")
    for fn, s in dirty:
        print(f"  FAIL {fn}  {s}  {why.get(fn, '')}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
