#!/usr/bin/env python
"""Summarise the SASS of our kernels: per kernel, a histogram of the mnemonics that prove the
Blackwell-native path (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, LDGMC = multimem.ld_reduce,
LDGSTS = cp.async, multimem = NVLS).  Usage: tools/sass_summary.py <file.sass> > profiles/sass/x.md"""
import collections
import re
import sys

KEY = re.compile(r"\b(LDGMC|REDGMC|UTC[A-Z]*MMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|LDGSTS|SYNCS|HMMA|FFMA|FMUL|FADD|LDG|STG|LDS|STS|BAR|ATOM|RED|MEMBAR|FENCE|ERRBAR|CCTL|MULTIMEM|LD\.E|ST\.E|SHFL|MUFU|ELECT|R2UR|NANOSLEEP)\b")
fn = None
hist = collections.OrderedDict()
for line in open(sys.argv[1]):
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        hist[fn] = collections.Counter()
        continue
    if fn and "/*" in line and ";" in line:
        body = line.split("*/", 1)[-1]
        op = body.strip().split()
        if not op:
            continue
        mn = op[0] if not op[0].startswith("@") else (op[1] if len(op) > 1 else "")
        base = mn.split(".")[0]
        hist[fn][base] += 1
        if "multimem" in line.lower() or ".MULTIMEM" in mn:
            hist[fn]["<multimem>"] += 1
print("| kernel | instructions | notable mnemonics |\n|---|---|---|")
for fn, h in hist.items():
    tot = sum(h.values())
    keys = {k: v for k, v in h.items() if KEY.search(k) or k.startswith("UT") or k.startswith("LDTM") or k == "<multimem>"}
    try:
        import subprocess

        short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
    except OSError:
        short = fn
    short = re.sub(r"\(anonymous namespace\)::", "", short)
    short = re.sub(r"\(.*$", "", short)[:150]   # drop the argument list
    print(f"| `{short}` | {tot} | " + ", ".join(f"{k}×{v}" for k, v in sorted(keys.items(), key=lambda kv: -kv[1])[:14]) + " |")
