#!/usr/bin/env python
"""Opcode histogram of the loops of one function in an object file (cuobjdump -sass): which pipe the hot loop loads.
usage: sass_loops.py <object> <substring of the mangled function name> [min_len]"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 150
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn, funcs = None, collections.OrderedDict()
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1)
        funcs[fn] = []
        continue
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and fn:
        funcs[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, ins in funcs.items():
    if pat not in fn:
        continue
    print(fn, len(ins))
    for a, t in ins:
        m = re.search(r"BRA\S*\s+.*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [x for x in ins if tgt <= x[0] <= a]
            if len(body) < min_len or len(body) > 2500:
                continue
            c = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", x[1]).split()[0] for x in body)
            heavy = sum(v for k, v in c.items() if k.startswith("IMAD"))
            print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr, IMAD* {heavy}, of which MOV {c.get('IMAD.MOV.U32', 0)}")
            print("   ", c.most_common(14))
