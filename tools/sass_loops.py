"""Summarise the loops (backward branches) of a cuobjdump -sass listing: address range, instruction mix.

    cuobjdump -sass -fun <mangled> lib.so > k.sass; python tools/sass_loops.py k.sass [min_len]
"""
import collections
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            body = ins[addr_index[tgt]: i + 1]
            if len(body) < min_len:
                continue
            mix = collections.Counter()
            for _, bt in body:
                op = bt.split()[1] if bt.startswith("@") else bt.split()[0]
                mix[op.split(".")[0]] += 1
            print("loop 0x%x..0x%x  %d instr: %s" % (tgt, a, len(body), ", ".join("%s %d" % kv for kv in mix.most_common(14))))
