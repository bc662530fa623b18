#!/usr/bin/env python
"""Static instruction mix of the loops in a cuobjdump -sass listing (finds backward branches)."""
import collections
import re
import sys


def parse(path):
    ins = []
    for ln in open(path):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            addr = int(m.group(1), 16)
            txt = m.group(2).strip()
            op = re.sub(r"^@!?U?P\d+\s+", "", txt).split()[0].split(".")[0]
            ins.append((addr, op, txt))
    return ins


def main():
    ins = parse(sys.argv[1])
    loops = []
    for addr, op, txt in ins:
        if op == "BRA":
            m = re.search(r"0x([0-9a-f]+)", txt)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
    FP64 = {"DFMA", "DMUL", "DADD", "DSETP"}
    for lo, hi in sorted(loops, key=lambda t: t[1] - t[0]):
        body = [i for i in ins if lo <= i[0] <= hi]
        cnt = collections.Counter(op for _, op, _ in body)
        fp = sum(cnt[k] for k in FP64)
        if len(body) < 50:
            continue
        top = ", ".join(f"{k} {v}" for k, v in cnt.most_common(14))
        print(f"loop 0x{lo:x}-0x{hi:x}: {len(body)} instr, FP64 {fp} ({100*fp/len(body):.0f} %), MUFU {cnt['MUFU']}\n   {top}")


if __name__ == "__main__":
    main()
