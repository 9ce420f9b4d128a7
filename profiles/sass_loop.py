#!/usr/bin/env python
"""Instruction mix of the hot loop of each kernel in a cubin / .so (cuobjdump -sass): finds the largest backward
branch per function and histograms the opcodes inside it.  Used to tune the issue-bound stencil kernels without a GPU.
    python profiles/sass_loop.py <file.o|.so> [name-filter]"""
import collections
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if flt and flt not in name:
            continue
        ins = []
        for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f):
            ins.append((int(m.group(1), 16), m.group(2).strip()))
        loops = []
        for addr, text in ins:
            mm = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", text)
            if mm:
                tgt = int(mm.group(1), 16)
                if tgt < addr and (addr - tgt) // 16 >= 100:
                    loops.append((tgt, addr))
        print(f"{name}  (total {len(ins)} instr)")
        for lo, hi in loops:
            body = [t for a, t in ins if lo <= a <= hi]
            hist = collections.Counter()
            for t in body:
                t = re.sub(r"^@!?U?P\d+\s+", "", t)
                op = t.split()[0].split(".")[0]
                hist[op] += 1
            fp = hist["DFMA"] + hist["DADD"] + hist["DMUL"]
            print(f"  loop @{lo:#x}: {len(body)} instr, fp64 {fp}: " + ", ".join(f"{k} {v}" for k, v in hist.most_common(22)))


if __name__ == "__main__":
    main()
