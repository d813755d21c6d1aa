#!/usr/bin/env python
"""Instruction mix of the hottest loop of one kernel, read from the built library's SASS (no GPU needed).

    python tools/sass_loop.py <mangled-name-substring> [--lib path] [--print]

The hottest loop is taken to be the backward branch spanning the most IMAD.WIDE / DFMA instructions. Prints the
per-opcode histogram, the totals by issue class used in DESIGN.md section 4.1 (FP64 / FMA-pipe integer / ALU / LSU /
MUFU / uniform datapath) and the dispatch-port model 2 x FP64 + 1 x other.
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("--lib", default=os.path.join(ROOT, "hedgehog.jl_b200", "libhedgehog_mc.so"))
    ap.add_argument("--print", action="store_true")
    ap.add_argument("--loop", type=int, default=None, help="index of the backward branch to analyse (default: heaviest)")
    args = ap.parse_args()
    names = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True).stdout
    funcs = re.findall(r"Function : (\S+)", names)
    match = [f for f in funcs if args.name in f]
    if not match:
        sys.exit(f"no kernel matching {args.name!r}")
    fn = match[0]
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, args.lib], capture_output=True, text=True).stdout
    ins = []
    for l in out.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for a, t in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t.split("BRA")[1])
            if m and int(m.group(1), 16) < a:
                lo = int(m.group(1), 16)
                body = [x for x in ins if lo <= x[0] <= a]
                w = sum(1 for _, x in body if re.search(r"IMAD\.WIDE|DFMA|DMUL|DADD", x))
                loops.append((w, lo, a, body))
    if not loops:
        sys.exit("no loop")
    # innermost loops only (no other backward branch strictly inside), heaviest first
    inner = [L for L in loops if not any(o is not L and L[1] <= o[1] and o[2] <= L[2] for o in loops)]
    w, lo, hi, body = max(inner) if args.loop is None else loops[args.loop]
    cnt = collections.Counter()
    cls = collections.Counter()
    for _, t in body:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        base = op.split(".")[0]
        cnt[op] += 1
        if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
            cls["fp64"] += 1
        elif base.startswith("U") and base not in ("UNPACK",):
            cls["uniform"] += 1
        elif base in ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2"):
            cls["fma_pipe"] += 1
        elif base in ("LDS", "LDG", "STS", "STG", "LDC", "LD", "ST"):
            cls["lsu"] += 1
        elif base == "MUFU":
            cls["mufu"] += 1
        elif base in ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "BAR"):
            cls["control"] += 1
        else:
            cls["alu"] += 1
        if args.print:
            print(t)
    n = len(body)
    print(f"{fn}\nloop 0x{lo:x}..0x{hi:x}: {n} instructions")
    print(dict(cnt.most_common()))
    print(dict(cls))
    other = n - cls["fp64"] - cls["uniform"]
    print(f"dispatch model: 2 x {cls['fp64']} FP64 + {other} other (uniform datapath excluded) = {2 * cls['fp64'] + other} cycles per warp-iteration")


if __name__ == "__main__":
    main()
