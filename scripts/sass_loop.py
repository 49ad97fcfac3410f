#!/usr/bin/env python
"""Static op mix of a kernel's hottest loop from its SASS (no GPU needed).

  python scripts/sass_loop.py [--fn SUBSTR] [--px N] [-- extra nvcc flags]

Compiles realsensetracker_b200/csrc/rst_icp_part0_iter_plain.cu to a cubin (sm_100a), dumps the SASS of the first
function whose demangled name contains SUBSTR (default: the plain level-0 ICP kernel), takes the backward-branch
region with the most FFMA/FFMA2/LDGSTS as "the loop" and prints instructions per pixel by opcode (--px = pixels one
thread handles per loop trip; k_icp_iter: 2 stages x 2*CPW pixels = 8).
"""
import argparse
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fn", default="k_icp_iter<(int)0, (bool)0, (bool)0, (bool)0, (bool)0>")
    ap.add_argument("--px", type=float, default=8.0)
    ap.add_argument("--src", default="realsensetracker_b200/csrc/rst_icp_part0_iter_plain.cu")
    ap.add_argument("--dump", default=None, help="write the function's SASS here")
    ap.add_argument("rest", nargs="*")
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as td:
        cubin = Path(td) / "k.cubin"
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-cubin",
               "-Xptxas", "-v", "-I", str(ROOT / "include"), "-ccbin", "/usr/bin/g++", *a.rest, "-o", str(cubin), str(ROOT / a.src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stderr)
        sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
        dem = subprocess.run(["cu++filt"], input=sass, capture_output=True, text=True).stdout
        # ptxas -v block of the function
        log = subprocess.run(["cu++filt"], input=r.stderr, capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", dem)
    blk = next((b for b in blocks[1:] if a.fn in b.splitlines()[0]), None)
    if blk is None:
        sys.exit("function not found; have:\n" + "\n".join(b.splitlines()[0] for b in blocks[1:]))
    print("function:", blk.splitlines()[0])
    lines = log.splitlines()
    for i, ln in enumerate(lines):
        if "Compiling entry function" in ln and a.fn in ln:
            print("  ", " | ".join(x.strip().replace("ptxas info    : ", "") for x in lines[i + 1:i + 4]))
            break
    if a.dump:
        Path(a.dump).write_text(blk)
    ins = []
    for ln in blk.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {ad: i for i, (ad, _) in enumerate(ins)}
    best = None
    for i, (ad, tx) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", tx)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < ad and tgt in addr:
                span = i - addr[tgt] + 1
                n_g = sum(1 for _, t in ins[addr[tgt]:i + 1] if "LDGSTS.E.128" in t or "LDGSTS.E.BYPASS.128" in t)
                score = -span if n_g >= 4 else -10 ** 9   # the innermost loop that holds the texel gathers
                if best is None or score > best[3]:
                    best = (span, addr[tgt], i, score)
    if best is None:
        sys.exit("no backward branch")
    span, i0, i1, _ = best
    ops = collections.Counter()
    for _, tx in ins[i0:i1 + 1]:
        t = tx.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    print(f"loop: {span} instructions ({ins[i0][0]:#x}..{ins[i1][0]:#x}) = {span / a.px:.1f} per pixel; whole function {len(ins)}")
    for op, n in ops.most_common():
        print(f"  {op:10s} {n / a.px:6.2f}")


if __name__ == "__main__":
    main()
