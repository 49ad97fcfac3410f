#!/usr/bin/env python
"""Counts the Blackwell-specific SASS opcodes per kernel of the shipped librst_align.so (no GPU needed):
UTMALDG (TMA tensor copy), UBLKCP (bulk async copy), SYNCS (mbarrier), FFMA2/FMUL2/FADD2 (packed fp32),
LDGSTS (cp.async), UCGABAR / cluster barriers, ACQBULK, tensor-core opcodes (expected: none)."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "realsensetracker_b200/_lib/librst_align.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = subprocess.run(["cu++filt"], input=sass, capture_output=True, text=True).stdout
KEYS = ["UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "LDGSTS", "UCGABAR", "CGAERRBAR", "ACQBULK", "HMMA", "UTCMMA", "QGMMA", "IMMA"]
print("arch:", sorted(set(re.findall(r"arch = (sm_\w+)", dem))))
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in KEYS))
tot = collections.Counter()
for blk in re.split(r"\n\s*Function : ", dem)[1:]:
    name = blk.splitlines()[0].strip()
    ops = collections.Counter()
    for ln in blk.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            for k in KEYS:
                if m.group(1).startswith(k): ops[k] += 1
    row = [ops.get(k, 0) for k in KEYS]
    tot.update({k: ops.get(k, 0) for k in KEYS})
    if any(row) and ("k_icp_iter<(int)0, (bool)0, (bool)0, (bool)0, (bool)0>" in name or "k_icp_fused<(int)0, (bool)0, (bool)0>" in name
                     or "k_preprocess" in name or "k_icp3d" in name or "k_icp_iter<(int)1, (bool)0, (bool)0, (bool)1" in name):
        short = re.sub(r"\(anonymous namespace\)::|rst::", "", name)[:70]
        print(f"{short:70s} " + " ".join(f"{v:8d}" for v in row))
print(f"{'ALL KERNELS of the library':70s} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
