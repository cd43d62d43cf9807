#!/usr/bin/env python
"""Mnemonic counts and excerpts of the hot kernels' SASS (no GPU needed): python tools/sass_excerpt.py > profiles/rNN_sass.txt"""
import collections, re, subprocess
so = "batrack_b200/libbatrack_ba.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WANT = [("k_edge_pass_v2ILb0ELi2", "edge pass (pose + depth), two tracks per lane"), ("k_edge_pass_v2ILb0ELi1", "edge pass (pose + depth)"),
        ("k_backsub", "back-substitution"), ("7k_schurE", "SIMT Schur"), ("k_schur_tc", "tcgen05 Schur"),
        ("k_solve_band_diag", "band Cholesky (FP64 DMMA), diagonal ownership"), ("k_solve_tiles", "shared-memory tile solver")]
MN = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "SYNCS", "LDGSTS", "UBLKCP", "DMMA", "REDG", "BAR.SYNC", "UCGABAR"]
print(f"# cuobjdump -sass {so} (sm_100a): mnemonic counts of the hot kernels and excerpts")
print("# UTCHMMA = tcgen05.mma kind::tf32 (5th-gen tensor core, TMEM accumulator), UTCBAR = tcgen05.commit -> mbarrier, LDTM = tcgen05.ld,")
print("# SYNCS = mbarrier ops, LDGSTS = cp.async, UBLKCP = cp.async.bulk, DMMA = mma.sync m8n8k4 f64, REDG = red.global (no return), UCGABAR = cluster barrier")
funcs = re.split(r"\n\s*Function : ", txt)
for key, label in WANT:
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if key not in name:
            continue
        ins = [l for l in f.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
        cnt = collections.Counter()
        for l in ins:
            for m in MN:
                if re.search(r"\b" + re.escape(m), l):
                    cnt[m] += 1
        print(f"\n== {name}  [{label}]  {len(ins)} instructions")
        print("   " + "  ".join(f"{m} {cnt[m]}" for m in MN if cnt[m]))
        shown = 0
        for l in ins:
            if any(re.search(r"\b" + m, l) for m in ("UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "DMMA", "REDG")) and shown < 8:
                print("   " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l).strip())
                shown += 1
