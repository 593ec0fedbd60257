#!/usr/bin/env python
"""cuobjdump -sass of the shipped library: per kernel, how often the SASS mnemonics that prove tcgen05 / TMEM / TMA /
cp.async / multimem / system-scope exchange traffic occur, and that no float atomic exists.
usage: python profiles/sass_markers.py > profiles/r2_sass_markers.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "neural_tensor_network_for_kb_completion_b200", "libntn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
PAT = collections.OrderedDict([
    ("UTCHMMA (tcgen05.mma)", r"\bUTCHMMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("UTMALDG (TMA tile load)", r"\bUTMALDG"),
    ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
    ("LDGMC (multimem.ld_reduce)", r"\bLDGMC"), ("ST*.STRONG.SYS (flags, peer / multicast stores)", r"\bSTG?\.E[.\w]*\.STRONG\.SYS"),
    ("LD*.STRONG.SYS (flag polls, peer loads)", r"\bLDG?\.E[.\w]*\.STRONG\.SYS"),
    ("float atomics (ATOM/RED .F32)", r"\b(ATOM|RED|ATOMG|REDG)\.[A-Z0-9.]*F32")])
cur, cnt = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    if cur:
        for k, p in PAT.items():
            if re.search(p, line):
                cnt[cur][k] += 1
dem = subprocess.run(["cu++filt"] + list(cnt), capture_output=True, text=True).stdout.splitlines()
print("cuobjdump -sass libntn_b200.so (sm_100a): occurrences per kernel; kernels of the library's own sources only (CUB's are\n"
      "omitted), kernels without any marker omitted.  multimem.st compiles to an STG...STRONG.SYS on the multicast alias.\n")
tot = collections.Counter()
for (k, c), d in zip(cnt.items(), dem):
    if "cub::" in d:
        continue
    tot.update(c)
    if sum(c.values()):
        print(re.sub(r"\((int|bool)\)", "", d)[:110])
        for kk, v in c.items():
            print(f"    {kk:50s} {v}")
print("\nTOTAL (library kernels)")
for kk in PAT:
    print(f"    {kk:50s} {tot[kk]}")
