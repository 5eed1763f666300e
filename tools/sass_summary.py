#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libhaslr_b200.so the instruction mix (packed DPX, warp collectives, streaming stores,
async copies, tensor / TMA mnemonics) and the row loop of the deep fill verbatim. usage: tools/sass_summary.py [lib.so] > profiles/...txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "haslr_b200/libhaslr_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["VIADDMNMX.S16x2", "VIADDMNMX", "VIMNMX.S16x2", "VIMNMX3", "VIMNMX", "VIADD.16x2", "PRMT", "CREDUX", "REDUX", "SHFL", "VOTE", "MATCH",
        "STG.E.EF.128", "STG", "LDG.E.128", "LDG", "LDS.128", "STS.128", "LDGSTS", "ATOM", "RED", "CCTL", "UBLKCP", "UTMALDG", "UTCHMMA", "HMMA", "LDTM", "BAR", "NANOSLEEP"]
fn, ins = None, collections.OrderedDict()
arch = re.findall(r"arch = (sm_\w+)", sass)
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = m.group(1); ins[fn] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m and fn:
        ins[fn].append((int(m.group(1), 16), m.group(2).strip()))
print(f"# {lib}: cubins for {sorted(set(arch))}")
print("# instruction mix per kernel (static SASS instruction counts)\n")
for fn, body in ins.items():
    if not re.search(r"k_poa|k0_|k1_|k2_|k4_|k_scan|k_exclusive", fn):
        continue
    cnt = collections.Counter()
    for _, t in body:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        for k in KEYS:
            if op.startswith(k):
                cnt[k] += 1
                break
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"{name}: {len(body)} instructions; " + ", ".join(f"{k} {v}" for k, v in cnt.items() if v))
# the deep fill's row loop: from the plan shuffle to the backward branch (inside k_poa_pool: rel_stripe<true> is inlined there)
print("\n# row loop of the deep fill (rel_stripe / row_rel, as inlined in k_poa_edges_deep): one iteration = one 512-cell row\n")
for fn, body in ins.items():
    if "k_poa_edges_deep" not in fn:
        continue
    stg = [i for i, (_, t) in enumerate(body) if "STG.E.EF.128" in t]
    if len(stg) < 2:
        continue
    last = stg[-1]
    end = next(i for i in range(last, len(body)) if re.search(r"\bBRA\b", body[i][1]) and re.search(r"0x([0-9a-f]+)", body[i][1]) and int(re.search(r"0x([0-9a-f]+)", body[i][1]).group(1), 16) < body[i][0])
    tgt = int(re.search(r"0x([0-9a-f]+)", body[end][1]).group(1), 16)
    for a, t in body:
        if tgt <= a <= body[end][0]:
            print(f"  /*{a:05x}*/ {t}")
