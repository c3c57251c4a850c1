"""SASS opcode summary of libbmpc.so (evidence that the kernels are sm_100a cubins using the FP64 tensor pipe, TMA bulk
copies, mbarriers and warp REDUX): `cuobjdump -sass` of every kernel, opcode histogram of the ones that matter.
usage: python tools/sass_summary.py [lib] > profiles/sass_r02_summary.json"""
import collections
import json
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "modelpredictivecontrol.jl_b200", "libbmpc.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Za-z0-9_.]+)?)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
watch = ["DMMA.8x8x4", "UBLKCP.S.G", "SYNCS", "REDUX", "CREDUX", "SHFL", "DFMA", "MUFU.RSQ64H", "MUFU.RCP64H", "LDS", "STS",
         "ATOMG", "MEMBAR", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "NANOSLEEP"]
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
rows, total = [], collections.Counter()
for name, c in kernels.items():
    tot = sum(c.values())
    sel = {}
    for w in watch:
        v = sum(n for op, n in c.items() if op == w or op.startswith(w + ".") or op.startswith(w))
        if v:
            sel[w] = v
            total[w] += v
    rows.append({"kernel": demangle(name)[:120], "instructions": tot, "opcodes": sel})
print(json.dumps({"library": os.path.relpath(lib), "arch": arch, "kernels": len(rows), "totals": dict(total),
                  "note": "DMMA.8x8x4 = mma.sync.m8n8k4.f64 (FP64 tensor pipe; tcgen05 has no fp64 kind, so no UTC*MMA / TMEM is expected); "
                          "UBLKCP.S.G = cp.async.bulk global->shared (TMA bulk copy) completing on an mbarrier (SYNCS.*); REDUX / CREDUX = "
                          "redux.sync; ST/LD.E.STRONG.SYS = the release / acquire accesses of the epoch-flag gather",
                  "per_kernel": sorted(rows, key=lambda r: -r["instructions"])[:40]}, indent=1))
