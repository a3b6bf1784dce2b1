"""SASS opcode histogram per kernel of liboblas_b200.so (cuobjdump -sass) -> profiles/rNN_sass_opcodes.json:
the evidence that the tensor-core kernels issue tcgen05 (UTCOMMA / UTCBAR / LDTM / STTM), that their operands
arrive by bulk copies (UBLKCP) and that the streaming kernels move 128 bits per access.
    python tools/sass_histogram.py profiles/r02_sass_opcodes.json"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "oblas_b200", "liboblas_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEY = ("UTCOMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UBLKCP", "SYNCS", "LDS.128", "STS.128", "LDG.E.128", "STG.E.128", "LDG", "STG",
       "PRMT", "LOP3", "SHF", "IMAD", "FADD", "SHFL", "BAR", "ELECT", "REDUX", "ATOM", "MEMBAR", "FENCE")
out, cur, name = {}, None, None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = out.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        cur["instructions"] += 1
        for k in KEY:
            if op == k or op.startswith(k + ".") or (k in ("LDG", "STG", "LDS.128", "STS.128", "LDG.E.128", "STG.E.128") and op.startswith(k)):
                cur[k] += 1
        if op.startswith("UTCOMMA") and ".2CTA" in ln:
            cur["UTCOMMA.2CTA"] += 1
res = {"library": "oblas_b200/liboblas_b200.so", "how": "cuobjdump -sass, opcode prefixes counted per kernel (static instruction counts)",
       "kernels": {k: dict(sorted(v.items())) for k, v in sorted(out.items()) if v["instructions"] > 0}}
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_opcodes.json")
json.dump(res, open(dst, "w"), indent=1)
for k, v in res["kernels"].items():
    tc = {x: v[x] for x in ("UTCOMMA", "UTCOMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UBLKCP") if x in v}
    if tc:
        print(k[-70:], tc)
