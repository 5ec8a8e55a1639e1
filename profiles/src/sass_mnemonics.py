"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass): which kernels carry tcgen05 (UTC*MMA, LDTM/STTM), bulk /
TMA copies (UBLKCP, UTMALDG), mbarriers (SYNCS), cp.async (LDGSTS) and mma.sync (HMMA).
   python profiles/src/sass_mnemonics.py > profiles/r2_sass_evidence.txt"""
import collections
import re
import subprocess
import sys

sys.path.insert(0, ".")
from multimodal_mtrssm_b200 import build

lib = "multimodal_mtrssm_b200/librssm_rollout.so"
KEYS = ["HMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDSM", "BAR", "ATOMG", "RED", "MUFU"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts: dict[str, collections.Counter] = {}
total: dict[str, int] = {}
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("rssm::", "").replace("void ", "")
        counts[cur], total[cur] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
print(f"library {lib}, build stamp {build.build_stamp()}, sm_100a; columns = static instruction counts")
print(f"{'kernel':86s} {'instrs':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
for name in sorted(counts):
    c = counts[name]
    print(f"{name[:86]:86s} {total[name]:7d} " + " ".join(f"{c[k]:7d}" if c[k] else f"{'.':>7s}" for k in KEYS))
