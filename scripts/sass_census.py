"""Per-kernel SASS census of ocrl_b200/csrc/libocrl_sa.so -> profiles/r2/sass_census.txt (run in the build container).
Counts the instructions that prove what a kernel runs on: UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP (bulk copies), HMMA (mma.sync), FFMA, SYNCS (mbarrier)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ocrl_b200", "csrc", "libocrl_sa.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
pat = [("UTC*MMA", r"\bUTC\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"),
       ("UTMASTG", r"\bUTMASTG\b"), ("UBLKCP", r"\bUBLKCP\b"), ("HMMA", r"\bHMMA\b"), ("FFMA", r"\bFFMA\b"),
       ("SYNCS", r"\bSYNCS\b"), ("LDGSTS", r"\bLDGSTS\b")]
out = ["SASS census of ocrl_b200/csrc/libocrl_sa.so (cuobjdump -sass, sm_100a), instructions per kernel",
       "kernel | " + " | ".join(p[0] for p in pat) + " | total"]
blocks = sass.split("Function : ")[1:]
for blk, name in zip(blocks, names):
    lines = [ln for ln in blk.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/", ln)]
    cnt = collections.OrderedDict((k, sum(1 for ln in lines if re.search(rx, ln))) for k, rx in pat)
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    out.append(f"{short} | " + " | ".join(str(v) for v in cnt.values()) + f" | {len(lines)}")
os.makedirs(os.path.join(ROOT, "profiles", "r2"), exist_ok=True)
open(os.path.join(ROOT, "profiles", "r2", "sass_census.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
