"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md):
    cuobjdump -sass attack_vc_b200/libavc_b200.so | python scripts/sass_summary.py > profiles/<name>.md"""
import collections
import re
import subprocess
import sys

pat = re.compile(r"\b(UTCHMMA|UTCBAR|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|FFMA|LDGSTS)\b")
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for k in pat.findall(line):
            cnt[cur][k] += 1


def dem(n):
    out = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", out.replace("(anonymous namespace)::", "")).replace("void ", "")


print("SASS mnemonic counts per kernel of attack_vc_b200/libavc_b200.so (`cuobjdump -sass`, sm_100a).  UTCHMMA = tcgen05.mma,")
print("UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = cp.async.bulk.tensor (TMA tensor load),")
print("UBLKCP = cp.async.bulk (TMA engine, linear), LDGSTS = cp.async.  Kernels without any of them and < 200 FFMA are omitted.\n")
print("| kernel | UTCHMMA | UTCBAR | LDTM | STTM | UTMALDG | UBLKCP | LDGSTS | FFMA |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
for k, c in sorted(cnt.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 100000 + kv[1]["UTMALDG"] * 1000 + kv[1]["FFMA"])):
    if c["UTCHMMA"] + c["UTMALDG"] + c["UBLKCP"] + c["LDTM"] == 0 and c["FFMA"] < 200:
        continue
    print(f"| `{dem(k)}` | {c['UTCHMMA']} | {c['UTCBAR']} | {c['LDTM']} | {c['STTM']} | {c['UTMALDG']} | {c['UBLKCP']} | {c['LDGSTS']} | {c['FFMA']} |")
