"""Family shares of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file ...`).
  python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt
gpu__time_duration.sum per launch is cold-cache and serialised: compare SHARES with the bench line's kernel_ms_per_nfe, not absolutes."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    v = float(r[mv].replace(",", ""))
    us = {"ns": v / 1e3, "us": v, "usecond": v, "nsecond": v / 1e3, "ms": v * 1e3, "msecond": v * 1e3}.get(r[mu], v / 1e3)
    name = re.sub(r"^void ", "", r[kn])
    name = re.sub(r"\(.*$", "", name).replace("oron::", "")
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
print(f"{len(rows)} launches, {total / 1e3:.2f} ms of kernel time (cold-cache, serialised)\n")
print(f"{'kernel':72s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>8s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:24]:
    print(f"{k[:72]:72s} {cnt[k]:8d} {v:12.1f} {100 * v / total:6.1f}% {v / cnt[k]:8.2f}")
fam = defaultdict(float)
for k, v in tot.items():
    f = ("gemm" if ("gemm" in k or "gconv" in k or "ffn2" in k) else "attention" if "attn" in k else "ln_modulate" if "ln_modulate" in k else "other")
    fam[f] += v
core = fam["gemm"] + fam["attention"] + fam["ln_modulate"]
print("\nfamilies: " + "  ".join(f"{f} {100 * v / total:.1f}%" for f, v in sorted(fam.items(), key=lambda kv: -kv[1])))
print("of gemm + attention + ln: " + "  ".join(f"{f} {fam[f] / core:.3f}" for f in ("gemm", "attention", "ln_modulate")))
