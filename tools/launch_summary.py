"""dev tool: aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) per kernel:
     python tools/launch_summary.py gpurun_out/r2b_launches_B256.csv profiles/r2b_launches_B256 "title"
   writes <out>_summary.md and <out>.csv.gz (the raw list)."""
import csv
import gzip
import re
import shutil
import sys
from collections import defaultdict

src, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
lines = [l for l in open(src, errors="replace") if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, cnt = defaultdict(float), defaultdict(int)
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"^void ", "", r[ik])
    name = re.split(r"[<(]", name)[0]
    ms = float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    agg[name] += ms
    cnt[name] += 1
tot = sum(agg.values())
ours = sum(v for k, v in agg.items() if k.startswith(("xq::", "xqv::", "xql::", "xqd::")))
cublas = sum(v for k, v in agg.items() if k.startswith(("nvjet", "cutlass", "sm100_", "sm90_", "cublas")))
md = [f"# {title}\n", f"Raw list: `{out.split('/')[-1]}.csv.gz` ({sum(cnt.values())} launches, {tot:.1f} ms; per-launch times are cold-cache and "
      "serialised -- compare SHARES, not absolutes).\n",
      f"libxqb200 kernels: {100 * ours / tot:.1f} %   cuBLAS GEMMs (nvjet / cutlass): {100 * cublas / tot:.1f} %   "
      f"other (torch elementwise / reductions, cuDNN): {100 * (tot - ours - cublas) / tot:.1f} %\n",
      "| share | ms | launches | kernel |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:40]:
    md.append(f"| {100 * v / tot:.1f} % | {v:.2f} | {cnt[k]} | `{k}` |")
open(out + "_summary.md", "w").write("\n".join(md) + "\n")
with open(src, "rb") as f, gzip.open(out + ".csv.gz", "wb") as g:
    shutil.copyfileobj(f, g)
print("\n".join(md[:30]))
