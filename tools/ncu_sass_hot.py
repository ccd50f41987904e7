"""Top stall-sample SASS lines of the first kernel in an .ncu-rep: python tools/ncu_sass_hot.py REP [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout.splitlines()
rows = list(csv.reader(out))
# split per kernel
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    h = b["rows"][0]
    body = b["rows"][1:]
    si, ci = h.index("Source"), h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(float(r[ci]) for r in body if len(r) > ci and r[ci].replace(".", "").isdigit())
    print("==", b["name"][:90], "samples", tot)
    agg = {}
    for r in body:
        for i in stall_cols:
            try:
                agg[h[i]] = agg.get(h[i], 0) + float(r[i])
            except (ValueError, IndexError):
                pass
    print("   stall totals:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    items = []
    for idx, r in enumerate(body):
        try:
            v = float(r[ci])
        except (ValueError, IndexError):
            continue
        st = sorted(((float(r[i]), h[i]) for i in stall_cols if r[i] not in ("", "0")), reverse=True)[:2]
        items.append((v, idx, r[si].strip()[:70], st))
    items.sort(reverse=True)
    for v, idx, s, st in items[:topn]:
        print(f"  {100 * v / max(tot, 1):5.1f}%  #{idx:5d} {s:70s} {st}")
