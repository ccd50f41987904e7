"""Summarise an ncu gpu__time_duration launch list: python tools/launch_table.py FILE [first_preprocess_index]"""
import csv
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1000.0) for r in rows]
idx = [i for i, n in enumerate(names) if "preprocess" in n[0]]
start = idx[which]
end = idx[which + 1] if which + 1 < len(idx) and which != -1 else len(names)
seg = names[start:end]
tot = sum(t for _, t in seg)
for i, (n, t) in enumerate(seg):
    short = n.split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    print(f"{i:3d} {short[:48]:48s} {t:8.1f} us")
print(f"total {tot:.1f} us over {len(seg)} launches")
