"""Print key metrics (and optionally top stall lines) of an .ncu-rep: python tools/ncu_summary.py REP [--source]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum")
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for i, h in enumerate(hdr):
        if h in keep or "stalled" in h and "per_issue_active" in h:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if "stalled" in h and v < 0.3:
                continue
            print(f"   {h:95s} {rows[1][i]:10s} {r[i]}")
if "--source" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    # find header row
    for k, r in enumerate(srows):
        if "Source" in r and any("Sampl" in c for c in r):
            h = r
            body = srows[k + 1:]
            break
    else:
        print("no source page")
        sys.exit()
    si = h.index("Source")
    ci = [i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
    ci = ci[0] if ci else [i for i, c in enumerate(h) if "Sampl" in c][0]
    tot = 0
    items = []
    for r in body:
        try:
            v = float(r[ci].replace(",", ""))
        except (ValueError, IndexError):
            continue
        tot += v
        items.append((v, r[si].strip()[:110]))
    items.sort(reverse=True)
    print("total samples", tot, "column:", h[ci])
    for v, s in items[:22]:
        print(f"  {100 * v / max(tot, 1):5.1f}%  {s}")
