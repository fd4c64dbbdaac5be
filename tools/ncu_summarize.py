"""Condenses an `ncu --page raw --csv` export (tools/gpu_r2.sh) into a per-kernel table for profiles/: duration, DRAM
bytes read + written, achieved DRAM GB/s and % of peak, warp slots in use, L1/L2 hit rates, top stall reasons."""
import csv, gzip, sys, collections

path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def find(sub):
    for h in hdr:
        if sub in h:
            return col[h]
    return None


def num(r, c):
    if c is None:
        return None
    try:
        return float(r[c].replace(",", ""))
    except ValueError:
        return None


C = {k: find(v) for k, v in dict(
    dur="gpu__time_duration.sum", rd="dram__bytes_read.sum", wr="dram__bytes_write.sum", dpct="dram__throughput.avg.pct_of_peak_sustained_elapsed",
    warps="sm__warps_active.avg.pct_of_peak_sustained_active", l1="l1tex__t_sector_hit_rate.pct", l2="lts__t_sector_hit_rate.pct",
    cyc="sm__cycles_elapsed.max", regs="launch__registers_per_thread", smem="launch__shared_mem_per_block_dynamic",
    ipc="sm__inst_executed.avg.per_cycle_elapsed").items()}
stall_cols = [(h.split("smsp__average_warps_issue_stalled_")[-1].split("_per_issue_active")[0], i) for h, i in col.items()
              if "smsp__average_warps_issue_stalled_" in h and "per_issue_active" in h and "not_issued" not in h]
unit = lambda c: units[c] if c is not None else ""
groups = collections.OrderedDict()
for r in data:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    key = (short[:110], r[col["Grid Size"]], r[col["Block Size"]])
    groups.setdefault(key, []).append(r)
print("source: %s (%d kernel instances; one 4096^2 PIC/FLIP step with the PCG capped at 3 iterations, ncu --set full --clock-control none)" % (path, len(data)))
print("units: duration %s, dram bytes %s/%s" % (unit(C["dur"]), unit(C["rd"]), unit(C["wr"])))
for (short, grid, block), rs in groups.items():
    durs = [num(r, C["dur"]) or 0.0 for r in rs]
    rs_all = rs
    rs = [rs[max(range(len(rs)), key=lambda q: durs[q])]]  # metrics of the longest instance (gated no-op launches are in the list too)
    def avg(c):
        v = [num(r, c) for r in rs]
        v = [x for x in v if x is not None]
        return sum(v) / len(v) if v else None
    dur, rd, wr = avg(C["dur"]), avg(C["rd"]), avg(C["wr"])
    scale_t = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit(C["dur"]), 1e-9)
    scale_b = lambda c: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit(c), 1.0)
    secs = dur * scale_t if dur else None
    bytes_ = (rd or 0) * scale_b(C["rd"]) + (wr or 0) * scale_b(C["wr"])
    gbs = bytes_ / secs / 1e9 if secs else 0
    stalls = sorted(((num(rs[0], i) or 0, n) for n, i in stall_cols), reverse=True)[:4]
    print("\n%s   grid %s block %s   x%d  durations (%s): %s" % (short, grid, block, len(rs_all), unit(C["dur"]), " ".join("%.3f" % d for d in durs)))
    print("   duration %.1f us | DRAM read %.1f MB write %.1f MB -> %.0f GB/s (%.1f %% of peak by ncu) | warps active %.1f %% | L1 hit %.1f %% L2 hit %.1f %% | regs %s smem %s %s | IPC %.2f" % (
        (secs or 0) * 1e6, (rd or 0) * scale_b(C["rd"]) / 1e6, (wr or 0) * scale_b(C["wr"]) / 1e6, gbs, avg(C["dpct"]) or 0, avg(C["warps"]) or 0,
        avg(C["l1"]) or 0, avg(C["l2"]) or 0, rs[0][C["regs"]] if C["regs"] is not None else "?", rs[0][C["smem"]] if C["smem"] is not None else "?", unit(C["smem"]), avg(C["ipc"]) or 0))
    print("   top stalls (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls))
