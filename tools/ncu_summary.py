"""Condense ncu CSVs.  usage: ncu_summary.py launches <launches.csv> | raw <raw.csv> [more raw.csv ...]"""
import csv, sys, collections
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.pct", "launch__grid_size", "launch__block_size",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]
def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
        name = r[ix["Kernel Name"]].split("(")[0][:70]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        if unit == "ns": v /= 1e3
        elif unit == "ms": v *= 1e3
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} us {100*t/tot:5.1f}%  x{c:<3d} {t/c:9.1f} us/launch  {k}")
def raw(path):
    rows = list(csv.reader(open(path)))
    h = rows[0]; units = rows[1]
    for r in rows[2:]:
        ix = {n: i for i, n in enumerate(h)}
        print("==", r[ix["Kernel Name"]][:90], r[ix["Grid Size"]], r[ix["Block Size"]])
        for k in KEYS:
            if k in ix: print(f"   {k:75s} {r[ix[k]]:>14s} {units[ix[k]]}")
if sys.argv[1] == "launches": launches(sys.argv[2])
else:
    for p in sys.argv[2:]: raw(p)
