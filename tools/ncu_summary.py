"""Summarise an .ncu-rep (first kernel): key metrics + instruction mix + hottest source lines.  usage: ncu_summary.py rep [n_lines]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        try:
            v = float(vals[i])
        except Exception:
            v = vals[i]
        if "issue_stalled" in h and isinstance(v, float) and v < 0.15:
            continue
        print("%-95s %s %s" % (h, vals[i], rows[1][i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc, ie, ism = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    try:
        data.append((r[isrc].strip(), int(r[ie]), int(r[ism])))
    except Exception:
        pass
tot = sum(e for _, e, _ in data); sm = sum(s for _, _, s in data)
by = collections.Counter(); bys = collections.Counter()
for s_, e, smp in data:
    op = (s_.split()[1] if s_.startswith("@") else s_.split()[0]).split(".")[0]
    by[op] += e; bys[op] += smp
print("total warp instr", tot, "samples", sm)
for op, e in by.most_common(14):
    print("  %-10s instr %5.1f%%  samples %5.1f%%" % (op, 100 * e / tot, 100 * bys[op] / max(sm, 1)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if n:
    cum = 0
    print("-- cumulative profile by position (every ~2% of samples)")
    acc = 0; acce = 0
    for i, (s_, e, smp) in enumerate(data):
        acc += smp; acce += e
        if acc >= sm / n:
            print("  @%4d %-40s samples+%4.1f%% instr+%4.1f%%" % (i, s_[:40], 100 * acc / sm, 100 * acce / tot))
            acc = 0; acce = 0
