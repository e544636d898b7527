"""Per-kernel hot spots of an ncu report with source import: python tools/ncu_hot.py report.ncu-rep [min_samples]
Prints, for every captured launch, headline raw metrics and the SASS instructions with the most stall samples (with the
preceding instructions for context)."""
import csv, subprocess, sys, io

rep = sys.argv[1]
min_s = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    print("== launch:", {h: r[i] for i, h in enumerate(hdr) if h in want})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) > 10:
        cur["data"].append(r)
stalls = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_math",
          "stall_dispatch", "stall_not_selected", "stall_selected", "stall_sleep", "stall_membar", "stall_tex", "stall_no_inst"]
for b in blocks:
    idx = {h: i for i, h in enumerate(b["hdr"])}
    S, I = idx["# Samples"], idx["Instructions Executed"]
    data = b["data"]
    tot = sum(int(r[S] or 0) for r in data)
    print("\n#### %s   total samples %d, SASS instructions %d" % (b["name"][:90], tot, len(data)))
    agg = {c: sum(int(r[idx[c]] or 0) for r in data) for c in stalls if c in idx}
    print("   stall mix:", {k[6:]: v for k, v in agg.items() if v})
    for i, r in enumerate(data):
        if int(r[S] or 0) >= min_s:
            st = " ".join("%s=%s" % (c[6:], r[idx[c]]) for c in stalls if c in idx and int(r[idx[c]] or 0) > 0)
            for q in data[max(0, i - 3):i]:
                print("      %s %5s %9s  %s" % (q[0][-5:], q[S], q[I], q[idx["Source"]][:90]))
            print("  >>> %s %5s %9s  %-70s %s" % (r[0][-5:], r[S], r[I], r[idx["Source"]][:70], st))
