"""Summarise an ncu --set full report into JSON (one entry per captured launch): python tools/ncu_summary.py report.ncu-rep > out.json
Keeps the metrics DESIGN.md / bench.py quote: duration, DRAM bytes, registers, issue / pipe utilisation, warps active and the
stall-reason mix of the sampled warps (source page)."""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "launch__registers_per_thread": "registers_per_thread", "launch__grid_size": "grid", "launch__block_size": "block",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "smsp__inst_executed.sum": "warp_instructions", "sm__cycles_elapsed.max": "cycles",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts"}
out = []
for r in rows[2:]:
    e = {}
    for i, h in enumerate(hdr):
        if h == "Kernel Name":
            e["kernel"] = r[i]
        elif h in want:
            v = r[i].replace(",", "")
            try:
                v = float(v)
            except ValueError:
                pass
            key = want[h]
            if key in ("dram_read", "dram_write"):
                key += "_" + units[i].lower()
            e[key] = v
    out.append(e)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) > 10:
        cur["data"].append(r)
def norm(name):
    return name.replace("void ", "").replace("gnnml3::", "").replace("(int)", "").replace(" ", "").split("(")[0]


# the source page lists one block per launch but not necessarily in launch order: pair by kernel name, in order of appearance
pools = {}
for b in blocks:
    pools.setdefault(norm(b["name"]), []).append(b)
for e in out:
    cand = pools.get(norm(e["kernel"]), [])
    if not cand:
        continue
    b = cand.pop(0)
    idx = {h: i for i, h in enumerate(b["hdr"])}
    tot = sum(int(r[idx["# Samples"]] or 0) for r in b["data"])
    mix = {}
    for c in idx:
        if c.startswith("stall_") and "Not Issued" not in c:
            v = sum(int(r[idx[c]] or 0) for r in b["data"])
            if v:
                mix[c[6:]] = round(v / max(tot, 1), 3)
    e["stall_mix_of_sampled_warps"] = dict(sorted(mix.items(), key=lambda kv: -kv[1]))
    e["samples"] = tot
json.dump(out, sys.stdout, indent=1)
