"""Stall samples of an ncu report (source page) summed over runs of SASS instructions: python tools/ncu_regions.py rep [run_len]
Prints address range, samples, executed count of the first instruction and the dominant opcodes of each run -- enough to see
which phase of a warp-specialised kernel the sampled warps sit in."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; run = int(sys.argv[2]) if len(sys.argv) > 2 else 48
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r
    elif hdr is not None and len(r) > 10: data.append(r)
idx = {h: i for i, h in enumerate(hdr)}
S, I, A, X = idx["# Samples"], idx["Instructions Executed"], idx["Address"], idx["Source"]
tot = sum(int(r[S] or 0) for r in data)
print("total samples", tot)
for i in range(0, len(data), run):
    blk = data[i:i + run]
    s = sum(int(r[S] or 0) for r in blk)
    if s < tot * 0.004: continue
    ops = collections.Counter(r[X].split()[1] if r[X].startswith("@") else r[X].split()[0] for r in blk)
    top = max(blk, key=lambda r: int(r[S] or 0))
    print("%s  samples %5d (%4.1f%%)  exec %8s  %s | hottest: %s (%s)" % (blk[0][A][-5:], s, 100.0 * s / tot, blk[0][I], dict(ops.most_common(4)), top[X][:50], top[S]))
