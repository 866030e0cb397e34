"""Dev tool: aggregate the per-instruction stall samples of an .ncu-rep source page by code region (generator / MMA).
usage: python tools/ncu_regions.py gpurun_out/prof.ncu-rep [block]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, data = rows[1], rows[2:]
ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
dm = [i for i, r in enumerate(data) if "DMMA" in r[ix["Source"]]]
first = dm[0] if dm else len(data)
# MMA region: from ~150 instructions before the first DMMA (fragment loads) to the end
regions = [("generator", 0, max(0, first - 150)), ("mma", max(0, first - 150), len(data))]
if blk:
    regions = [(f"[{lo},{min(lo + blk, len(data))})", lo, min(lo + blk, len(data))) for lo in range(0, len(data), blk)]
for name, lo, hi in regions:
    c = collections.Counter()
    samples = inst = 0
    for r in data[lo:hi]:
        samples += int(r[ix["# Samples"]])
        inst += int(r[ix["Instructions Executed"]])
        for s in stalls:
            c[s] += int(r[ix[s]])
    tot = max(1, sum(c.values()))
    print(f"{name:12s} instrs {hi - lo:5d} samples {samples:7d} warp-inst {inst:10d} | " + "  ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in c.most_common(8)))
