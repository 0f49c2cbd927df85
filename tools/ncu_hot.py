"""Top SASS instructions by warp-stall samples for one kernel of an .ncu-rep (needs --import-source on).
Usage: python tools/ncu_hot.py file.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
fn, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", fn, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name"')
for blk in blocks[1:2]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    print(rows[0][1][:100])
    hdr = rows[1]
    si, ii, ai, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Address"), hdr.index("Instructions Executed")
    data = [(int(r[ii] or 0), k, r[si].strip(), int(r[ei] or 0)) for k, r in enumerate(rows[2:]) if len(r) > ii]
    tot = sum(d[0] for d in data)
    print("total samples", tot, "instructions", len(data), "executed", sum(d[3] for d in data))
    for smp, k, src, ex in sorted(data, reverse=True)[:top]:
        print(f"{100*smp/tot:5.1f}%  #{k:4d} x{ex:8d}  {src[:90]}")
