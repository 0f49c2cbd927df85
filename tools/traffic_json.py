"""profiles/*_traffic.json from an ncu launch list of the bench command (developer tool):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras
    python tools/traffic_json.py launches.csv[.gz] bench_line.json c3 > profiles/rNN_traffic.json

Per stage: measured DRAM bytes (read + write) per step next to the algorithmic bytes the bench line
reports; for the pivoted-QR passes also per launch.  A "step" = one qr_init_kernel launch in the list."""
import collections
import csv
import gzip
import json
import sys

STAGE_OF = [("block_tree_kernel", "stats"), ("block_top_kernel", "stats"), ("node_table_kernel", "stats"),
            ("row_means", "centre"), ("center_given", "centre"),
            ("gram_", "gram"), ("backproject_", "backproject"),
            ("qr_gemv_kernel", "qr_passes"), ("qr_apply", "qr_passes")]


def main(fn, bench_json, key):
    op = gzip.open if fn.endswith(".gz") else open
    rows = list(csv.reader(l for l in op(fn, "rt") if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(int(r[ii]), {"name": r[ki]})
        try:
            d[r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    steps = sum(1 for d in per.values() if "qr_init_kernel" in d["name"])
    agg = collections.OrderedDict()
    for d in per.values():
        nm = d["name"].replace("void ", "").replace("omb::", "")
        for pat, st in STAGE_OF:
            if nm.startswith(pat):
                a = agg.setdefault(st, [0, 0.0, 0.0])
                a[0] += 1
                a[1] += d.get("gpu__time_duration.sum", 0.0)
                a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
                break
    line = json.loads(open(bench_json).read().strip().splitlines()[-1])
    roof = line["roofline"]
    out = {"qr_block": line["config"]["qr_block"], "qr_lazy": bool(roof.get("lazy", {}).get("on")),
           "qr_lazy_alpha": roof.get("lazy", {}).get("alpha"), "steps_in_list": steps,
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none of "
                     "`python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras` (%d placements); "
                     "algorithmic bytes from the bench line of the same build" % steps,
           "stages": {}}
    for st, (cnt, t_ns, byt) in agg.items():
        alg = roof["algorithmic_bytes_per_step"] if st == "qr_passes" else roof["stages"].get(st, {}).get("algorithmic_bytes")
        ent = {"launches_per_step": cnt / steps, "dram_bytes_per_step": byt / steps, "ncu_time_ms_per_step": t_ns / steps / 1e6,
               "algorithmic_bytes_per_step": alg, "traffic_over_algorithmic": (byt / steps / alg) if alg else None}
        if st == "qr_passes":
            ent["schedule_launches_per_step"] = roof["launches_per_step"]      # the passes of the schedule (bench line)
            ent["dram_bytes_per_launch"] = byt / steps / roof["launches_per_step"]
            ent["note"] = ("launches include the conditional catch-up passes (no-ops unless a panel could not certify its "
                           "pivot); algorithmic bytes = apply passes (every column) + what the read-only passes visited, "
                           "counted by the kernels (lazy norm down-dates)")
            out["qr_passes"] = ent
        else:
            out["stages"][st] = ent
    print(json.dumps({key: out}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "c3")
