"""ncu launch list (--csv --log-file, metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum)
-> a per-launch text summary under profiles/ and, for the descent, profiles/r02_descent_traffic.json (DRAM bytes per row
of the level kernels), which bench.py reports as roofline.traffic.

    python tools/summarise_launches.py gpurun_out/launches.csv profiles/r02_launches_summary.txt "<command line>" [rows] [queries]

`rows` = descriptor rows the listed descent launches processed (all listed steps together); `queries` = queries the listed
retrieval launches scored (writes profiles/r02_scorer_traffic.json from the score_shard* launches).
"""
import csv
import json
import os
import re
import sys


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    queries = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    by_name, sc_bytes, sc_ms = {}, 0.0, 0.0
    lines = [l for l in open(src) if l.startswith('"')]
    recs = {}
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        e = recs.setdefault(k, {"name": r["Kernel Name"]})
        e[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        e["unit_" + r["Metric Name"]] = r["Metric Unit"]
    out = ["ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none", "    " + cmd,
           "Per-launch times under ncu are serialised and cold-cache: compare SHARES with bench.py's live CUDA-event times.", ""]
    tot_ms, tot_rd, tot_wr, lvl_bytes, lvl_ms = 0.0, 0.0, 0.0, 0.0, 0.0
    for k in sorted(recs):
        e = recs[k]
        name = re.sub(r"\(.*", "", e["name"].replace("void <unnamed>::", "")).strip()
        t = e.get("gpu__time_duration.sum", 0.0)
        unit = e.get("unit_gpu__time_duration.sum", "ns")
        ms = t / 1e6 if unit in ("ns", "nsecond") else (t / 1e3 if unit in ("us", "usecond") else t)
        rd, wr = e.get("dram__bytes_read.sum", 0.0), e.get("dram__bytes_write.sum", 0.0)
        for m, v in (("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr")):
            u = e.get("unit_" + m, "byte")
            f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            if v == "rd":
                rd *= f
            else:
                wr *= f
        tot_ms, tot_rd, tot_wr = tot_ms + ms, tot_rd + rd, tot_wr + wr
        if "level_pair_kernel" in name or "level_mma_kernel" in name:
            lvl_bytes += rd + wr
            lvl_ms += ms
        g = by_name.setdefault(name[:60], [0, 0.0, 0.0])
        g[0], g[1], g[2] = g[0] + 1, g[1] + ms, g[2] + rd + wr
        if "score_shard" in name or "score_query" in name:
            sc_bytes += rd + wr
            sc_ms += ms
        out.append("%3d %-40s %8.3f ms  DRAM read %8.3f GB  write %7.3f GB  (%.2f TB/s)"
                   % (k, name[:40], ms, rd / 1e9, wr / 1e9, (rd + wr) / max(ms, 1e-9) / 1e9))
    out.append("")
    out.append("total %.3f ms, DRAM read %.2f GB + write %.2f GB = %.2f GB" % (tot_ms, tot_rd / 1e9, tot_wr / 1e9, (tot_rd + tot_wr) / 1e9))
    if lvl_ms:
        out.append("descent level kernels: %.3f ms (%.1f %% of the listed launches), %.2f GB of DRAM traffic"
                   % (lvl_ms, 100 * lvl_ms / tot_ms, lvl_bytes / 1e9))
    out.append("")
    out.append("by kernel (launches, total ms, share, DRAM GB):")
    for nm, g in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
        out.append("  %-60s %4d  %9.3f ms  %5.1f %%  %8.3f GB" % (nm, g[0], g[1], 100 * g[1] / max(tot_ms, 1e-9), g[2] / 1e9))
    if queries and sc_bytes:
        out.append("scorer launches: %.3f ms, %.2f GB of DRAM traffic over %d queries = %.2f MB per query"
                   % (sc_ms, sc_bytes / 1e9, queries, sc_bytes / queries / 1e6))
        j = {"dram_bytes_per_query": sc_bytes / queries, "queries": queries,
             "source": "ncu launch list %s (dram__bytes_read.sum + dram__bytes_write.sum of the scorer launches, %d queries)"
                       % (os.path.basename(dst), queries)}
        json.dump(j, open(os.path.join(os.path.dirname(dst), "r02_scorer_traffic.json"), "w"), indent=1)
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[-14:]))
    if rows and lvl_bytes:
        j = {"level_kernel_dram_bytes_per_row": lvl_bytes / rows, "step_dram_bytes_per_row": (tot_rd + tot_wr) / rows, "rows": rows,
             "source": "ncu launch list %s (dram__bytes_read.sum + dram__bytes_write.sum of the descent level launches of one step, "
                       "%d rows), scaled by rows" % (os.path.basename(dst), rows)}
        json.dump(j, open(os.path.join(os.path.dirname(dst), "r02_descent_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
