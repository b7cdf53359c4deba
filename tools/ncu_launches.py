"""Per-launch summary of an `ncu --csv` launch list of the specialised kernels of ONE tape (run here, no GPU needed).

usage: ncu_launches.py <launches.csv> <prof_one.json line file> <out.txt> [traffic.json tape-name]
The csv comes from
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,
      sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
      --clock-control none -k regex:ccu_seg --csv --log-file <launches.csv> python tools/prof_one.py <tape> 1 0 0 0 <N> 2
(two evaluations; the LAST `segments` launches -- the second evaluation of the first tile -- are summarised: per-launch times
are cold-cache and serialised, so shares are meaningful, absolutes are not).  Writes the table and, optionally, the DRAM
bytes per evaluation of the plan into profiles/r2_traffic.json (bench.py's roofline.traffic)."""
import csv
import json
import sys


def main():
    path, info_path, out = sys.argv[1], sys.argv[2], sys.argv[3]
    info = None
    for ln in open(info_path):
        if ln.startswith("{"):
            info = json.loads(ln)
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    col = {k: h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
    launches = {}
    for r in rows[1:]:
        launches.setdefault(int(r[col["ID"]]), {})[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
    ids = sorted(launches)
    S = info["jit_segments"]
    tile = min(info["N"], info["jit_tile"] or info["N"])
    ids = ids[-S * ((info["N"] + tile - 1) // tile):][:S]  # first tile of the last evaluation
    tot_t = sum(launches[i]["gpu__time_duration.sum"] for i in ids)
    rd = sum(launches[i]["dram__bytes_read.sum"] for i in ids)
    wr = sum(launches[i]["dram__bytes_write.sum"] for i in ids)

    def scale(v, unit_hint):  # ncu prints bytes in the unit of the column; --csv raw values are plain numbers
        return v
    with open(out, "w") as f:
        f.write("%s tape, automatic plan: %d segments, %d threads/CTA, max %d registers, scratch slots %d, cross loads %d stores %d; "
                "one tile of %d instances\n" % (info["tape"], S, info["jit_threads"], info["jit_max_regs"], info["jit_scratch_slots"],
                                                  info["jit_cross_loads"], info["jit_cross_stores"], tile))
        f.write("DRAM bytes per evaluation: read %.0f + write %.0f = %.0f; device time of the launches %.1f us\n\n" % (
            rd / tile, wr / tile, (rd + wr) / tile, tot_t / 1e3))
        f.write("seg   time_us   rd_MB   wr_MB  dram%  fp64%  issue%  regs\n")
        for n, i in enumerate(ids):
            m = launches[i]
            f.write("%3d  %8.1f  %6.1f  %6.1f  %5.1f  %5.1f  %6.1f  %4d\n" % (
                n, m["gpu__time_duration.sum"] / 1e3, m["dram__bytes_read.sum"] / 1e6, m["dram__bytes_write.sum"] / 1e6,
                m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0),
                m.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0),
                m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0), m.get("launch__registers_per_thread", 0)))
    if len(sys.argv) > 5:
        tj, name = sys.argv[4], sys.argv[5]
        try:
            d = json.load(open(tj))
        except Exception:
            d = {}
        d[name] = {"dram_bytes_per_eval": round((rd + wr) / tile), "segments": S, "scratch_slots": info["jit_scratch_slots"],
                   "source": out}
        json.dump(d, open(tj, "w"), indent=1, sort_keys=True)
    print(open(out).read())


if __name__ == "__main__":
    main()
