"""Markdown table of one bench.py JSON line (headline + the `configs` array): what DESIGN.md section 4 quotes.
usage: python tools/bench_table.py profiles/r2_bench_final.json"""
import json
import sys


def row(name, c, cpu, e2e):
    r = c["roofline"]
    f = lambda v, fmt="%.3g": "–" if v is None else fmt % v
    return "| %s | %s | %s %.3f | %s | %s | %s |" % (
        name, f(c.get("value")), r["bound"], r["frac"], f((e2e or {}).get("value")),
        f((cpu or {}).get("value")), f(((cpu or {}).get("serial") or {}).get("value")))


def main(path):
    d = None
    for line in open(path):
        if line.startswith("{"):
            d = json.loads(line)
    print("| config | device-resident evals/s | roofline bound, frac | e2e evals/s (plugin, pageable) | reference openmp (%s cores) | reference serial |"
          % (d.get("cpu_baseline") or {}).get("cores", "?"))
    print("|---|---|---|---|---|---|")
    print(row(d["config"]["name"] + " (headline)", d, d.get("cpu_baseline"), d.get("e2e")))
    for c in d.get("configs", []):
        if c.get("value") is None:
            print("| %s | not measured: %s | | | | |" % (c["config"]["name"], c.get("skipped") or c.get("error")))
            continue
        print(row(c["config"]["name"], c, c.get("cpu_baseline"), c.get("e2e")))
    it = d.get("interpreter")
    ck = d.get("clocks") or {}
    print()
    print("headline: %.4g evals/s, %.2f ms/step, roofline.frac %.3f (at the sustained clock %.3f), traffic %s B/eval, SM %s / %s MHz %s, "
          "interpreter %s evals/s, %d launches in the timed region" % (
              d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["fp64"].get("frac_at_sustained_clock", float("nan")),
              d["roofline"].get("traffic"), ck.get("sm_mhz"), ck.get("sm_max_mhz"), ck.get("reasons"),
              "%.3g" % it["value"] if it else "–", d.get("gpu_launches", 0)))


if __name__ == "__main__":
    main(sys.argv[1])
