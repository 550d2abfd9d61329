"""Median kernel durations per phase of scripts/run_graph_bwd.py from an ncu launch list (gpurun_out/graph_bwd.csv)."""
import csv, statistics, sys
rows = list(csv.reader(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/graph_bwd.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
seq = []
for r in rows[hi + 1:]:
    if len(r) < len(ix):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("iapt::", "").replace("void ", "")
    seq.append((name, r[ix["Grid Size"]], float(r[ix["Metric Value"]]) / 1000.0))
out, cur = [], []
for name, grid, us in seq:  # a call = optional inverse-index kernels followed by gradient kernel(s)
    cur.append((name, grid, us))
    if "grad" in name and not name.startswith("graph_feature_grad_q"):
        key = tuple((n, g) for n, g, _ in cur)
        if out and out[-1][0] == key:
            out[-1][1].append([u for _, _, u in cur])
        else:
            out.append([key, [[u for _, _, u in cur]]])
        cur = []
for key, ts in out:
    med = [statistics.median(x[i] for x in ts) for i in range(len(key))]
    print(len(ts), " + ".join("%s%s %.1f" % (n[:28], g, m) for (n, g), m in zip(key, med)), "= %.1f us" % sum(med))
