"""Aggregate an `ncu --csv --metrics ...` log into JSON: per (kernel name, grid, block) the number of sampled launches and
the mean of every metric.  Usage: python scripts/ncu_csv_summary.py in.csv [out.json]"""
import csv
import io
import json
import sys


def summarise(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    if start < 0:
        raise SystemExit(f"{path}: no ncu CSV header found")
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = {}
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("eda::(anonymous namespace)::", "").replace("void ", "")
        key = f'{name} grid={r["Grid Size"]} block={r["Block Size"]}'
        e = agg.setdefault(key, {"ids": set(), "m": {}})
        e["ids"].add(r["ID"])
        try:
            val = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        m = e["m"].setdefault(f'{r["Metric Name"]} [{r["Metric Unit"]}]', [0.0, 0])
        m[0] += val
        m[1] += 1
    out = {}
    for key, e in agg.items():
        out[key] = {"launches": len(e["ids"]), **{k: v[0] / v[1] for k, v in sorted(e["m"].items())}}
    return out


if __name__ == "__main__":
    res = summarise(sys.argv[1])
    js = json.dumps(res, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(js)
    else:
        print(js)
