"""Summaries of ncu --csv logs (read here, after the GPU run):

    python profiles/tools/ncu_csv_summary.py list    gpurun_out/r2_final_launches.csv      > profiles/r2/...summary.txt
    python profiles/tools/ncu_csv_summary.py metrics gpurun_out/r2_final_conv_metrics.csv  > profiles/r2/...table.txt

`list`: launches / total us / share per kernel (template arguments kept, parameter list dropped).
`metrics`: per kernel: launches, mean us, DRAM bytes read / written per launch, tensor-pipe activity (time-weighted),
DRAM throughput %, and the achieved DRAM GB/s (bytes / duration)."""
import collections
import csv
import re
import sys


def rows(path):
    with open(path, newline="") as f:
        lines = f.readlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    yield from csv.DictReader(lines[start:])


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").strip()
    return name


def to_ns(value, unit):
    v = float(value.replace(",", ""))
    return v * {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}.get(unit, 1.0)


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main():
    mode, path = sys.argv[1], sys.argv[2]
    per_id = collections.OrderedDict()
    for r in rows(path):
        d = per_id.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r["Grid Size"]})
        m, u, v = r["Metric Name"], r["Metric Unit"], r["Metric Value"]
        if m == "gpu__time_duration.sum":
            d["ns"] = to_ns(v, u)
        elif m.startswith("dram__bytes_read"):
            d["rd"] = to_bytes(v, u)
        elif m.startswith("dram__bytes_write"):
            d["wr"] = to_bytes(v, u)
        elif m.startswith("sm__pipe_tensor_cycles_active"):
            d["tensor"] = float(v)
        elif m.startswith("gpu__dram_throughput"):
            d["dram_pct"] = float(v)
    agg = collections.OrderedDict()
    for d in per_id.values():
        a = agg.setdefault(d["kernel"], collections.defaultdict(float))
        a["n"] += 1
        a["ns"] += d.get("ns", 0.0)
        a["rd"] += d.get("rd", 0.0)
        a["wr"] += d.get("wr", 0.0)
        a["tensor_w"] += d.get("tensor", 0.0) * d.get("ns", 0.0)
        a["dram_w"] += d.get("dram_pct", 0.0) * d.get("ns", 0.0)
    total = sum(a["ns"] for a in agg.values())
    order = sorted(agg.items(), key=lambda kv: -kv[1]["ns"])
    if mode == "list":
        print(f"{len(per_id)} launches, {total / 1e6:.2f} ms (gpu__time_duration.sum; serialised, cold caches)\n")
        print(f"{'kernel':64s} {'launches':>8s} {'us':>10s} {'share':>7s}")
        for k, a in order:
            print(f"{k[:64]:64s} {int(a['n']):8d} {a['ns'] / 1e3:10.1f} {100 * a['ns'] / total:6.1f}%")
    else:
        print(f"{'kernel':44s} {'n':>4s} {'us/launch':>10s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>7s} {'dram %':>7s} {'tensor %':>8s}")
        for k, a in order:
            n, ns = a["n"], max(a["ns"], 1.0)
            print(f"{k[:44]:44s} {int(n):4d} {a['ns'] / n / 1e3:10.1f} {a['rd'] / n / 1e6:9.2f} {a['wr'] / n / 1e6:9.2f} "
                  f"{(a['rd'] + a['wr']) / ns:7.0f} {a['dram_w'] / ns:7.1f} {a['tensor_w'] / ns:8.1f}")
        print(f"\nall: {len(per_id)} launches, {total / 1e6:.3f} ms, DRAM {sum(a['rd'] for a in agg.values()) / 1e9:.3f} GB read, "
              f"{sum(a['wr'] for a in agg.values()) / 1e9:.3f} GB written")


if __name__ == "__main__":
    main()
