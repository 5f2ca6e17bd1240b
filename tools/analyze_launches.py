"""Summarise an ncu per-launch CSV of tools/profile_step.py (one denoise step): per-launch and per-op-group times."""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    data = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = data.setdefault(int(row["ID"]), {"name": row["Kernel Name"]})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        d[row["Metric Name"]] = v
    return data


def labels(n_layers=8, n_ctrl=0, fused=False):
    head = ["packx", "tsemb", "te0", "te2", "packemb", "mod", "embed"]
    blk = ["lnT", "qkv", "smq", "smk", "ctx", "apply", "lnmod", "saout"]
    blk += ["fused_ca_ffn"] if fused else ["ln", "caq", "smq2", "caapply", "lnmod2", "caout", "lin1", "lin2", "lnmod3", "ffnout"]
    tail = ["packh", "out"] if fused else ["out"]
    return head + [b + str(i) for i in range(n_layers) for b in blk] + tail


def main():
    data = load(sys.argv[1])
    verbose = len(sys.argv) > 2
    labs = labels(fused=any("fused_block" in d["name"] for d in data.values()))
    tot = 0.0
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for i, d in enumerate(data.values()):
        t = d["gpu__time_duration.sum"]
        lab = labs[i] if i < len(labs) else "?"
        base = lab.rstrip("0123456789")
        tot += t
        a = agg[base]
        a[0] += t; a[1] += 1
        a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        a[3] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0))
        if verbose and (i < 25 or i == len(data) - 1):
            print(f"{lab:10s} {d['name'][:24]:24s} {t:9.1f} us")
    print(f"launches {len(data)}  total {tot:.1f} us")
    print(f"{'op':10s} {'us/launch':>10s} {'n':>3s} {'total us':>10s} {'share':>6s} {'GB/s':>8s} {'tensor%':>8s}")
    for k, (t, n, byts, tens) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:10s} {t / n:10.1f} {n:3d} {t:10.1f} {100 * t / tot:5.1f}% {byts / t / 1e3:8.0f} {tens / n:8.1f}")


if __name__ == "__main__":
    main()
