#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log into a per-kernel launch list (text)."""
import csv, sys, collections

def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(r[ci["Kernel Name"]], [0, 0.0])
        a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f"# {title}")
    print("# (gpu__time_duration.sum, --clock-control none; cold-cache serialised times: compare SHARES)")
    print(f"# total kernel time {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches\n")
    print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {ms:10.3f} {100 * ms / tot:6.1f}%")

if __name__ == "__main__":
    main()
