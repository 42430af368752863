"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals / shares."""
import collections, csv, re, sys

def main(path, skip=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n += 1
        if n <= skip:
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {s/1e3:.2f} ms total (cold-cache, serialised: compare SHARES)")
    for k, v in tot.most_common(25):
        print(f"{v:11.1f} us {100*v/s:5.1f}%  n={cnt[k]:4d}  avg {v/cnt[k]:9.1f} us  {k[:110]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
