"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None, title=""):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*$", "", name)
        name = re.sub(r"^void ", "", name)
        name = name.replace("eigb200::<unnamed>::", "").replace("eigb200::", "")
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        tot[name] += val * scale
        cnt[name] += 1
    total = sum(tot.values())
    lines = [f"# {title}", "", f"source: `{path}` (ncu gpu__time_duration.sum, --clock-control none; per-launch times are "
             "cold-cache and serialised: compare SHARES)", "", f"total kernel time {total:.1f} ms over {sum(cnt.values())} launches",
             "", "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k in sorted(tot, key=lambda k: -tot[k]):
        lines.append(f"| `{k[:110]}` | {cnt[k]} | {tot[k]:.2f} | {100*tot[k]/total:.1f}% |")
    text = "\n".join(lines) + "\n"
    if out:
        with open(out, "w") as f:
            f.write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "")
