"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, median /
total duration and share of the captured window (times are cold-cache and serialised by ncu: shares only)."""
import csv, sys, re, statistics as st
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1]
    name = re.sub(r"<.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
    rows.append((name, v))
tot = sum(v for _, v in rows)
agg = {}
for n, v in rows:
    agg.setdefault(n, []).append(v)
print(f"# {len(rows)} launches, {tot:.1f} us total (ncu-serialised)")
print(f"{'kernel':28s} {'launches':>8s} {'median_us':>10s} {'total_us':>10s} {'share':>7s}")
for n, vs in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{n:28s} {len(vs):8d} {st.median(vs):10.2f} {sum(vs):10.1f} {100 * sum(vs) / tot:6.1f}%")
