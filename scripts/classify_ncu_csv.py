"""Classify the kernel names in ncu CSVs with the reference's OWN clean_kernel_names (utils/plot_kernels.py:95-109, staged in
baseline/_ref/plot_kernels.py; only that function is extracted — the module imports matplotlib, which this image lacks)."""
import ast
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "baseline", "_ref", "plot_kernels.py")).read()
fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "clean_kernel_names")
ns = {}
exec(compile(ast.Module(body=[fn], type_ignores=[]), "plot_kernels.py", "exec"), ns)
clean = ns["clean_kernel_names"]

print("| csv | kernel (demangled, as ncu reports it) | reference's class | launches | gpu__time_duration us (sum) |")
print("|---|---|---|---|---|")
for path in sys.argv[1:]:
    lines = open(path).read().splitlines()
    hdr = next((i for i, l in enumerate(lines) if l.startswith('"ID","Process ID","Process Name"')), None)
    if hdr is None:
        print(f"| {os.path.basename(path)} | (no ncu table found) | | | |")
        continue
    rows = list(csv.DictReader(lines[hdr:]))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    for k, (n, us) in agg.items():
        print(f"| {os.path.basename(path)} | `{k[:110]}` | **{clean(k)}** | {n} | {us:.1f} |")
