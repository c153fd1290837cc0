"""Per-kernel table from an ncu CSV log that carries several metrics per launch (`--metrics a,b,c --csv`).

    python tools/ncu_metrics_table.py gpurun_out/inst.csv > profiles/kernel_instructions_<tag>.txt

Averages every metric over the launches of a kernel and adds `issue_us` = warp instructions / (SM sub-partitions x
clock): the time the whole GPU would need to ISSUE that kernel's instructions at 1 per cycle per sub-partition —
the currency of a stage that runs many images concurrently and is issue-bound rather than HBM-bound.
"""
import csv
import re
import sys
from collections import OrderedDict

SMSP = 148 * 4
CLOCK_HZ = 1.965e9


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    return name.replace("nttt::", "").replace("void ", "").strip()


def main() -> None:
    with open(sys.argv[1]) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per = OrderedDict()   # kernel -> metric -> [values]
    ids = {}
    for r in csv.DictReader(lines):
        k = short(r["Kernel Name"])
        if "at::" in k or "cub::" in k:
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        per.setdefault(k, OrderedDict()).setdefault(r["Metric Name"], []).append(v)
        ids.setdefault(k, set()).add(r["ID"])
    metrics = []
    for k in per:
        for m in per[k]:
            if m not in metrics:
                metrics.append(m)
    print("kernel | launches | " + " | ".join(metrics) + " | issue_us")
    tot_issue = 0.0
    for k, ms in per.items():
        row = [k, str(len(ids[k]))]
        inst = None
        for m in metrics:
            vals = ms.get(m)
            if not vals:
                row.append("-")
                continue
            avg = sum(vals) / len(vals)
            if m.startswith("sm__inst_executed.sum") or m == "smsp__inst_executed.sum":
                inst = avg
            row.append(f"{avg:.4g}")
        iu = inst / SMSP / CLOCK_HZ * 1e6 if inst is not None else float("nan")
        if inst is not None and not k.startswith(("aa_", "proto_prepare")):
            tot_issue += iu
        row.append(f"{iu:.2f}")
        print(" | ".join(row))
    print(f"# sum of issue_us over the stage's kernels (one launch each): {tot_issue:.1f} us")


if __name__ == "__main__":
    main()
