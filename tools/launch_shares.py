"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into per-kernel averages and shares.

    python tools/launch_shares.py gpurun_out/launches.csv "command that was profiled" > profiles/launch_shares_<tag>.csv

The one-off set-up kernels (antialias tables, prototype preparation) and torch's own kernels are listed but kept out
of the stage sum.  ncu serialises launches and runs them cold-cache: compare SHARES, not absolute times.
"""
import csv
import re
import sys
from collections import OrderedDict

SETUP = ("aa_table_kernel", "aa_transpose_kernel", "aa_group_kernel", "aa_pack_kernel", "proto_prepare_kernel",
         "fill_pool_kernel", "fill_scatter_kernel", "fill_finalize_kernel", "unpack_masks_kernel")  # (+ the fill phase)


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)          # drop the argument list
    name = name.replace("nttt::", "").replace("void ", "")
    return name.strip()


def main() -> None:
    path = sys.argv[1]
    what = sys.argv[2] if len(sys.argv) > 2 else "bench.py"
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) / 1e3))
    agg = OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    ours = {k: v for k, v in agg.items() if "at::" not in k and "cub::" not in k and not k.startswith(SETUP)}
    images = max(v[0] for k, v in ours.items() if k.startswith("lowres_pack"))
    stage_us = sum(v[1] for v in ours.values()) / images
    print(f"# ncu launch list (gpu__time_duration.sum, --clock-control none) of `{what}`")
    print("# per-kernel average over the captured launches; cold-cache and serialised: compare SHARES, not absolutes")
    print(f"# stage passes captured: {images}; sum of the stage's kernels per pass: {stage_us:.1f} us")
    print("kernel,launches,avg_us,us_per_pass,share_of_stage")
    for k, (n, tot) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"{k},{n},{tot / n:.1f},{tot / images:.1f},{tot / images / stage_us:.3f}")
    print("# not part of the stage (set-up, torch):")
    for k, (n, tot) in agg.items():
        if k not in ours:
            print(f"# {k},{n},{tot / n:.1f}")


if __name__ == "__main__":
    main()
