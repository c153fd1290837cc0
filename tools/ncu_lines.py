"""Per-source-line instruction / stall-sample shares of one kernel from an ncu report (needs -lineinfo + --import-source).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--top 40] [--file fullres.cu]

Reads `ncu -i <rep> --page source --print-source cuda,sass --csv` (first profiled launch) and prints, per CUDA source
line, its share of executed warp instructions and of warp-stall samples, plus the dominant stall reason.
"""
import argparse
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--file", default=None, help="only lines of source files whose path contains this")
    ap.add_argument("--kernel", default=None, help="only launches of this kernel (ncu -k)")
    args = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--print-source", "cuda,sass", "--csv"] +
                         (["-k", args.kernel] if args.kernel else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr, lines, launches = None, None, {}, 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            launches += 1
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        key = (cur_file, int(r[0]))
        e = lines.setdefault(key, dict(src=r[1].strip(), inst=0, smp=0, stalls={}))
        e["inst"] += int(d["Instructions Executed"] or 0)
        e["smp"] += int(d["# Samples"] or 0)
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "-", "0"):
                e["stalls"][k] = e["stalls"].get(k, 0) + int(v)
    tot_i = sum(e["inst"] for e in lines.values()) or 1
    tot_s = sum(e["smp"] for e in lines.values()) or 1
    print(f"total warp instructions {tot_i}  samples {tot_s}  (tables seen: {launches})")
    items = [(k, e) for k, e in lines.items() if args.file is None or args.file in (k[0] or "")]
    for (f, ln), e in sorted(items, key=lambda kv: -kv[1]["smp"])[:args.top]:
        top = sorted(e["stalls"].items(), key=lambda kv: -kv[1])[:2]
        st = " ".join(f"{k[6:]}={v}" for k, v in top)
        print(f"{100 * e['inst'] / tot_i:5.1f}% inst {100 * e['smp'] / tot_s:5.1f}% smp  {f.split('/')[-1]}:{ln:<4} {e['src'][:90]}   [{st}]")


if __name__ == "__main__":
    main()
