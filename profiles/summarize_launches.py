"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-frame table.

    python profiles/summarize_launches.py gpurun_out/launches.csv [frame_index] > profiles/<name>.md

A frame = the launches from one stem7x7_kernel to the next.  ncu serialises launches and flushes caches, so
only the SHARES are meaningful, not the absolute times (B200_PROFILING.md).
"""
import collections
import csv
import sys


def short(n):
    n = n.replace("void ", "").replace("deft::", "")
    for eng in ("igemm_simt_kernel", "tc_igemm_kernel", "tc_ts_kernel"):
        if eng in n:
            ap = "Dcn" if "DcnAP" in n else ("Pair" if "PairAP" in n else "Conv")
            tail = n.split("AP, ")[1].split(">")[0] if "AP, " in n else "?"
            return f"{eng}<{ap},{tail}>"
    return n.split("(")[0].split("<")[0][:60]


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if l.startswith('"')]))
    stems = [i for i, r in enumerate(rows) if "stem7x7" in r["Kernel Name"]]
    if len(stems) < 2:
        sys.exit("need at least two frames in the capture")
    pairs = list(zip(stems[:-1], stems[1:]))
    a, b = pairs[which]
    frame = rows[a:b]
    tot = sum(float(r["Metric Value"]) for r in frame)
    agg = collections.OrderedDict()
    for r in frame:
        k = short(r["Kernel Name"])
        c, s = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, s + float(r["Metric Value"]))
    print(f"# launch list summary: {path}, frame {which} ({len(frame)} launches, {tot / 1e3:.1f} us serialised)\n")
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {s / 1e3:.1f} | {100 * s / tot:.1f}% |")
    gem = sum(s for k, (c, s) in agg.items() if "Dcn" in k)
    print(f"\nDCN gather-GEMM share of the frame: {100 * gem / tot:.1f}%")


if __name__ == "__main__":
    main()
