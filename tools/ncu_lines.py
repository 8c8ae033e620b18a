"""Per-source-line instruction / stall-sample shares of one kernel of an .ncu-rep.

usage: python tools/ncu_lines.py REPORT KERNEL_REGEX [min_share_pct]
(reads `ncu --page source --print-source cuda,sass --csv`; needs -lineinfo and --import-source on)
"""
import csv, subprocess, sys

def main():
    rep, rx = sys.argv[1], sys.argv[2]
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = ""
    agg = {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].rsplit("/", 1)[-1]
            continue
        if len(r) < 8 or not r[0] or r[2] != "-":
            continue
        try:
            key = (fname, int(r[0]), r[1][:90])
            s, i = int(r[4]), int(r[7])
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0])
        a[0] += s
        a[1] += i
    ti = sum(v[1] for v in agg.values()) or 1
    ts = sum(v[0] for v in agg.values()) or 1
    print(f"total warp instructions {ti}, stall samples {ts}")
    for (f, ln, src), (s, i) in sorted(agg.items()):
        if 100 * i / ti >= thr or 100 * s / ts >= thr:
            print(f"{f}:{ln:4d} {100*i/ti:5.1f}%inst {100*s/ts:5.1f}%smpl  {src}")

if __name__ == "__main__":
    main()
