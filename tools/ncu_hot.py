#!/usr/bin/env python
"""Hot spots of one kernel from an .ncu-rep source page (warp-state samples per SASS region).

    python tools/ncu_hot.py gpurun_out/x.ncu-rep <kernel-regex> [region=25] [top=0]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    region = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # several launches are concatenated: keep the first block
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        print("no kernel matched")
        return
    end = starts[1] if len(starts) > 1 else len(rows)
    print(rows[starts[0]][1][:100])
    hdr = rows[starts[0] + 1]
    body = [r for r in rows[starts[0] + 2:end] if len(r) == len(hdr)]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall_cols = [(j, h[6:]) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in body) or 1
    print(f"samples {tot}, instructions {len(body)}")
    for k in range(0, len(body), region):
        blk = body[k:k + region]
        v = sum(int(r[si]) for r in blk)
        if v < tot * 0.01:
            continue
        st = collections.Counter()
        for r in blk:
            for j, name in stall_cols:
                if r[j] not in ("", "0"):
                    st[name] += int(r[j])
        top = ", ".join(f"{n} {100 * c / v:.0f}%" for n, c in st.most_common(3))
        hot = max(blk, key=lambda r: int(r[si]))
        print(f"{k:5d} {100 * v / tot:5.1f}%  exec {blk[0][ie]:>8s}  [{top}]  hot: {hot[src].strip()[:60]}")


if __name__ == "__main__":
    main()
