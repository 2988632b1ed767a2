"""Top SASS instructions of a kernel by warp-stall samples, from an .ncu-rep taken with --import-source on:
    python tools/ncu_hot_sass.py gpurun_out/x.ncu-rep [top N] > profiles/x.hot.txt"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    samp = [int(r[ci["# Samples"]] or 0) for r in body]
    total = sum(samp) or 1
    print(f"kernel {name}: {len(body)} SASS instructions, {total} warp-stall samples")
    print("  share   cum   samples  shared-wavefronts(excess)  instruction")
    order = sorted(range(len(body)), key=lambda i: -samp[i])[:top]
    cum = 0
    for i in order:
        cum += samp[i]
        r = body[i]
        wv, ex = r[ci["L1 Wavefronts Shared"]], r[ci["L1 Wavefronts Shared Excessive"]]
        print(f"  {100.0 * samp[i] / total:5.1f}% {100.0 * cum / total:5.1f}% {samp[i]:8d}  {wv:>10s}({ex:>9s})  {r[ci['Source']].strip()[:90]}")
    # by opcode class
    cls = {}
    for r, n in zip(body, samp):
        op = r[ci["Source"]].strip().split()
        op = (op[1] if op and op[0].startswith("@") else (op[0] if op else "?")).split(".")[0]
        cls[op] = cls.get(op, 0) + n
    print("by opcode:", ", ".join(f"{k} {100.0 * v / total:.1f}%" for k, v in sorted(cls.items(), key=lambda kv: -kv[1])[:12]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
