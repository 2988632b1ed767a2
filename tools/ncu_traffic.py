"""Write / update profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`) from an `ncu --set full` capture:

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep <kernel substring> <units per launch> [key]

`units per launch`: how many roofline "launches" one captured kernel launch stands for (the persistent greedy-MI kernels
run n_picks iterations per CUDA launch and bench.py accounts one iteration as one launch; 1 for ordinary kernels).
The entry holds dram__bytes_read.sum + dram__bytes_write.sum per unit, the capture's file name and the commit."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path, substr, units, key=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, unit_row = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    r_col, w_col = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, n = 0.0, 0
    for r in rows[2:]:
        if substr in r[name_col]:
            total += float(r[r_col]) * scale[unit_row[r_col]] + float(r[w_col]) * scale[unit_row[w_col]]
            n += 1
    if n == 0:
        raise SystemExit("no kernel matching %r in %s" % (substr, path))
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    dst = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    data = json.load(open(dst)) if os.path.exists(dst) else {}
    data[key or substr] = {"dram_bytes_per_launch": total / n / float(units), "captured_launches": n,
                           "units_per_captured_launch": float(units),
                           "source": "ncu --set full, %s, tree at %s" % (os.path.basename(path), commit)}
    json.dump(data, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[key or substr]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else None)
