#!/usr/bin/env python3
"""Refresh profiles/traffic.json (read by bench.py for `roofline.traffic`) from an `ncu --set full` capture.

    # on the GPU box (after the same command has exited 0 without ncu):
    ncu --set full --clock-control none --import-source on -k regex:"ib_(cn|vn)[0-9]*_n4" -s 12 -c 2 -f -o gpurun_out/prof_c1 \
        python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e
    # here:
    ncu -i gpurun_out/prof_c1.ncu-rep --page raw --csv > /tmp/raw_c1.csv
    python profiles/capture_traffic.py /tmp/raw_c1.csv c1 65536 "r02 final kernels"

The file records the sha256 of the kernel sources (bench.source_hash); bench.py reports `traffic: null` with a note when
the current sources differ from the ones the capture was taken on.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    raw, workload, frames, source = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    import bench
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw, ik, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def nbytes(r, i):
        return float(r[i].replace(",", "")) * scale[units[i]]

    best = max(data, key=lambda r: float(r[it].replace(",", "")))          # the dominant kernel = the longest launch
    kernel = best[ik].replace("void ", "").split("(")[0].replace("ibldpc::", "")
    kind = "ib_cn" if "ib_cn" in kernel or "kernel<0" in kernel else "ib_vn"
    path = os.path.join(ROOT, "profiles", "traffic.json")
    tj = json.load(open(path)) if os.path.exists(path) else {}
    tj[workload] = {"kernel": kind + " " + kernel, "bytes": nbytes(best, ir) + nbytes(best, iw), "bytes_read": nbytes(best, ir),
                    "bytes_written": nbytes(best, iw), "frames_per_launch": frames, "source": source,
                    "source_hash": bench.source_hash()}
    json.dump(tj, open(path, "w"), indent=1)
    print(json.dumps(tj[workload], indent=1))


if __name__ == "__main__":
    main()
