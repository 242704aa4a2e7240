"""Writes profiles/r2_ncu_traffic.json from `ncu --page raw --csv` exports (gpurun_out/<tag>_<kernel>_raw.csv): DRAM bytes read + written
and L2 bytes of ONE launch per kernel, tagged with the hash of the kernel sources they were captured from (bench.py refuses the
figures when the sources it timed hash differently).

    python tools/make_ncu_traffic.py <tag> k_spatial_merge k_temporal_merge ..."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def num(x):
    return float(x.replace(",", ""))


def main():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    tag, kernels = sys.argv[1], sys.argv[2:]
    out = {"kernel_sources_sha16": bench.kernel_sources_hash(), "captured_with": "ncu --set full --clock-control none, one launch after warm-up (tools/run_%s.sh)" % tag,
           "kernels": {}}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for k in kernels:
        rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", "%s_%s_raw.csv" % (tag, k)))))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

        def get(name):
            u, v = d[name]
            return num(v) * scale.get(u, 1.0)
        rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
        ent = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "duration_us": num(d["gpu__time_duration.sum"][1]) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}[d["gpu__time_duration.sum"][0]]}
        out["kernels"][k] = ent
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, json.dumps(out["kernels"]))


if __name__ == "__main__":
    main()
