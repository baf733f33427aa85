#!/usr/bin/env python
"""tools/make_traffic.py SUMMARY.json [windows] [landmarks] -> profiles/r02_traffic.json

bench.py's `roofline.traffic` = dram__bytes_read.sum + dram__bytes_write.sum per launch of each window kernel, taken from
the `ncu --set full` capture of tools/profile_driver.py (same batch as the bench's headline workload: 9472 windows, seed
1000, L ~ U{750..1250}); the file is stamped with the commit the capture was taken on."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(s):
    v, u = s.split()
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def main():
    d = json.load(open(sys.argv[1]))
    out = {"_source": f"ncu --set full capture of tools/profile_driver.py ({os.path.basename(sys.argv[1])}): "
                      "dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip(),
           "windows": int(sys.argv[2]) if len(sys.argv) > 2 else 9472}
    if len(sys.argv) > 3:
        out["landmarks"] = int(sys.argv[3])
    for k, v in d.items():
        name = k.replace("void ", "").split("<")[0]     # template instantiations: the first one profiled (the bench's) wins
        if "dram__bytes_read.sum" in v and name not in out:
            out[name] = int(to_bytes(v["dram__bytes_read.sum"]) + to_bytes(v["dram__bytes_write.sum"]))
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
