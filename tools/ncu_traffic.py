"""profiles/traffic.json: DRAM bytes per launch of both kernels, from `ncu --set full` captures of the CURRENT kernels.

usage: python tools/ncu_traffic.py config3:10000000=gpurun_out/prof_r2_config3.ncu-rep [more ...]
The table is keyed by bench.kernel_hash() (sha256 of regk_core.cuh + regk_kernels.cuh) so that bench.py reports
`roofline.traffic` only for the kernel sources the capture was taken from (null otherwise)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    mine = table.setdefault(bench.kernel_hash(), {})
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        entry = {}
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = d["Kernel Name"]
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(d[m].replace(",", "")) * SCALE[units[hdr.index(m)]]
            which = "path" if "path_kernel" in name else "json" if "json_kernel" in name else None
            if which and which not in entry:
                entry[which] = int(tot)
                entry[which + "_us"] = float(d["gpu__time_duration.sum"].replace(",", "")) * (
                    1e-3 if units[hdr.index("gpu__time_duration.sum")] in ("nsecond", "ns") else 1.0)
        mine[key] = entry
        print(key, entry)
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
