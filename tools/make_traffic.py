"""profiles/traffic.json from the ncu captures of tools/gpu_final_evidence.sh (gpurun_out/final/):
DRAM bytes per unit (LP, or launch of a pivot prefix) of each kernel, keyed by the hash of the kernel
sources (and of the library binary) the captures were taken on.  bench.py reports roofline.traffic only
when the loaded build comes from the same sources."""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "final")
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None   # capture names to refresh; others are kept
old = {}
try:
    for e in json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["captures"]:
        old[e["tag"]] = e
except Exception:
    pass
sha = open(os.path.join(src, "lib_sha16.txt")).read().strip()
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (the source hash; run this while the sources are the ones the captures were built from)
def dram(name):
    rows = list(csv.reader(open(os.path.join(src, "ncu_%s_raw.csv" % name))))
    rows = [r for r in rows if len(r) > 10]
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        i = hdr.index(key)
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[units[i]]
        out[key] = float(vals[i].replace(",", "")) * mult
    return out
caps = []
for name, tag, kernel in (("warp_c5", "c5", "dz_batch_kernel<1,true,8,6> (warp per LP, config-5 LPs)"),
                          ("warp_c2", "c2", "dz_batch_kernel<1,true,4,3> (warp per LP, config-2 LPs)"),
                          ("core_c5", "c5-core", "dz_core_kernel<6,256> (coupled core on chip, 2 CTAs/SM)"),
                          ("fast_c2", "c2_fast", "dz_fast_kernel<128> (opt-in fast numerics, config-2 LPs)"),
                          ("fast_c5", "c5_fast", "dz_fast_kernel<512> (opt-in fast numerics, config-5 LPs)"),
                          ("fast_c5_dmma", "c5_fast_dmma", "dz_fast_kernel<512>, blocked elimination on the FP64 tensor cores (worker_warps=2)"),
                          ("grid_c4", "c4", "dz_grid_kernel (config 4, 60-pivot prefix)"),
                          ("grid_c3", "c3", "dz_grid_kernel (config 3, 60-pivot prefix)")):
    if only is not None and name not in only:
        if tag in old:
            caps.append(old[tag])
        continue
    try:
        d = dram(name)
        log = open(os.path.join(src, "one_%s.log" % name)).read()
    except Exception as e:
        print("skip", name, e)
        continue
    m = re.search(r" B (\d+) ", log) or re.search(r"prefix (\d+) ", log)
    n = int(m.group(1))
    total = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    single = name.startswith("grid")
    caps.append({"tag": tag, "kernel": kernel, "lib_sha16": sha, "src_sha16": bench._src_sha16(tag), "units_in_capture": n,
                 "unit": "launch of a %d-pivot prefix" % n if single else "LP",
                 "dram_bytes_read": d["dram__bytes_read.sum"], "dram_bytes_write": d["dram__bytes_write.sum"],
                 "dram_bytes_per_unit": total if single else total / n,
                 "prefix": n if single else None,
                 "kernel_ms_under_ncu": d["gpu__time_duration.sum"] * 1e3})
json.dump({"captures": caps, "note": "ncu --set full --clock-control none, one launch each (tools/gpu_final_evidence.sh)"},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(caps, indent=1))
