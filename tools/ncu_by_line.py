"""Aggregate an ncu report's per-instruction samples by CUDA source line.

usage: ncu_by_line.py report.ncu-rep lib.so 'kernel-substring' [top] [source.cu (default dz_kernel.cu)]
Needs the .so the report was captured from (built with -lineinfo).
Lines of inlined CUDA headers (shuffles, __ldg ...) are charged to the nearest
preceding dz_kernel.cu line; a per-function summary follows the per-line table.
"""
import csv, io, os, re, subprocess, sys, tempfile

rep, so, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
SRC = sys.argv[5] if len(sys.argv) > 5 else "dz_kernel.cu"
STEM = SRC.rsplit(".", 1)[0]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if STEM + ".sm" not in f and STEM not in f:
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, want = None, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            want = pat in m.group(1)
            continue
        if not want:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if m.group(1).endswith(SRC):
                cur_line = int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur_line
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
tot_s = tot_i = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    line = line_of.get(off)
    a = agg.setdefault(line, dict(samples=0, instr=0, stalls={}))
    s = int(r[ix["# Samples"]] or 0)
    n = int(r[ix["Instructions Executed"]] or 0)
    a["samples"] += s; a["instr"] += n
    tot_s += s; tot_i += n
    for h in stall_cols:
        v = int(r[ix[h]] or 0)
        if v:
            a["stalls"][h] = a["stalls"].get(h, 0) + v
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dantzig_b200", "csrc", SRC)).read().splitlines()
print(kname, "samples", tot_s, "warp-instr", tot_i)
for line, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
    st = " ".join("%s=%.0f%%" % (k.replace("stall_", ""), 100 * v / max(a["samples"], 1)) for k, v in st)
    text = src[line - 1].strip()[:70] if line and line <= len(src) else "?"
    print("%5.1f%% smp %5.1f%% ins  L%-4s %-70s %s" % (100 * a["samples"] / tot_s, 100 * a["instr"] / tot_i, line, text, st))

# per-function summary: a line belongs to the last function header above it
starts = [(i + 1, m.group(1)) for i, l in enumerate(src)
          for m in [re.match(r"(?:__device__ __forceinline__|__global__|static __device__)\s+[\w:<> ]*?\b(\w+)\(", l.strip())
                    or re.match(r"(dz_batch_kernel|dz_core_kernel)\(", l.strip())] if m]
fagg = {}
for line, a in agg.items():
    name = "?"
    for ln0, nm in starts:
        if line and ln0 <= line:
            name = nm
    f = fagg.setdefault(name, [0, 0])
    f[0] += a["samples"]; f[1] += a["instr"]
print("\nby function (inlined into the kernel):")
for name, (sm, ins) in sorted(fagg.items(), key=lambda kv: -kv[1][0]):
    print("%5.1f%% smp %5.1f%% ins  %s" % (100 * sm / tot_s, 100 * ins / tot_i, name))
