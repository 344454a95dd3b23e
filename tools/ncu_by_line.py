"""Attribute an ncu report's per-instruction counts to source lines: joins `ncu --page source --print-source sass`
(instruction order) with `nvdisasm -g` of the same cubin (line info).  usage: ncu_by_line.py report.ncu-rep cubin mangled_name [top]"""
import csv, io, os, re, subprocess, sys
rep, cubin, fun = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], ("?", 0), False
for l in dis:
    if l.startswith(".text."):
        on = l.strip() == f".text.{fun}:"
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, it, iss = (hdr.index(k) for k in ("Instructions Executed", "Thread Instructions Executed", "# Samples"))
data = [r for r in rows[2:] if len(r) > it and r[ia].isdigit()]
n = min(len(data), len(lines))
print(f"sass instructions: ncu {len(data)}, nvdisasm {len(lines)}")
agg = {}
for i in range(n):
    a = agg.setdefault(lines[i], [0, 0, 0])
    a[0] += int(data[i][ia]); a[1] += int(data[i][it]); a[2] += int(data[i][iss])
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
srcs = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2 if "--by-samples" in sys.argv else 0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open(os.path.join(os.environ.get("PRT_SRC_DIR", "/root/repo/physics-based-ray-tracing_b200/csrc"), f)).read().splitlines()
        except OSError:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print(f"{a[0] / tot * 100:5.2f}% inst {a[2] / max(tots, 1) * 100:5.2f}% smp thr={a[1] / max(a[0], 1):4.1f}  {f}:{ln}  {text}")
