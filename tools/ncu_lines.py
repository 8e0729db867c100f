"""Join an ncu SASS-level source page (csv) with nvdisasm line info to get per-CUDA-line instruction / sample counts.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <object file .o> [launch index]"""
import collections, csv, re, subprocess, sys
rep, kre, obj = sys.argv[1:4]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kre.split(chr(73))[0]}", "-s", skip, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname = rows[0][1]
hdr = rows[1]
ai, si, ii, sa = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [(int(r[ai], 16), r[si].strip(), int(r[ii] or 0), int(r[sa] or 0)) for r in rows[2:] if len(r) > ii and r[ii].isdigit()]
base = sass[0][0]
# nvdisasm with line info for the matching function
import os, tempfile
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(obj)} >/dev/null 2>&1", shell=True)
cubs = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubs[0]], capture_output=True, text=True)
# pick the function whose demangled-ish name matches pieces of kname
want = re.sub(r"[^A-Za-z0-9_]", " ", kname).split()
cur_fn, cur_line, fn_lines = None, None, collections.defaultdict(dict)
for ln in dis.stdout.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
    if m:
        cur_fn = m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur_fn:
        fn_lines[cur_fn][int(m.group(1), 16)] = cur_line
# choose function with the most address matches
best = max(fn_lines, key=lambda f: (all(w in f for w in [kre.split("|")[0][:12]]), len(fn_lines[f]))) if fn_lines else None
cands = [f for f in fn_lines if kre.split("|")[0] in f]
n = len(sass)
cands = sorted(cands, key=lambda f: abs(len(fn_lines[f]) - n))
fn = cands[0] if cands else best
lines = fn_lines[fn]
agg = collections.defaultdict(lambda: [0, 0])
ti = ts = 0
for a, s, ni, ns in sass:
    key = lines.get(a - base, ("?", 0))
    agg[key][0] += ni; agg[key][1] += ns; ti += ni; ts += ns
print(f"kernel {kname[:100]}\nfunction {fn}\ninstructions executed {ti}  samples {ts}")
print("by line (top 40 by samples):")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f"  {k[0]:16s}:{k[1]:<5d} inst {100 * v[0] / max(ti, 1):5.1f}%  samples {100 * v[1] / max(ts, 1):5.1f}%")
