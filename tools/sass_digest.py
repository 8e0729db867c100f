"""profiles/sass_digest.txt: SASS evidence per kernel (cuobjdump -sass of the sm_100a objects in csrc/build/).
UBLKCP = cp.async.bulk (TMA bulk copies), SYNCS = mbarrier, UCGABAR = cluster barrier, MAPA / .CLUSTER = DSMEM addressing,
MATCH = match.any, DFMA/DADD/DMUL = fp64 pipe, ATOM/RED = global atomics, ATOMS = shared atomics."""
import collections, re, subprocess, sys
from pathlib import Path
build = Path(__file__).resolve().parents[1] / "noetic-slam_b200" / "csrc" / "build"
pats = ["UBLKCP", "SYNCS", "UCGABAR", "MAPA", "MATCH", "DFMA", "DADD", "DMUL", "MUFU.RCP64H", "F2F.F64.F32", "ATOMG|ATOM\\.|RED\\.", "ATOMS", "SHFL", "REDUX", "VIMNMX", "FMNMX",
        "LDG", "STG", "LDS", "STS", "BAR.SYNC", "CCTL", "HMMA|UTCHMMA|UTCQMMA|UTCMMA"]
ver = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2]
print(f"# cuobjdump -sass digest of noetic-slam_b200/csrc/build/*.o ({ver}; -gencode arch=compute_100a,code=sm_100a)")
print("# columns: instructions, then the count of every mnemonic class that occurs in the kernel")
for o in ["api.o", "index.o", "radix_sort.o", "knn.o", "covariance.o", "linearize.o", "keyframe.o", "filters.o"]:
    f = build / o
    if not f.exists():
        continue
    sass = subprocess.run(["cuobjdump", "-sass", str(f)], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1); kernels[cur] = collections.Counter(); continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            kernels[cur]["n"] += 1
            for p in pats:
                if re.search(p, ln):
                    kernels[cur][p] += 1
    names = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"\n## {o}: {len(kernels)} kernels, {tot['n']} instructions; " + ", ".join(f"{p.split('|')[0]} {tot[p]}" for p in pats if tot[p]))
    for (k, c), nm in zip(kernels.items(), names):
        nm = re.sub(r"\(anonymous namespace\)::|ngicp::|<unnamed>::", "", nm)
        nm = re.sub(r"\((?!int\)|bool\)).*$", "", nm).replace("void ", "").replace("(int)", "").replace("(bool)", "")
        print(f"  {nm[:70]:70s} {c['n']:6d}  " + " ".join(f"{p.split('|')[0]}={c[p]}" for p in pats[:12] if c[p]))
