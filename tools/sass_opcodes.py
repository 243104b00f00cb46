"""cuobjdump opcode histogram per kernel of libmss_b200.so -> profiles/rNN_sass_opcodes.txt, so that tcgen05 (UTCHMMA,
LDTM/STTM), TMA (UTMALDG, UBLKCP) and the rest of the instruction mix are provable without the (git-ignored) .so."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "multishiftseg_b200", "libmss_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kern, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU", "FFMA", "DFMA", "DADD", "DMUL",
       "VOTE", "SHFL", "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR"]
with open(out, "w") as f:
    f.write(f"# SASS opcode histogram of multishiftseg_b200/libmss_b200.so (cuobjdump -sass), architectures: {', '.join(arch)}\n")
    f.write("# columns: total instructions, then the opcodes that identify the hardware path (0 omitted)\n")
    tot = collections.Counter()
    for k, h in hist.items():
        tot.update(h)
        cols = " ".join(f"{o}={h[o]}" for o in KEY if h[o])
        f.write(f"{k}\n    total={sum(h.values())} {cols}\n")
    f.write("\n# whole library\n    " + " ".join(f"{o}={tot[o]}" for o in KEY if tot[o]) + "\n")
print(open(out).read()[-700:])
