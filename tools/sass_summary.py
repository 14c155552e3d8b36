"""Opcode summary of the shipped library's SASS: per kernel, the Blackwell-native instructions (tcgen05 MMAs = UTCHMMA[.2CTA],
TMEM loads = LDTM, TMA loads = UTMALDG, tcgen05 commits = UTCBAR, mbarrier ops = SYNCS) and the widths of its global memory
accesses.  `make -C unet-phasegen_b200/csrc sass-summary` writes profiles/r02_sass_opcodes.txt (the .so itself is git-ignored)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "unet-phasegen_b200/csrc/libphasegen.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = {}
cur = None
counts = collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    if base in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "SYNCS"):
        counts[cur][op if base.startswith("UTC") else base] += 1
    elif base in ("STG", "LDG"):
        w = "128" if ".128" in op else "64" if ".64" in op else "16" if ".U16" in op or ".S16" in op else "8" if ".U8" in op else "32"
        counts[cur][f"{base}.{w}"] += 1
    counts[cur]["instructions"] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
for mangled, pretty in zip(counts, dem):
    pretty = re.sub(r"\(.*", "", pretty).replace("void ", "")
    c = counts[mangled]
    keys = sorted(k for k in c if k != "instructions")
    print(f"{pretty:45s} {c['instructions']:6d} instr  " + "  ".join(f"{k}={c[k]}" for k in keys))
