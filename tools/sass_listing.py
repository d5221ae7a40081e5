"""SASS listing of one kernel of the built library with an opcode histogram on top (evidence that can be read without a GPU):
    python tools/sass_listing.py KERNEL_SUBSTRING [TEMPLATE_ARGS_SUBSTRING] > profiles/r02_sass_<kernel>.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "optical-flow-2d-data-generation_b200", "csrc", "libofdg.so")
want, targ = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
sect, take, fn = [], False, None
for l in out.splitlines():
    if "Function :" in l:
        take = want in l and targ in l and fn is None
        if take:
            fn = l.split("Function :")[1].strip()
    elif take and ".........." in l and sect:
        take = False
    if take:
        sect.append(l)
ins = [l.split("*/", 1)[1].split(";")[0].strip() for l in sect if re.match(r"\s*/\*[0-9a-f]{4}\*/", l)]
ops = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for t in ins if t)
demangled = subprocess.run(["c++filt", fn or ""], capture_output=True, text=True).stdout.strip()
print("# kernel:", demangled)
print("# sm_100a SASS from", os.path.relpath(lib, ROOT), "(cuobjdump -sass);", len(ins), "instructions (static)")
print("# opcode histogram:", ", ".join(f"{k} {v}" for k, v in ops.most_common()))
print("# local-memory (spill) instructions: LDL", ops.get("LDL", 0), "STL", ops.get("STL", 0),
      "| global: LDG", ops.get("LDG", 0), "STG", ops.get("STG", 0), "| shared: LDS", ops.get("LDS", 0), "STS", ops.get("STS", 0),
      "| bulk async copy (UBLKCP)", ops.get("UBLKCP", 0), "| DP: DADD", ops.get("DADD", 0), "DMUL", ops.get("DMUL", 0), "DFMA", ops.get("DFMA", 0))
# the listing itself: address + instruction (encodings dropped)
for l in sect:
    m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m:
        print(m.group(1), m.group(2).strip())
    elif l.strip().startswith(".L_") or "Function :" in l:
        print(l.strip())
