"""Opcode evidence per kernel of libb200diff.so: counts of the tensor-core / TMA / TMEM / cluster SASS mnemonics (the ones
/opt/skills/guides/B200_PROFILING.md names) and the ten most frequent opcodes.  python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusers_b200", "_C", "libb200diff.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "UTMALDG", "UTMASTG", "UTMACCTL", "UBLKPF", "UTMAPF", "LDTM", "STTM", "UTCALLOC", "SYNCS", "UCGABAR",
       "ACQBULK", "ELECT", "MUFU", "FFMA2", "FADD2", "FMNMX3", "USETMAXREG", "HMMA", "LDGSTS")
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = collections.Counter()
        kern[m.group(1)] = cur
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(so)}: {len(kern)} kernels (sm_100a)")
for (name, c), dn in zip(kern.items(), demangle):
    if sum(c.values()) < 50:
        continue
    short = re.sub(r"\(.*", "", dn).replace("void b200::", "")
    keys = {k: c[k] for k in KEY if c[k]}
    print(f"{short[:70]:70s} {sum(c.values()):6d} instr  {keys}")
