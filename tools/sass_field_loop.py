"""Prints the SASS of ms_load_kernel's field loop (the basic blocks around I2F.F64) of an object / library, and its size."""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?) ;", out)]
k = next(i for i, (_, t) in enumerate(ins) if "I2F.F64.U32" in t)
# the loop: the backward branch after k and its target
for j in range(k, len(ins)):
    m = re.search(r"BRA (?:P\d, )?(0x[0-9a-f]+)", ins[j][1])
    if m and int(m.group(1), 16) < ins[k][0]:
        lo = int(m.group(1), 16); hi = ins[j][0]; break
body = [(a, t) for a, t in ins if lo <= a <= hi]
if "-v" in sys.argv:
    for a, t in body: print(f"{a:05x} {t}")
print(len(body), "instructions in the loop body (incl. cold blocks inside)")
