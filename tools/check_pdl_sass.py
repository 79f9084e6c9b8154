"""Static audit of the PDL kernels: list every global load that the compiler scheduled BEFORE griddepcontrol.wait (SASS
ACQBULK).  Only loads of data that is static during a solve may appear there (PDL rule).
    python tools/check_pdl_sass.py            # prints, per kernel, the pre-wait global loads with their source lines"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "admm_optim_b200", "libadmm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels = re.split(r"\n\s*Function : ", out)[1:]
bad = 0
for k in kernels:
    name = k.split("\n", 1)[0].strip()
    if "ACQBULK" not in k:
        continue
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem)
    lines = k.split("\n")
    first_wait = next(i for i, l in enumerate(lines) if "ACQBULK" in l)
    pre = [l for l in lines[:first_wait] if re.search(r"\b(LDG|LD\.E|ATOMG|STG|RED)\b", l)]
    kinds = sorted({re.search(r"(LDG[.\w]*|STG[.\w]*|ATOMG[.\w]*|RED[.\w]*)", l).group(1) for l in pre})
    print("%-60s pre-wait global ops: %3d  %s" % (dem[:60], len(pre), " ".join(kinds)))
    if any(re.search(r"\b(STG|ATOMG|RED)\b", l) for l in pre):
        bad += 1
        print("   !!! store/atomic before the dependency wait")
sys.exit(1 if bad else 0)
