"""Build libsmplk.so (or a -D variant of it) with ptxas -v and print the fused kernel's resource line.
Usage: python tools/build_variant.py [out.so] [-DNAME=VALUE ...]   (default: the product library)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smplk import _lib as L

args = sys.argv[1:]
out = L.LIB_PATH
if args and not args[0].startswith("-"):
    out = args.pop(0)
cmd = ["nvcc"] + L.NVCC_FLAGS + args + ["-Xptxas", "-v", "-o", out, os.path.join(L.CSRC, "smplk_api.cu")]
r = subprocess.run(cmd, capture_output=True, text=True)
lines = (r.stdout + r.stderr).splitlines()
for i, l in enumerate(lines):
    if "error" in l or ("warning" in l and "ptxas" not in l):
        print(l)
    if "Compiling" in l and any(k in l for k in ("blend_skin_fused_kernelILi20670", )):
        print(out, "|", " ".join(x.strip() for x in lines[i + 2:i + 4]))
sys.exit(r.returncode)
