"""Builds an experiment variant of the library: tools/build_variant.py NAME [--src FILE.cu] -DMACRO [-DMACRO2 ...]
-> variants/libwfl_NAME.so (select with WFL_LIB=...).  Only FILE.cu (default attention.cu) is recompiled with the macros."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wfl_asr_b200 import build as b
args = sys.argv[1:]
name = args.pop(0)
src = "attention.cu"
if args and args[0] == "--src":
    src = args[1]
    args = args[2:]
defs = args
b.build()
out_dir = os.path.join(os.path.dirname(b.PKG_DIR), "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, f"{src[:-3]}_{name}.o")
subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + defs + ["-c", os.path.join(b.CSRC, src), "-o", obj])
objs = [os.path.join(b.PKG_DIR, "build", s[:-3] + ".o") for s in b.SOURCES if s != src and os.path.exists(os.path.join(b.CSRC, s))]
so = os.path.join(out_dir, f"libwfl_{name}.so")
subprocess.check_call([b._nvcc(), "-shared", "-cudart", "shared", "-o", so, obj] + objs)
print(so)
