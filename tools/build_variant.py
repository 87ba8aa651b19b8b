"""Builds an experiment variant of the library: tools/build_variant.py NAME -DMACRO [-DMACRO2 ...]
-> gpurun_out/variants/libwfl_NAME.so (select with WFL_LIB=...).  Only attention.cu is recompiled with the macros."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wfl_asr_b200 import build as b
name, defs = sys.argv[1], sys.argv[2:]
b.build()
out_dir = os.path.join(os.path.dirname(b.PKG_DIR), "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, f"attention_{name}.o")
subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + defs + ["-c", os.path.join(b.CSRC, "attention.cu"), "-o", obj])
objs = [os.path.join(b.PKG_DIR, "build", s[:-3] + ".o") for s in b.SOURCES if s != "attention.cu" and os.path.exists(os.path.join(b.CSRC, s))]
so = os.path.join(out_dir, f"libwfl_{name}.so")
subprocess.check_call([b._nvcc(), "-shared", "-cudart", "shared", "-o", so, obj] + objs)
print(so)
