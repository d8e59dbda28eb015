#!/usr/bin/env python
"""Build tools/bin/libexp_GTR.so: the library with %globaltimer trace points in superglue_tc_kernel (-DT2P_SGT_TRACE; CTA 0, steps
20..23); tools/diag_sgt_trace.py prints the per-group timeline.  The product library is not touched."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from text2pos_cvpr2022_b200 import build as b
b.build()
obj_dir = os.path.join(b.HERE, "build")
out_obj = os.path.join(ROOT, "tools", "bin", "superglue_tc_trace.o")
os.makedirs(os.path.dirname(out_obj), exist_ok=True)
src = os.path.join(b.HERE, "csrc", "superglue_tc.cu")
subprocess.check_call([b.NVCC, *b.ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-DT2P_SGT_TRACE", *sys.argv[1:], "-c", src, "-o", out_obj])
objs = [os.path.join(obj_dir, os.path.basename(s)[:-3] + ".o") for s in b.sources() if not s.endswith("superglue_tc.cu")]
out = os.path.join(ROOT, "tools", "bin", "libexp_GTR.so")
subprocess.check_call([b.NVCC, *b.ARCH_FLAGS, "-shared", "-o", out, out_obj, *objs])
print(out)
