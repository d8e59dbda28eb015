#!/usr/bin/env python
"""Build tools/bin/libexp_SCT.so: the library with %globaltimer trace points in retrieve_scan_tc_kernel (-DT2P_SCAN_TRACE; CTA
(0,0)); tools/diag_scan_trace.py runs one scan with it and prints the timeline.  The product library is not touched."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from text2pos_cvpr2022_b200 import build as b
b.build()
obj_dir = os.path.join(b.HERE, "build")
out_obj = os.path.join(ROOT, "tools", "bin", "retrieval_tc_trace.o")
os.makedirs(os.path.dirname(out_obj), exist_ok=True)
src = os.path.join(b.HERE, "csrc", "retrieval_tc.cu")
subprocess.check_call([b.NVCC, *b.ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-DT2P_SCAN_TRACE", "-c", src, "-o", out_obj])
objs = [os.path.join(obj_dir, os.path.basename(s)[:-3] + ".o") for s in b.sources() if not s.endswith("retrieval_tc.cu")]
out = os.path.join(ROOT, "tools", "bin", "libexp_SCT.so")
subprocess.check_call([b.NVCC, *b.ARCH_FLAGS, "-shared", "-o", out, out_obj, *objs])
print(out)
