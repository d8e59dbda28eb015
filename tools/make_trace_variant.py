#!/usr/bin/env python
"""Build tools/bin/libexp_TR.so: the library with clock64 trace points in sa_edge_tc_kernel (per role, per work item, CTA 0), read
back with t2p_debug_trace (tools/diag_trace.py).  The product sources are not modified."""
import os, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "text2pos_cvpr2022_b200/csrc/sa_tc.cu")).read()
WHICH = sys.argv[1] if len(sys.argv) > 1 else "!DENSE && N == 128 && K == 128"  # which instantiation records
def rep(a, b, count=1):
    global src
    assert a in src, a[:60]
    src = src.replace(a, b, count)
rep("namespace t2p {\n\nusing namespace sm100;\n", """namespace t2p {

using namespace sm100;
__device__ long long sat_trace[16 * 256];
#define TR(ev) do { if ((%s) && blockIdx.x == 0 && blockIdx.y == 0 && it < 256 && lane == 0) sat_trace[(ev) * 256 + it] = clock64(); } while (0)
""" % WHICH)
rep("""        if ((it & 1) != which) continue;  // the other row-table warp's item
        const int buf = it % SAT_NTAB;
        mbar_wait(&bars->rows_empty[buf], (uint32_t)(((it / SAT_NTAB) & 1) ^ 1));
""", """        if ((it & 1) != which) continue;  // the other row-table warp's item
        const int buf = it % SAT_NTAB;
        mbar_wait(&bars->rows_empty[buf], (uint32_t)(((it / SAT_NTAB) & 1) ^ 1));
        TR(0);
""")
rep("""          rw->t_last = (e_base + SAT_ROWS >= E) ? 1 : 0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->rows_full[buf]);
""", """          rw->t_last = (e_base + SAT_ROWS >= E) ? 1 : 0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->rows_full[buf]);
        TR(1);
""")
rep("""      const int n_valid = rows[buf].n_valid;
      if (n_valid < 0) break;""", """      TR(2);
      const int n_valid = rows[buf].n_valid;
      if (n_valid < 0) break;""")
rep("""        mbar_wait(&bars->tmem_empty[acc], (uint32_t)(((nv >> 1) & 1) ^ 1));
        tc_fence_after_sync();""", """        mbar_wait(&bars->tmem_empty[acc], (uint32_t)(((nv >> 1) & 1) ^ 1));
        TR(3);
        tc_fence_after_sync();""")
rep("""          mbar_wait(&bars->full[stage], ph);
          tc_fence_after_sync();""", """          mbar_wait(&bars->full[stage], ph);
          if (kc == 0) TR(4);
          tc_fence_after_sync();""")
rep("""        ++nv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_empty[buf]);
    }
  } else if (warp < SAT_EPI_WARP0) {""", """        ++nv;
        TR(5);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_empty[buf]);
    }
  } else if (warp < SAT_EPI_WARP0) {""")
rep("""      const SatRows* rw = rows + buf;
      if (rw->n_valid < 0) break;
      if (rw->n_valid > 0) {
        const int t_row0""", """      const SatRows* rw = rows + buf;
      if (pw == 0) TR(6);
      if (rw->n_valid < 0) break;
      if (rw->n_valid > 0) {
        const int t_row0""")
rep("""          mbar_wait(&bars->empty[stage], ph ^ 1);
          uint8_t* st = stages + stage * STAGE_BYTES;""", """          mbar_wait(&bars->empty[stage], ph ^ 1);
          if (pw == 0 && kc == 0) TR(7);
          uint8_t* st = stages + stage * STAGE_BYTES;""")
rep("""          sat_fence_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->full[stage]);""", """          if (pw == 0 && kc == 0) TR(8);
          sat_fence_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->full[stage]);
          if (pw == 0 && kc == NKC - 1) TR(9);""")
rep("""        const int acc = nv & 1;
        mbar_wait(&bars->tmem_full[acc], (uint32_t)((nv >> 1) & 1));
        tc_fence_after_sync();
        for (int cc = eh;""", """        const int acc = nv & 1;
        if (warp == SAT_EPI_WARP0) TR(10);
        mbar_wait(&bars->tmem_full[acc], (uint32_t)((nv >> 1) & 1));
        if (warp == SAT_EPI_WARP0) TR(11);
        tc_fence_after_sync();
        for (int cc = eh;""")
rep("""        if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        ++nv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_empty[buf]);
    }
  }
  tc_fence_before_sync();""", """        if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        if (warp == SAT_EPI_WARP0) TR(12);
        ++nv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_empty[buf]);
    }
  }
  tc_fence_before_sync();""")
src = src.rstrip() + """
extern "C" int t2p_debug_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, t2p::sat_trace, sizeof(long long) * 16 * 256);
}
"""
os.makedirs("/tmp/exp", exist_ok=True)
tmp = os.path.join(ROOT, "text2pos_cvpr2022_b200/csrc/_sa_tc_trace_tmp.cu")
open(tmp, "w").write(src)
try:
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                           "-c", tmp, "-o", "/tmp/exp/sa_tc_TR.o"])
finally:
    os.remove(tmp)
objs = [o for o in glob.glob(os.path.join(ROOT, "text2pos_cvpr2022_b200/build/*.o")) if not o.endswith("/sa_tc.o")]
os.makedirs(os.path.join(ROOT, "tools/bin"), exist_ok=True)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", os.path.join(ROOT, "tools/bin/libexp_TR.so"), *objs, "/tmp/exp/sa_tc_TR.o"])
print("built tools/bin/libexp_TR.so, tracing", WHICH)
