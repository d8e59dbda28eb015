"""CPU: the C-ABI library loads, exports every symbol the header declares, and the ctypes structs match the C layout."""
import ctypes
import os
import re
import subprocess
import tempfile

from conftest import ROOT
from text2pos_cvpr2022_b200 import _lib

HEADER = os.path.join(ROOT, "include", "text2pos_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(t2p_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_bound_and_exported(lib):
    names = header_functions()
    assert len(names) >= 20
    assert sorted(_lib.PROTOTYPES.keys()) == names
    for n in names:
        assert hasattr(lib, n), f"{n} not exported by {_lib.LIB_PATH}"
    assert lib.t2p_version() >= 100


def test_struct_layouts_match_c():
    code = r'''
#include <stdio.h>
#include "text2pos_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(t2p_linear_desc), sizeof(t2p_pointnet2_desc), sizeof(t2p_objenc_desc),
         sizeof(t2p_cellagg_desc), sizeof(t2p_lstm_desc), sizeof(t2p_superglue_desc), sizeof(t2p_peers));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(code)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    py = [ctypes.sizeof(t) for t in (_lib.LinearDesc, _lib.PointNet2Desc, _lib.ObjEncDesc, _lib.CellAggDesc, _lib.LstmDesc, _lib.SuperGlueDesc, _lib.Peers)]
    assert sizes == py


def test_errors_are_loud_without_gpu_or_bad_args(lib):
    # argument validation happens before any CUDA call, so it works on the CPU box too
    rc = lib.t2p_retrieve_topk(None, None, 1, 1, 1, 1, 0, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.t2p_last_error()
    rc = lib.t2p_topk_merge(None, None, 1, 1, 1, 1, None, None, None)
    assert rc == -1
