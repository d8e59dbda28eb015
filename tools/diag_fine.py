#!/usr/bin/env python
"""Diagnosis (GPU box): where the CUDA fine matcher and the oracle part ways on the pipeline test's matching-regime weights."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.nn.functional as F

import oracle
import pipeline_common as pc
from oracle import cells as ocells, superglue as osg, text as otext
from oracle.models import _pack
from text2pos_cvpr2022_b200 import _lib, pipeline_eval as pe, synthetic as syn
from text2pos_cvpr2022_b200.object_encoder import object_encoder_forward
from text2pos_cvpr2022_b200.pointnet2 import pointnet2_forward
from text2pos_cvpr2022_b200.object_encoder import obj_cell_start_from_offsets
from text2pos_cvpr2022_b200.superglue import superglue_forward

ds, loader = pc.scene(11)
args = pc.pipeline_args()
fm, sd = pc.fine_state_dict()
kw = fm.language_encoder.known_words
fm = fm.to("cuda")
retr = [[c.id for c in ds.all_cells[:5]]]
s = pe.TopKDataset(ds.all_poses[:1], ds.all_cells, retr, None, args)[0]
rgb, pos, ctr, col = _pack(s["objects"], s["object_points"])
cells = syn.pack_cells(s["objects"], s["object_points"]).to("cuda")
weights, desc = fm.t2p_packed()
with torch.no_grad():
    f2_ref = torch.cat([oracle.pointnet.pointnet2_features(sd, "object_encoder.pointnet.", r, p, True) for r, p in zip(rgb, pos)])
start = obj_cell_start_from_offsets(cells.cell_offsets.to("cuda"))
f2 = pointnet2_forward(weights, desc["pointnet"], cells.pos, cells.rgb, start, fm)
print("features2: max|ref|", float(f2_ref.abs().max()), "max err", float((f2.cpu() - f2_ref).abs().max()), "finite", bool(torch.isfinite(f2).all()))
rel = (f2.cpu() - f2_ref).abs().max() / f2_ref.abs().max()
print("features2 rel err", float(rel))
with torch.no_grad():
    obj_ref = ocells.object_encoder(sd, "object_encoder.", rgb, pos, ctr, col)
obj = object_encoder_forward(weights, desc["pointnet"], desc["objenc"], cells, fm)
print("obj emb: max|ref|", float(obj_ref.abs().max()), "max err", float((obj.cpu() - obj_ref).abs().max()))
on, orn = F.normalize(obj.cpu(), dim=-1), F.normalize(obj_ref, dim=-1)
print("obj emb normalised max err", float((on - orn).abs().max()))
hint_ref = torch.stack([otext.language_encoder(sd, "language_encoder.", *otext.tokenize(h, kw)) for h in s["hint_descriptions"]])
out = fm(s["objects"], s["hint_descriptions"], s["object_points"])
ref = osg.superglue_match_forward(sd, hint_ref, obj_ref.reshape(5, 16, 128), 6, 50)
print("matches equal", bool(np.array_equal(out.matches0.cpu().numpy(), ref["matches0"].numpy())))
print("P max err", float((out.P.cpu() - ref["P"]).abs().max()))
# SuperGlue alone on the ORACLE's normalised encodings
o = F.normalize(obj_ref.reshape(5, 16, 128), dim=-1).cuda()
h = F.normalize(hint_ref, dim=-1).cuda()
sg = superglue_forward(weights, desc["superglue"], o, h, fm, return_scores=True)
ref_sg = osg.superglue_forward(sd, "superglue.", o.cpu(), h.cpu(), 6, 50)
print("superglue on oracle inputs: scores max err", float((sg["scores"].cpu() - ref_sg["scores"]).abs().max()), "scores std",
      float(ref_sg["scores"].std()), "matches equal", bool(np.array_equal(sg["matches0"].cpu().numpy(), ref_sg["matches0"].numpy())))
from text2pos_cvpr2022_b200.modules import lstm_encode, tokenize
flat = [x for hh in s["hint_descriptions"] for x in hh]
tok, ln = tokenize(flat, kw)
he = lstm_encode(weights, desc["lstm"], torch.from_numpy(tok).cuda(), torch.from_numpy(ln).cuda(), True, fm)
print("hint enc normalised max err", float((he.cpu() - F.normalize(hint_ref.reshape(-1, 128), dim=-1)).abs().max()))
# per-layer debug of the pointnet: which set-abstraction level diverges first
out_dbg, dbg = pointnet2_forward(weights, desc["pointnet"], cells.pos, cells.rgb, start, fm, debug=True)
with torch.no_grad():
    ref_layers = oracle.pointnet.pointnet2_debug(sd, "object_encoder.pointnet.", rgb[0], pos[0], True) if hasattr(oracle.pointnet, "pointnet2_debug") else None
for l in range(3):
    x = dbg["x"][l]
    print(f"SA{l+1} out: max {float(x.abs().max()):.4g} finite {bool(torch.isfinite(x).all())}")
    if ref_layers is not None:
        n0 = rgb[0].shape[0]
        r = ref_layers["x"][l]
        print(f"   vs oracle (cell 0): max err {float((x[:n0].cpu() - r).abs().max()):.4g} rel {float((x[:n0].cpu() - r).abs().max() / r.abs().max()):.3g}")
