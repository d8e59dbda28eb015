"""Coarse-to-fine evaluation pipeline over the B200 modules (BASELINE config 5): the reference's
``evaluation/pipeline.py`` ``run_coarse`` (``:37-137``) and ``run_fine`` (``:171-279``), its top-k dataset
(``dataloading/kitti360pose/eval.py:117-189``) and ``calc_sample_accuracies`` (``evaluation/utils.py:31-54``).

Two ways through the fine stage:

* ``run_fine`` -- the drop-in path: one ``model(objects, hints, object_points)`` call per query over its ``max(top_k)``
  retrieved cells, exactly the reference's loop; works with any module that has the reference's ``forward`` signature.
* ``run_fine_cached`` -- the B200 path: the query-independent object encodings of every cell
  are computed ONCE (``SuperGlueMatch.encode_cells``), the fine stage per batch of queries is hint LSTM + gather + SuperGlue
  head + offset MLP + pose head (``forward_cached``), and the thresholded accuracies are reduced on the device
  (``t2p_pose_accuracy``).  Same results as ``run_fine`` given the same padding objects and resampled points.

What differs from the reference on purpose: padding objects and ``FixedPoints`` resampling are SEEDED per cell (the
reference draws them from numpy's global RNG per ``__getitem__``, so its own output changes run to run); pass
``padding_factory`` / a PyG-style ``transform`` to restore the reference's behaviour (used by the CPU test that compares
this file with the unmodified reference functions).

The data side is duck-typed like the reference: ``dataloader.dataset`` has ``all_poses`` (``pose_w``, ``cell_id``,
``descriptions`` with ``direction / object_color_text / object_label`` or plain hint strings) and ``all_cells``
(``id``, ``objects``, ``bbox_w``, ``cell_size``).
"""
import copy
import time
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import synthetic as syn


# ----------------------------------------------------------------------------------------------------------------
# shared helpers
# ----------------------------------------------------------------------------------------------------------------
def calc_sample_accuracies(pose, top_cells, pos_in_cells, top_k, threshs):
    """``evaluation/utils.py:31-54``: world-frame prediction per retrieved cell, distance to the ground-truth pose, cells of
    other scenes discarded, hit if any of the first k distances is within the threshold."""
    pose_w = pose.pose_w
    assert len(top_cells) == max(top_k) == len(pos_in_cells)
    pred_w = np.array([top_cells[i].bbox_w[0:2] + pos_in_cells[i, :] * top_cells[i].cell_size for i in range(len(top_cells))])
    dists = np.linalg.norm(pose_w[0:2] - pred_w, axis=1)
    pose_scene = pose.cell_id.split("_")[0]
    cell_scenes = np.array([cell.id.split("_")[0] for cell in top_cells])
    dists[pose_scene != cell_scenes] = np.inf
    return {k: {t: np.min(dists[0:k]) <= t for t in threshs} for k in top_k}


def create_hint_description(pose) -> List[str]:
    """``Kitti360BaseDataset.create_hint_description`` (dataloading/kitti360pose/base.py:57-66)."""
    hints = []
    for d in pose.descriptions:
        if isinstance(d, str):
            hints.append(d)
        else:
            hints.append(f"The pose is {d.direction} of a {d.object_color_text} {d.object_label}.")
    return hints


def _mean_accuracies(acc, top_k, threshs):
    return {k: {t: np.mean(acc[k][t]) for t in threshs} for k in top_k}


def padded_objects(cell, pad_size: int, padding_factory: Callable[[], object]) -> list:
    """Cut / pad the object list of ``cell`` to ``pad_size`` (dataloading/kitti360pose/eval.py:146-157).  Returns a new list;
    the cell itself is not modified."""
    objects = list(cell.objects[0:pad_size])
    while len(objects) < pad_size:
        objects.append(padding_factory())
    return objects


def seeded_padding_factory(seed: int, cell_id: str) -> Callable[[], object]:
    """Padding objects of a cell from a generator seeded with (seed, cell id): the same cell always gets the same padding."""
    rng = np.random.default_rng([int(seed)] + [ord(c) for c in str(cell_id)])
    return lambda: syn.SynthObject3d.create_padding(rng)


def _points_of(objects, transform, seed: int, obj_id0: int):
    """``batch_object_points(objects, transform)`` (dataloading/kitti360pose/utils.py:89-110).  ``transform`` None:
    FixedPoints(256) + NormalizeScale with the counter-based sampling indices of the device data path (object ``i`` is global
    object ``obj_id0 + i``, so the host path and ``CellStore.batch_object_points`` resample identically); a callable: the
    reference's per-object PyG-style transform."""
    if transform is None:
        return syn.batch_object_points_idx(objects, seed, obj_id0)
    xs, ps = [], []
    for obj in objects:
        d = transform(_PointData(torch.tensor(obj.rgb, dtype=torch.float), torch.tensor(obj.xyz, dtype=torch.float)))
        xs.append(d.x)
        ps.append(d.pos)
    batch = torch.cat([torch.full((p.shape[0],), i, dtype=torch.long) for i, p in enumerate(ps)])
    return syn.PointBatch(torch.cat(xs), torch.cat(ps), batch)


class _PointData:
    """The ``Data(x=rgb, pos=xyz)`` container a PyG-style transform mutates (``num_nodes``, item access, iteration)."""

    def __init__(self, x, pos):
        self.x, self.pos = x, pos

    @property
    def num_nodes(self):
        return int(self.pos.shape[0])

    def keys(self):
        return ["x", "pos"]

    def __iter__(self):
        yield "x", self.x
        yield "pos", self.pos

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)


class TopKDataset:
    """``Kitti360TopKDataset`` (dataloading/kitti360pose/eval.py:117-189): item ``i`` = pose ``i`` against each of its
    ``max(top_k)`` retrieved cells, objects cut / padded to ``args.pad_size``."""

    def __init__(self, poses, cells, retrievals, transform, args, seed: int = 0,
                 padding_factory: Optional[Callable[[], object]] = None):
        self.poses, self.retrievals = poses, retrievals
        assert len(poses) == len(retrievals)
        assert len(retrievals[0]) == max(args.top_k), "Retrievals where not trimmed to max(top_k)"
        self.cells_dict = {cell.id: cell for cell in cells}
        assert len(self.cells_dict) == len(cells), "Cell-IDs are not unique"
        self.cell_number = {cell.id: i for i, cell in enumerate(cells)}
        self.transform, self.args, self.seed, self.padding_factory = transform, args, seed, padding_factory

    def load_pose_and_cell(self, pose, cell):
        factory = self.padding_factory or seeded_padding_factory(self.seed, cell.id)
        padded = copy.copy(cell)  # the reference deep-copies the cell and pads its object list in place
        padded.objects = padded_objects(cell, self.args.pad_size, factory)
        return {
            "poses": pose,
            "objects": padded.objects,
            "object_points": _points_of(padded.objects, self.transform, self.seed, self.cell_number[cell.id] * self.args.pad_size),
            "hint_descriptions": create_hint_description(pose),
            "cells": padded,
        }

    def __getitem__(self, idx):
        pose = self.poses[idx]
        data = [self.load_pose_and_cell(pose, self.cells_dict[cid]) for cid in self.retrievals[idx]]
        return {k: [d[k] for d in data] for k in data[0]}

    def __len__(self):
        return len(self.poses)


# ----------------------------------------------------------------------------------------------------------------
# coarse stage
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def run_coarse(model, dataloader, args, eval_epoch_fn: Optional[Callable] = None):
    """``evaluation/pipeline.py:37-137`` (the model-driven branch and the ``coarse_oracle`` branch): top cells per pose and the
    cell-centre accuracies ``{k: {thresh: acc}}``."""
    from .coarse_eval import eval_epoch

    model.eval()
    dataset = dataloader.dataset
    all_cells_dict = {cell.id: cell for cell in dataset.all_cells}
    if getattr(args, "coarse_oracle", False):
        max_k = max(args.top_k)
        retrievals = [[pose.cell_id for _ in range(max_k)] for pose in dataset.all_poses]
    else:
        _, _, retrievals = (eval_epoch_fn or eval_epoch)(model, dataloader, args)
        retrievals = [retrievals[idx] for idx in range(len(retrievals))]
        assert len(retrievals) == len(dataset.all_poses)

    accuracies = {k: {t: [] for t in args.threshs} for k in args.top_k}
    for i_sample in range(len(retrievals)):
        pose = dataset.all_poses[i_sample]
        top_cells = [all_cells_dict[cell_id] for cell_id in retrievals[i_sample]]
        pos_in_cells = 0.5 * np.ones((len(top_cells), 2))  # predict the cell centres
        accs = calc_sample_accuracies(pose, top_cells, pos_in_cells, args.top_k, args.threshs)
        for k in args.top_k:
            for t in args.threshs:
                accuracies[k][t].append(accs[k][t])
    return retrievals, _mean_accuracies(accuracies, args.top_k, args.threshs)


# ----------------------------------------------------------------------------------------------------------------
# fine stage, drop-in path
# ----------------------------------------------------------------------------------------------------------------
def get_pos_in_cell(objects, matches0, offsets):
    from .superglue_matcher import get_pos_in_cell as f

    return f(objects, matches0, offsets)


@torch.no_grad()
def run_fine(model, retrievals, dataloader, args, transform=None, seed: int = 0,
             padding_factory: Optional[Callable[[], object]] = None, return_details: bool = False):
    """``evaluation/pipeline.py:171-279``: ``(accuracies_mean, accuracies_offset, accuracies_mean_conf)``; with
    ``return_details`` also a dict of the per-query matches / offsets / confidences / in-cell positions."""
    dataset = dataloader.dataset
    dataset_topk = TopKDataset(dataset.all_poses, dataset.all_cells, retrievals, transform, args, seed, padding_factory)
    num_samples = max(args.top_k)

    matches, offsets, confidences, cell_ids, poses_w, padded_cells = [], [], [], [], [], []
    for i_sample in range(len(dataset_topk)):
        sample = dataset_topk[i_sample]
        output = model(sample["objects"], sample["hint_descriptions"], sample["object_points"])
        out_matches = output.matches0.detach().cpu().numpy()
        assert out_matches.ndim == 2
        matches.append(out_matches)
        offsets.append(output.offsets.detach().cpu().numpy())
        confs = np.sum(out_matches >= 0, axis=1)
        assert len(confs) == num_samples
        confidences.append(confs)
        cell_ids.append([cell.id for cell in sample["cells"]])
        poses_w.append(sample["poses"][0].pose_w)
        padded_cells.append(sample["cells"])
    assert len(matches) == len(offsets) == len(retrievals)

    all_cells_dict = {cell.id: cell for cell in dataset.all_cells}
    acc_mean = {k: {t: [] for t in args.threshs} for k in args.top_k}
    acc_offset = {k: {t: [] for t in args.threshs} for k in args.top_k}
    acc_mean_conf = {1: {t: [] for t in args.threshs}}
    pos_mean_all, pos_off_all = [], []
    for i_sample in range(len(retrievals)):
        pose = dataset.all_poses[i_sample]
        top_cells = [all_cells_dict[cell_id] for cell_id in retrievals[i_sample]]
        assert np.all(np.array([cell.id for cell in top_cells]) == np.array(cell_ids[i_sample]))
        assert np.allclose(pose.pose_w, poses_w[i_sample])
        pos_mean, pos_off = [], []
        for i_cell in range(len(top_cells)):
            # the matcher may have matched a padding object: use the padded object list it saw (the reference re-pads the
            # cell with fresh random padding objects here, pipeline.py:231-236; theirs differ by < 1e-3 cell units)
            objs = padded_cells[i_sample][i_cell].objects
            m, o = matches[i_sample][i_cell], offsets[i_sample][i_cell]
            pos_mean.append(get_pos_in_cell(objs, m, np.zeros_like(o)))
            pos_off.append(get_pos_in_cell(objs, m, o))
        pos_mean, pos_off = np.array(pos_mean), np.array(pos_off)
        pos_mean_all.append(pos_mean)
        pos_off_all.append(pos_off)
        a_mean = calc_sample_accuracies(pose, top_cells, pos_mean, args.top_k, args.threshs)
        a_off = calc_sample_accuracies(pose, top_cells, pos_off, args.top_k, args.threshs)
        ci = int(np.argmax(confidences[i_sample]))
        a_conf = calc_sample_accuracies(pose, top_cells[ci:ci + 1], pos_mean[ci:ci + 1], top_k=[1], threshs=args.threshs)
        for k in args.top_k:
            for t in args.threshs:
                acc_mean[k][t].append(a_mean[k][t])
                acc_offset[k][t].append(a_off[k][t])
                acc_mean_conf[1][t].append(a_conf[1][t])
    res = (_mean_accuracies(acc_mean, args.top_k, args.threshs), _mean_accuracies(acc_offset, args.top_k, args.threshs),
           _mean_accuracies(acc_mean_conf, [1], args.threshs))
    if return_details:
        return res + (dict(matches=np.array(matches), offsets=np.array(offsets), confidences=np.array(confidences),
                           pos_mean=np.array(pos_mean_all), pos_offsets=np.array(pos_off_all)),)
    return res


# ----------------------------------------------------------------------------------------------------------------
# fine stage, cached / device path
# ----------------------------------------------------------------------------------------------------------------
class FineCellCache:
    """Per-cell tables for the fine stage, resident on the device: normalised object encodings ``[n_cells, pad, D]`` of the
    fine model (query independent: ``models/superglue_matcher.py:101-103``), object centres ``[n_cells, pad, 2]`` for the pose
    head, world-frame cell origin / size / scene id for the accuracies."""

    def __init__(self, model, cells: Sequence, args, transform=None, seed: int = 0, cells_per_call: int = 256,
                 padding_factory: Optional[Callable[[], object]] = None):
        self.cell_index = {cell.id: i for i, cell in enumerate(cells)}
        pad = int(args.pad_size)
        dev = model.t2p_device()
        enc, ctr = [], []
        for i0 in range(0, len(cells), cells_per_call):
            chunk = cells[i0:i0 + cells_per_call]
            objs = [padded_objects(c, pad, padding_factory or seeded_padding_factory(seed, c.id)) for c in chunk]
            pts = [_points_of(o, transform, seed, (i0 + j) * pad) for j, o in enumerate(objs)]
            e, c2 = model.encode_cells(objs, pts)
            enc.append(e)
            ctr.append(c2)
        self._finish(torch.cat(enc), torch.cat(ctr), cells, dev)

    @classmethod
    def from_store(cls, model, store, seed: int = 0, cells_per_call: int = 256):
        """The same tables from a PADDED ``CellStore`` resident on the device (``CellStore.from_cells(cells, pad_size,
        padding_factory)``): raw points -> ``t2p_batch_object_points`` -> object encoder, no host loop over objects."""
        self = cls.__new__(cls)
        self.cell_index = {cid: i for i, cid in enumerate(store.cell_ids)}
        co = store.cell_offsets.cpu()
        pad = int(co[1] - co[0])
        if not bool(((co[1:] - co[:-1]) == pad).all()):
            raise ValueError("FineCellCache.from_store needs a store whose cells all hold pad_size objects")
        enc, ctr = [], []
        for c0 in range(0, store.num_cells, cells_per_call):
            cells, ctr64, _ = store.batch_object_points(c0, min(store.num_cells, c0 + cells_per_call), seed=seed, return_extras=True)
            enc.append(model.encode_cells_packed(cells, pad))
            ctr.append(ctr64[:, 0:2].reshape(-1, pad, 2))

        class _C:  # world-frame fields of the store, in the shape _finish reads them
            def __init__(s, i):
                s.id, s.bbox_w, s.cell_size = store.cell_ids[i], store.bbox_w[i], store.cell_size[i]

        self._finish(torch.cat(enc), torch.cat(ctr).contiguous(), [_C(i) for i in range(store.num_cells)], store.device)
        return self

    def _finish(self, obj_enc, centers, cells, dev):
        self.obj_enc = obj_enc.contiguous()     # [n_cells, pad, D] unit rows
        self.centers = centers.contiguous()     # [n_cells, pad, 2] float64
        self.origin = torch.tensor(np.array([c.bbox_w[0:2] for c in cells]), dtype=torch.float64, device=dev)
        self.cell_size = torch.tensor([float(c.cell_size) for c in cells], dtype=torch.float64, device=dev)
        scenes = sorted({c.id.split("_")[0] for c in cells})
        self.scene_index = {s: i for i, s in enumerate(scenes)}
        self.scene = torch.tensor([self.scene_index[c.id.split("_")[0]] for c in cells], dtype=torch.int32, device=dev)
        self.device = dev


@torch.no_grad()
def run_fine_cached(model, retrievals, dataloader, args, cache: Optional[FineCellCache] = None, transform=None, seed: int = 0,
                    queries_per_call: int = 64, query_range: Optional[range] = None, return_details: bool = False):
    """Same results as ``run_fine``; the object encoder runs once per CELL (not once per retrieved (query, cell) pair), the
    matcher runs on batches of ``queries_per_call * max(top_k)`` samples, and in-cell positions + threshold hits are computed
    on the device (``t2p_pose_head`` / ``t2p_pose_accuracy``).  ``query_range``: the slice of queries this replica owns
    (fine stage = replicas only); returns SUMS over that slice as well so that replicas can be combined."""
    from .superglue_matcher import pose_accuracy

    dataset = dataloader.dataset
    if cache is None:
        cache = FineCellCache(model, dataset.all_cells, args, transform, seed)
    K = max(args.top_k)
    rng_q = range(len(retrievals)) if query_range is None else query_range
    top_k, threshs = list(args.top_k), list(args.threshs)
    sums = {name: np.zeros((len(top_k), len(threshs)), dtype=np.int64) for name in ("mean", "offset")}
    sums["mean_conf"] = np.zeros((1, len(threshs)), dtype=np.int64)
    details = dict(matches=[], offsets=[], confidences=[], pos_mean=[], pos_offsets=[])
    qs = list(rng_q)
    for q0 in range(0, len(qs), queries_per_call):
        chunk = qs[q0:q0 + queries_per_call]
        hints = [create_hint_description(dataset.all_poses[q]) for q in chunk]
        cell_idx = torch.tensor([[cache.cell_index[cid] for cid in retrievals[q]] for q in chunk], dtype=torch.int64)
        assert cell_idx.shape[1] == K
        out = model.forward_cached(cache, cell_idx.to(cache.device), hints)
        pose_w = torch.tensor(np.array([dataset.all_poses[q].pose_w[0:2] for q in chunk]), dtype=torch.float64, device=cache.device)
        pose_scene = torch.tensor([cache.scene_index.get(dataset.all_poses[q].cell_id.split("_")[0], -1) for q in chunk],
                                  dtype=torch.int32, device=cache.device)
        hits = pose_accuracy(cache, cell_idx.to(cache.device), out, pose_w, pose_scene, top_k, threshs)  # [3, Q, nk, nt] int32
        h = hits.cpu().numpy()
        sums["mean"] += h[0].sum(0)
        sums["offset"] += h[1].sum(0)
        sums["mean_conf"] += h[2].sum(0)[:1]
        if return_details:
            Q = len(chunk)
            details["matches"].append(out.matches0.reshape(Q, K, -1).cpu().numpy())
            details["offsets"].append(out.offsets.reshape(Q, K, -1, 2).cpu().numpy())
            details["confidences"].append(out.confidence.reshape(Q, K).cpu().numpy())
            details["pos_mean"].append(out.pos_mean.reshape(Q, K, 2).cpu().numpy())
            details["pos_offsets"].append(out.pos_offsets.reshape(Q, K, 2).cpu().numpy())
    n = max(1, len(qs))
    to_dict = lambda a, ks: {k: {t: a[i, j] / n for j, t in enumerate(threshs)} for i, k in enumerate(ks)}
    res = (to_dict(sums["mean"], top_k), to_dict(sums["offset"], top_k), to_dict(sums["mean_conf"], [1]))
    extra = dict(sums=sums, n_queries=len(qs))
    if return_details:
        extra.update({k: np.concatenate(v) if v else np.zeros((0,)) for k, v in details.items()})
    return res + (extra,)


def combine_replica_sums(parts: Sequence[Dict], top_k, threshs):
    """Combine the ``extra`` dicts of ``run_fine_cached`` from several replicas (each over its own ``query_range``)."""
    n = sum(p["n_queries"] for p in parts)
    out = []
    for name, ks in (("mean", list(top_k)), ("offset", list(top_k)), ("mean_conf", [1])):
        s = sum(p["sums"][name] for p in parts)
        out.append({k: {t: s[i, j] / max(1, n) for j, t in enumerate(threshs)} for i, k in enumerate(ks)})
    return tuple(out)


# ----------------------------------------------------------------------------------------------------------------
# the whole pipeline on R GPUs (BASELINE config 5)
# ----------------------------------------------------------------------------------------------------------------
class _DatasetView:
    """``dataloader``-shaped view (``.dataset.all_poses / all_cells``) for the stage functions."""

    def __init__(self, dataset):
        self.dataset = dataset


@torch.no_grad()
def run_pipeline_distributed(coarse_model, fine_model, dataset, args, group=None, seed: int = 0, query_batch: int = 64,
                             return_details: bool = False):
    """Coarse -> fine evaluation of ``dataset`` (``all_cells`` / ``all_poses``) on the ranks of a ``torch.distributed`` group,
    one rank per GPU:

    1. DB build: the RAW cells are sharded like the embeddings (``CellStore.shard``) -- every rank resamples / encodes its own
       block on the device, so the cell embeddings are born in place; no collective.
    2. Coarse retrieval: queries are data-parallel (rank r owns a contiguous block of the poses); per batch one all-gather of
       the query embeddings, the local top-k of all queries against the shard, one all-gather of the per-shard lists and the
       merge (``ShardedCellDatabase.topk_dp``).
    3. Fine stage: replicas only.  The query-independent object encodings of the (padded) cells are computed once, each rank
       its block, and all-gathered (8 KB per cell); every rank then matches its own queries against their retrieved cells
       (``run_fine_cached``: hint LSTM + SuperGlue gather kernel + pose head + accuracies on the device).
    4. The hit counts are summed over the ranks (one all-reduce).

    Returns ``(coarse_acc, acc_mean, acc_offset, acc_mean_conf, info)`` with the reference's accuracy dicts
    (``{k: {thresh: acc}}``); ``info`` holds this rank's retrievals and stage timings (seconds)."""
    import torch.distributed as dist

    from .cell_store import CellStore, build_cell_database
    from .retrieval import ShardedCellDatabase, shard_bounds

    R, r = dist.get_world_size(group), dist.get_rank(group)
    dev = coarse_model.t2p_device()
    cells, poses = dataset.all_cells, dataset.all_poses
    n_cells, n_q = len(cells), len(poses)
    top_k, threshs = list(args.top_k), list(args.threshs)
    K = max(top_k)
    lo, hi = shard_bounds(n_cells, R)[r]
    q_lo, q_hi = shard_bounds(n_q, R)[r]
    per_q = shard_bounds(n_q, R)[0][1]  # queries per rank (the last rank may own fewer: its batches are padded)
    times = {}

    def tick(name, t0):
        torch.cuda.synchronize(dev)
        times[name] = time.perf_counter() - t0

    # 1. sharded DB build ---------------------------------------------------------------------------------------------
    t0 = time.perf_counter()
    obj_before = sum(len(c.objects) for c in cells[:lo])
    shard = CellStore.from_cells(cells[lo:hi]) if hi > lo else None
    if shard is not None:
        shard.obj_id_offset = obj_before
        local_emb = build_cell_database(coarse_model, shard.to(dev), seed=seed)
    else:
        local_emb = torch.empty(0, coarse_model.embed_dim, device=dev)
    tick("db_build_s", t0)

    # 2. coarse retrieval, queries data-parallel ----------------------------------------------------------------------
    t0 = time.perf_counter()
    sdb = ShardedCellDatabase(local_emb, n_cells, group)
    cell_ids = np.array([c.id for c in cells])
    own = list(range(q_lo, q_hi))
    retrievals = {}
    for b0 in range(0, per_q, query_batch):
        idx = [own[min(b0 + i, len(own) - 1)] if own else 0 for i in range(min(query_batch, per_q - b0))]  # padded with repeats
        texts = [" ".join(create_hint_description(poses[q])) for q in idx]
        q_enc = coarse_model.encode_text(texts)
        top, _ = sdb.topk_dp(q_enc, K)
        top = top.cpu().numpy()
        for i, q in enumerate(idx):
            if b0 + i < len(own):
                retrievals[q] = cell_ids[top[i]]
    tick("coarse_s", t0)
    cells_dict = {c.id: c for c in cells}
    coarse_sums = np.zeros((len(top_k), len(threshs)), dtype=np.int64)
    for q in own:
        accs = calc_sample_accuracies(poses[q], [cells_dict[c] for c in retrievals[q]], 0.5 * np.ones((K, 2)), top_k, threshs)
        coarse_sums += np.array([[int(accs[k][t]) for t in threshs] for k in top_k])

    # 3. fine stage: cache built in blocks, all-gathered; replicas over the queries --------------------------------------
    t0 = time.perf_counter()
    pad = int(args.pad_size)
    D = fine_model.embed_dim
    per_c = shard_bounds(n_cells, R)[0][1]
    enc_blk = torch.zeros(per_c, pad, D, dtype=torch.float32, device=dev)
    ctr_blk = torch.zeros(per_c, pad, 2, dtype=torch.float64, device=dev)
    if hi > lo:
        pstore = CellStore.from_cells(cells[lo:hi], pad, lambda cell: seeded_padding_factory(seed, cell.id))
        pstore.obj_id_offset = lo * pad
        part = FineCellCache.from_store(fine_model, pstore.to(dev), seed=seed)
        enc_blk[: hi - lo] = part.obj_enc
        ctr_blk[: hi - lo] = part.centers
    enc_all = torch.empty(R * per_c, pad, D, dtype=torch.float32, device=dev)
    ctr_all = torch.empty(R * per_c, pad, 2, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(enc_all, enc_blk, group=group)
    dist.all_gather_into_tensor(ctr_all, ctr_blk, group=group)
    cache = FineCellCache.__new__(FineCellCache)
    cache.cell_index = {c.id: i for i, c in enumerate(cells)}
    cache._finish(enc_all[:n_cells], ctr_all[:n_cells], cells, dev)
    tick("fine_cache_s", t0)
    t0 = time.perf_counter()
    fine = run_fine_cached(fine_model, retrievals, _DatasetView(dataset), args, cache=cache, queries_per_call=query_batch,
                           query_range=range(q_lo, q_hi), return_details=return_details)
    tick("fine_s", t0)

    # 4. reduce the hit counts ------------------------------------------------------------------------------------------
    sums = fine[3]["sums"]
    flat = np.concatenate([coarse_sums.reshape(-1), sums["mean"].reshape(-1), sums["offset"].reshape(-1), sums["mean_conf"].reshape(-1)])
    t = torch.tensor(flat, dtype=torch.int64, device=dev)
    dist.all_reduce(t, group=group)
    flat = t.cpu().numpy()
    nk, nt = len(top_k), len(threshs)
    parts, o = [], 0
    for ks in (top_k, top_k, top_k, [1]):
        n = len(ks) * nt
        a = flat[o: o + n].reshape(len(ks), nt)
        parts.append({k: {th: a[i, j] / max(1, n_q) for j, th in enumerate(threshs)} for i, k in enumerate(ks)})
        o += n
    info = dict(retrievals=retrievals, times=times, query_range=(q_lo, q_hi), cell_range=(lo, hi), local_emb=local_emb)
    if return_details:
        info["details"] = fine[3]
    return parts[0], parts[1], parts[2], parts[3], info
