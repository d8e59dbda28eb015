"""Drop-in for ``training.coarse.eval_epoch`` (reference ``training/coarse.py:68-167``): top-k retrieval of every pose of a
dataset against all of its cells, with the reference's return values -- ``(accuracies, accuracies_close, top_retrievals
[, cell_encodings, text_encodings])``.

What changes underneath: the encodings stay on the GPU, and the reference's per-query float64 numpy mat-vec + full
``argsort`` (``:134-140``) is one ``t2p_retrieve_topk`` call (same float64 ranking, certified); the hit / close-by accuracy
bookkeeping (``:142-160``) is the reference's, on the host.

The data side is duck-typed exactly as the reference uses it: ``dataloader`` yields dict-of-lists batches with ``"texts"``
and ``"cell_ids"`` and exposes ``.dataset`` with ``all_poses`` (``pose.pose_w``), ``all_cells`` and ``get_cell_dataset()``;
the cell dataset has ``.cells`` (``cell.id``, ``cell.cell_size``, ``cell.get_center()``) and items with ``"objects"``,
``"object_points"``, ``"cell_ids"``.
"""
from typing import Dict, List

import numpy as np
import torch

from .retrieval import CellDatabase


def _collate(items: List[dict]) -> Dict[str, list]:
    """``Kitti360CoarseDataset.collate_fn`` (dataloading/kitti360pose/base.py:81-85): dict of lists."""
    return {k: [it[k] for it in items] for k in items[0]}


@torch.no_grad()
def eval_epoch_store(model, dataloader, args, return_encodings: bool = False, seed: int = 0):
    """``eval_epoch`` with the database side on the device data path: the raw cells go into a ``CellStore`` once and
    ``t2p_batch_object_points`` resamples / normalises them on the GPU (counter-based indices, ``seed``) instead of the
    per-object host transforms of the cell dataset.  Same return values."""
    from .cell_store import CellStore, build_cell_database

    dataset = dataloader.dataset
    store = CellStore.from_cells(dataset.all_cells).to(model.t2p_device())
    cell_enc = build_cell_database(model, store, seed=seed)
    return eval_epoch(model, dataloader, args, return_encodings, _cell_enc=(cell_enc, [c.id for c in dataset.all_cells]))


@torch.no_grad()
def eval_epoch(model, dataloader, args, return_encodings: bool = False, _cell_enc=None):
    assert getattr(args, "ranking_loss", "pairwise") != "triplet"  # training/coarse.py:81
    model.eval()
    top_k = list(args.top_k)
    accuracies = {k: [] for k in top_k}
    accuracies_close = {k: [] for k in top_k}

    dataset = dataloader.dataset
    cells_dataset = dataset.get_cell_dataset()
    cells_dict = {cell.id: cell for cell in cells_dataset.cells}
    cell_size = cells_dataset.cells[0].cell_size
    query_poses_w = np.array([pose.pose_w[0:2] for pose in dataset.all_poses])

    # query side (training/coarse.py:108-118), kept on the device
    text_enc, query_cell_ids = [], []
    for batch in dataloader:
        text_enc.append(model.encode_text(batch["texts"]))
        query_cell_ids.extend(batch["cell_ids"])
    text_enc = torch.cat(text_enc)
    query_cell_ids = np.array(query_cell_ids, dtype="<U32")

    # database side (:121-131)
    if _cell_enc is not None:  # encoded by the caller (eval_epoch_store)
        cell_enc, db_cell_ids = _cell_enc
    else:
        cell_enc, db_cell_ids = [], []
        bs = int(args.batch_size)
        for i0 in range(0, len(cells_dataset), bs):
            batch = _collate([cells_dataset[i] for i in range(i0, min(i0 + bs, len(cells_dataset)))])
            cell_enc.append(model.encode_objects(batch["objects"], batch["object_points"]))
            db_cell_ids.extend(batch["cell_ids"])
        cell_enc = torch.cat(cell_enc)
    db_cell_ids = np.array(db_cell_ids, dtype="<U32")
    assert len(db_cell_ids) == len(dataset.all_cells)  # :137

    # all-pairs scores + top-k (:134-140) on the GPU; float64 ranking, (score desc, index asc)
    kmax = min(int(np.max(top_k)), len(db_cell_ids))  # a DB smaller than max(top_k) returns all of its cells, like argsort[0:k]
    db = CellDatabase(cell_enc, db_cell_ids)
    sorted_indices, _ = db.topk(text_enc, kmax)
    sorted_indices = sorted_indices.cpu().numpy()

    top_retrievals = {}
    for query_idx in range(len(text_enc)):
        retrieved_cell_ids = db_cell_ids[sorted_indices[query_idx]]
        target_cell_id = query_cell_ids[query_idx]
        for k in top_k:
            accuracies[k].append(target_cell_id in retrieved_cell_ids[0:k])
        top_retrievals[query_idx] = retrieved_cell_ids
        # close-by accuracy (:150-160)
        target_pose_w = query_poses_w[query_idx]
        retrieved_cell_poses = [cells_dict[cell_id].get_center()[0:2] for cell_id in retrieved_cell_ids]
        dists = np.linalg.norm(target_pose_w - retrieved_cell_poses, axis=1)
        for k in top_k:
            accuracies_close[k].append(np.any(dists[0:k] <= cell_size / 2))

    for k in top_k:
        accuracies[k] = np.mean(accuracies[k])
        accuracies_close[k] = np.mean(accuracies_close[k])

    if return_encodings:
        return accuracies, accuracies_close, top_retrievals, cell_enc.cpu().numpy().astype(np.float64), text_enc.cpu().numpy().astype(np.float64)
    return accuracies, accuracies_close, top_retrievals
