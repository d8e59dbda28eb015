"""Packed raw cell store + device-side ``batch_object_points``.

The reference keeps cells as pickled lists of ``Object3d`` and turns them into network inputs on the CPU, object by object
(``dataloading/kitti360pose/utils.py:89-110``: ``torch.tensor`` copies, ``FixedPoints``, ``NormalizeScale``,
``Batch.from_data_list``; plus the ``np.mean`` calls of ``models/object_encoder.py:121-131``).  ``CellStore`` is the packed
replacement: all raw points of all objects of all cells in two ragged float32 arrays plus offsets, uploaded once;
``batch_object_points`` is then ONE kernel launch per batch of cells (``t2p_batch_object_points``) that produces the
``PackedCells`` the encoders consume.  ``save`` / ``load`` give the store an on-disk format (one ``.npz``) that replaces the
pickles for evaluation.

Sampling indices come from a counter-based generator (``t2p_fixed_points_index``: splitmix64 of (seed, object, i)), so a cell
resamples identically wherever and whenever it is encoded; ``fixed_points_indices`` is the same function on the host.
"""
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .synthetic import NUM_POINTS, PackedCells

_M64 = (1 << 64) - 1


def fixed_points_indices(seed: int, obj_ids: np.ndarray, n_points: np.ndarray, P: int = NUM_POINTS) -> np.ndarray:
    """Host mirror of ``t2p_fixed_points_index``: [n_obj, P] int32 sampling indices (object ``o`` has ``n_points[o]`` raw points)."""
    obj = np.asarray(obj_ids, dtype=np.uint64)[:, None]
    i = np.arange(P, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M64) + np.uint64(0x9E3779B97F4A7C15) * (obj * np.uint64(P) + i + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        idx = ((z >> np.uint64(32)) * np.asarray(n_points, dtype=np.uint64)[:, None]) >> np.uint64(32)
    return idx.astype(np.int32)


class CellStore:
    """Ragged raw points of a list of cells.

    ``raw_xyz`` / ``raw_rgb`` [total_points, 3] float32, ``obj_offsets`` [n_obj+1] int64 (points of an object),
    ``cell_offsets`` [n_cells+1] int32 (objects of a cell), ``cell_ids`` (strings), ``bbox_w`` [n_cells, 6] / ``cell_size``
    [n_cells] float64 (world frame, for the pose accuracies)."""

    def __init__(self, raw_xyz, raw_rgb, obj_offsets, cell_offsets, cell_ids, bbox_w=None, cell_size=None, obj_id_offset: int = 0):
        self.obj_id_offset = int(obj_id_offset)  # global id of object 0 (a shard resamples its objects like the whole store does)
        self.raw_xyz = torch.as_tensor(raw_xyz, dtype=torch.float32).contiguous()
        self.raw_rgb = torch.as_tensor(raw_rgb, dtype=torch.float32).contiguous()
        self.obj_offsets = torch.as_tensor(obj_offsets, dtype=torch.int64).contiguous()
        self.cell_offsets = torch.as_tensor(cell_offsets, dtype=torch.int32).contiguous()
        self.cell_offsets_host = self.cell_offsets.cpu()  # kept on the host: reading offsets never synchronises the device
        self.cell_ids = [str(c) for c in cell_ids]
        self.bbox_w = None if bbox_w is None else np.asarray(bbox_w, dtype=np.float64)
        self.cell_size = None if cell_size is None else np.asarray(cell_size, dtype=np.float64)
        assert self.cell_offsets.numel() == len(self.cell_ids) + 1
        assert int(self.cell_offsets[-1]) == self.obj_offsets.numel() - 1
        assert int(self.obj_offsets[-1]) == self.raw_xyz.shape[0] == self.raw_rgb.shape[0]

    # ---- construction / persistence ---------------------------------------------------------------------------------
    @classmethod
    def from_cells(cls, cells: Sequence, pad_size: Optional[int] = None,
                   padding_factory: Optional[Callable[[object], Callable[[], object]]] = None) -> "CellStore":
        """``cells``: ``Cell`` duck types (``id``, ``objects`` with ``xyz`` / ``rgb``, ``bbox_w``, ``cell_size``).  With
        ``pad_size`` the object list of every cell is cut / padded like ``Kitti360TopKDataset`` does
        (dataloading/kitti360pose/eval.py:146-157); ``padding_factory(cell)`` returns the per-cell padding-object factory."""
        xyz, rgb, counts, cell_counts = [], [], [], []
        for cell in cells:
            objects = list(cell.objects)
            if pad_size is not None:
                objects = objects[:pad_size]
                make = padding_factory(cell) if padding_factory is not None else None
                while len(objects) < pad_size:
                    if make is None:
                        raise ValueError("pad_size needs a padding_factory")
                    objects.append(make())
            if not objects:
                raise ValueError(f"cell {cell.id} has no objects")
            for obj in objects:
                p = np.asarray(obj.xyz, dtype=np.float32).reshape(-1, 3)
                if p.shape[0] < 1:
                    raise ValueError(f"cell {cell.id}: object without points")
                xyz.append(p)
                rgb.append(np.asarray(obj.rgb, dtype=np.float32).reshape(-1, 3))
                counts.append(p.shape[0])
            cell_counts.append(len(objects))
        return cls(np.concatenate(xyz), np.concatenate(rgb), np.concatenate([[0], np.cumsum(counts)]),
                   np.concatenate([[0], np.cumsum(cell_counts)]), [c.id for c in cells],
                   np.array([c.bbox_w for c in cells]) if hasattr(cells[0], "bbox_w") else None,
                   np.array([c.cell_size for c in cells]) if hasattr(cells[0], "cell_size") else None)

    def save(self, path: str) -> None:
        np.savez(path, raw_xyz=self.raw_xyz.cpu().numpy(), raw_rgb=self.raw_rgb.cpu().numpy(),
                 obj_offsets=self.obj_offsets.cpu().numpy(), cell_offsets=self.cell_offsets.cpu().numpy(),
                 cell_ids=np.array(self.cell_ids), bbox_w=np.zeros((0, 6)) if self.bbox_w is None else self.bbox_w,
                 cell_size=np.zeros(0) if self.cell_size is None else self.cell_size)

    @classmethod
    def load(cls, path: str) -> "CellStore":
        z = np.load(path, allow_pickle=False)
        return cls(z["raw_xyz"], z["raw_rgb"], z["obj_offsets"], z["cell_offsets"], [str(s) for s in z["cell_ids"]],
                   z["bbox_w"] if z["bbox_w"].shape[0] else None, z["cell_size"] if z["cell_size"].shape[0] else None)

    # ---- views -------------------------------------------------------------------------------------------------------
    @property
    def num_cells(self) -> int:
        return len(self.cell_ids)

    @property
    def num_objects(self) -> int:
        return self.obj_offsets.numel() - 1

    @property
    def device(self):
        return self.raw_xyz.device

    def to(self, device) -> "CellStore":
        s = CellStore.__new__(CellStore)
        s.__dict__.update(self.__dict__)
        s.raw_xyz, s.raw_rgb = self.raw_xyz.to(device), self.raw_rgb.to(device)
        s.obj_offsets, s.cell_offsets = self.obj_offsets.to(device), self.cell_offsets.to(device)
        return s

    def shard(self, rank: int, world: int) -> "CellStore":
        """The contiguous block of cells rank ``rank`` of ``world`` owns (same blocks as ``retrieval.shard_bounds``): the raw
        cells are sharded like the embeddings, so a sharded DB build produces every embedding in place."""
        from .retrieval import shard_bounds

        lo, hi = shard_bounds(self.num_cells, world)[rank]
        return self.slice_cells(lo, hi)

    def slice_cells(self, lo: int, hi: int) -> "CellStore":
        co = self.cell_offsets_host
        o0, o1 = int(co[lo]), int(co[hi])
        oo = self.obj_offsets.cpu()
        p0, p1 = int(oo[o0]), int(oo[o1])
        return CellStore(self.raw_xyz[p0:p1], self.raw_rgb[p0:p1], (self.obj_offsets[o0:o1 + 1] - p0),
                         (self.cell_offsets[lo:hi + 1] - o0), self.cell_ids[lo:hi],
                         None if self.bbox_w is None else self.bbox_w[lo:hi], None if self.cell_size is None else self.cell_size[lo:hi],
                         obj_id_offset=self.obj_id_offset + o0)

    # ---- the device data path ---------------------------------------------------------------------------------------
    def batch_object_points(self, cell_lo: int = 0, cell_hi: Optional[int] = None, seed: int = 0, P: int = NUM_POINTS,
                            choice: Optional[torch.Tensor] = None, obj_id_base: Optional[int] = None,
                            return_extras: bool = False):
        """Cells ``[cell_lo, cell_hi)`` -> ``PackedCells`` on the store's device (one kernel).  ``choice`` [n_obj, P] int32
        overrides the counter-based sampling; ``obj_id_base``: global id of the first object (default: its index in the
        store + the store's ``obj_id_offset``, so a shard samples exactly like the whole store).  ``return_extras``: also (centers64 [n_obj,3] float64, choice [n_obj,P] int32)."""
        lib = _lib.load()
        _lib.require_cuda(self.raw_xyz, "cell store")
        cell_hi = self.num_cells if cell_hi is None else cell_hi
        dev = self.device
        co = self.cell_offsets[cell_lo:cell_hi + 1]
        co_host = self.cell_offsets_host[cell_lo:cell_hi + 1]
        o0, o1 = int(co_host[0]), int(co_host[-1])
        n_obj = o1 - o0
        pos = torch.empty(n_obj, P, 3, dtype=torch.float32, device=dev)
        rgb = torch.empty(n_obj, P, 3, dtype=torch.float32, device=dev)
        ctr = torch.empty(n_obj, 3, dtype=torch.float32, device=dev)
        col = torch.empty(n_obj, 3, dtype=torch.float32, device=dev)
        ctr64 = torch.empty(n_obj, 3, dtype=torch.float64, device=dev) if return_extras else None
        ch_out = torch.empty(n_obj, P, dtype=torch.int32, device=dev) if return_extras else None
        if choice is not None:
            choice = choice.to(dev, torch.int32).contiguous()
            assert tuple(choice.shape) == (n_obj, P)
        offs = self.obj_offsets[o0:o1 + 1]
        with torch.cuda.device(dev):
            _lib.check(
                lib.t2p_batch_object_points(_lib.ptr(self.raw_xyz), _lib.ptr(self.raw_rgb), offs.data_ptr(), n_obj, P,
                                            _lib.ptr(choice), int(seed) & _M64, int(self.obj_id_offset + o0 if obj_id_base is None else obj_id_base),
                                            _lib.ptr(pos), _lib.ptr(rgb), _lib.ptr(ctr), _lib.ptr(col), _lib.ptr(ctr64),
                                            _lib.ptr(ch_out), _lib.stream_ptr(dev)),
                "batch_object_points",
            )
        cells = PackedCells(pos, rgb, ctr, col, (co - o0).to(torch.int32), [int(x) - o0 for x in co_host.tolist()])
        return (cells, ctr64, ch_out) if return_extras else cells


def build_cell_database(model, store: CellStore, seed: int = 0, cells_per_call: int = 512) -> torch.Tensor:
    """DB build over a (shard of a) cell store, everything on the device: raw points -> ``batch_object_points`` kernel ->
    PointNet++ / object encoder / cell aggregation -> ``[n_cells, D]`` unit-norm embeddings (models/cell_retrieval.py:77-107)."""
    out = []
    for c0 in range(0, store.num_cells, cells_per_call):
        cells = store.batch_object_points(c0, min(store.num_cells, c0 + cells_per_call), seed=seed)
        out.append(model.encode_cells_packed(cells))
    D = model.embed_dim
    return torch.cat(out) if out else torch.empty(0, D, device=store.device)
