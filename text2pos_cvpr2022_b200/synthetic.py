"""Synthetic KITTI360Pose-shaped inputs and weights (there is no dataset / checkpoint offline).

Everything here is deterministic given a seed (numpy ``default_rng`` = PCG64, stable
across numpy versions) so that tests, the golden-vector generator and ``bench.py``
can rebuild identical inputs without shipping them.

Shapes follow the reference's data layer:
* objects: ``Object3d`` duck type (``datapreparation/kitti360pose/imports.py:8-83``) --
  ``.xyz .rgb .label get_center() get_color_rgb() get_color_text()``;
* object points: ``FixedPoints(256)`` (sampling WITH replacement) + ``NormalizeScale``
  (``evaluation/pipeline.py:290-293``, ``dataloading/kitti360pose/utils.py:89-110``),
  batched per cell into a PyG-``Batch``-like ``.x .pos .batch`` container;
* hints: ``"The pose is {direction} of a {color} {class}."``
  (``dataloading/kitti360pose/base.py:63-65``).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# class / colour / direction vocabularies of KITTI360Pose
# (datapreparation/kitti360pose/utils.py:48-71,208; datapreparation/kitti360pose/select.py:16-27)
KNOWN_CLASSES = (
    "building,pole,traffic light,traffic sign,garage,stop,smallpole,lamp,trash bin,vending machine,box,"
    "road,sidewalk,parking,wall,fence,guard rail,bridge,tunnel,vegetation,terrain,pad"
).split(",")
COLOR_NAMES = ["dark-green", "gray", "gray-green", "bright-gray", "gray", "black", "green", "beige"]
COLOR_CENTERS = (
    np.array(
        [
            [47.2579917, 49.75368454, 42.4153065],
            [136.32696657, 136.95241796, 126.02741229],
            [87.49822126, 91.69058836, 80.14558512],
            [213.91030679, 216.25033052, 207.24611073],
            [110.39218852, 112.91977458, 103.68638249],
            [27.47505158, 28.43996795, 25.16840296],
            [66.65951839, 70.22342483, 60.20395996],
            [171.00852191, 170.05737735, 155.00130334],
        ]
    )
    / 255.0
)
DIRECTIONS = ["east", "west", "north", "south", "on-top"]
NUM_POINTS = 256  # training/args.py:53
PAD_SIZE = 16  # training/args.py:42
NUM_HINTS = 6  # training/args.py:41


def known_words() -> List[str]:
    """Vocabulary as ``Kitti360BaseDataset.get_known_words`` would collect it from all template hints."""
    words = set("the pose is of a".split())
    words.update(DIRECTIONS)
    words.update(COLOR_NAMES)
    for c in KNOWN_CLASSES:
        words.update(c.split())
    return sorted(words)


# ----------------------------------------------------------------------------------------------
# duck-typed value types
# ----------------------------------------------------------------------------------------------
class SynthObject3d:
    """Duck type of the reference ``Object3d`` (the fields/methods the hot path reads)."""

    def __init__(self, id: int, xyz: np.ndarray, rgb: np.ndarray, label: str):
        self.id = id
        self.instance_id = id
        self.xyz = xyz
        self.rgb = rgb
        self.label = label

    def get_color_rgb(self):
        return np.mean(self.rgb, axis=0)

    def get_center(self):
        return np.mean(self.xyz, axis=0)

    def get_color_text(self):
        d = np.linalg.norm(np.mean(self.rgb, axis=0) - COLOR_CENTERS, axis=1)
        return COLOR_NAMES[int(np.argmin(d))]

    @classmethod
    def create_padding(cls, rng: np.random.Generator):
        """Seeded version of ``Object3d.create_padding`` (imports.py:74-83): 8 points x 0.001 scale, rgb 0."""
        return cls(-1, rng.random((8, 3)) * 0.001, np.zeros((8, 3)), "pad")


class SynthCell:
    """Duck type of the reference ``Cell`` (imports.py:221-247)."""

    def __init__(self, idx: int, scene_name: str, objects: List[SynthObject3d], cell_size: float, bbox_w: np.ndarray):
        self.scene_name = scene_name
        self.id = f"{scene_name}_{idx:05.0f}"
        self.objects = objects
        self.cell_size = cell_size
        self.bbox_w = bbox_w

    def get_center(self):
        return 0.5 * (self.bbox_w[0:3] + self.bbox_w[3:6])


class PointBatch:
    """Minimal stand-in for ``torch_geometric.data.Batch``: ``.x`` (rgb), ``.pos`` (xyz), ``.batch``."""

    def __init__(self, x: torch.Tensor, pos: torch.Tensor, batch: torch.Tensor):
        self.x, self.pos, self.batch = x, pos, batch

    def to(self, device):
        self.x, self.pos, self.batch = self.x.to(device), self.pos.to(device), self.batch.to(device)
        return self

    @property
    def num_graphs(self) -> int:
        return int(self.batch.max().item()) + 1 if self.batch.numel() else 0


# ----------------------------------------------------------------------------------------------
# point clouds
# ----------------------------------------------------------------------------------------------
def fixed_points_normalize(xyz: np.ndarray, rgb: np.ndarray, rng: np.random.Generator, num: int = NUM_POINTS):
    """``Compose([FixedPoints(num), NormalizeScale()])``: resample WITH replacement, centre, scale to (-1,1)."""
    return fixed_points_normalize_idx(xyz, rgb, rng.integers(0, xyz.shape[0], size=num))


def fixed_points_normalize_idx(xyz: np.ndarray, rgb: np.ndarray, choice: np.ndarray):
    """The same transform for given sampling indices (what ``t2p_batch_object_points`` computes on the device: float32 mean
    accumulated row by row, scale = float32((1 / max|pos|) * 0.999999))."""
    pos = xyz[choice].astype(np.float32)
    col = rgb[choice].astype(np.float32)
    pos = pos - pos.mean(axis=0, keepdims=True)
    scale = np.float32((1.0 / max(float(np.abs(pos).max()), 1e-12)) * 0.999999)
    return pos * scale, col


def batch_object_points_idx(objects: Sequence["SynthObject3d"], seed: int, obj_id0: int, num: int = NUM_POINTS) -> "PointBatch":
    """``batch_object_points`` with the counter-based sampling of the device data path (``cell_store.fixed_points_indices``):
    object ``i`` of the list is global object ``obj_id0 + i``."""
    from .cell_store import fixed_points_indices

    n_pts = np.array([np.asarray(o.xyz).reshape(-1, 3).shape[0] for o in objects])
    choice = fixed_points_indices(seed, obj_id0 + np.arange(len(objects)), n_pts, num)
    xs, ps = [], []
    for o, ch in zip(objects, choice):
        pos, col = fixed_points_normalize_idx(np.asarray(o.xyz, dtype=np.float32), np.asarray(o.rgb, dtype=np.float32), ch)
        ps.append(pos)
        xs.append(col)
    batch = np.repeat(np.arange(len(objects)), num)
    return PointBatch(torch.from_numpy(np.concatenate(xs)), torch.from_numpy(np.concatenate(ps)), torch.from_numpy(batch))


def synth_object(rng: np.random.Generator, kind: Optional[int] = None, n_src: Optional[int] = None, obj_id: int = 0):
    """A raw object: ``kind`` 0=planar patch, 1=thin pole, 2=blob; ``n_src`` distinct points."""
    kind = int(rng.integers(0, 3)) if kind is None else kind
    n_src = int(rng.choice([8, 40, 400, 2000])) if n_src is None else n_src
    if kind == 0:
        p = rng.random((n_src, 3)) * np.array([0.3, 0.25, 0.004])
    elif kind == 1:
        p = rng.random((n_src, 3)) * np.array([0.01, 0.01, 0.35])
    else:
        p = rng.normal(size=(n_src, 3)) * 0.05
    centre = rng.random(3)
    xyz = p - p.mean(axis=0) + centre
    base = rng.random(3)
    rgb = np.clip(base + rng.normal(size=(n_src, 3)) * 0.08, 0.0, 1.0)
    label = KNOWN_CLASSES[int(rng.integers(0, len(KNOWN_CLASSES) - 1))]
    return SynthObject3d(obj_id, xyz, rgb, label)


def synth_cell(rng: np.random.Generator, idx: int, n_obj: Optional[int] = None, scene: str = "0010", cell_size=30.0,
               origin: Optional[np.ndarray] = None):
    n_obj = int(rng.integers(6, 17)) if n_obj is None else n_obj
    objects = [synth_object(rng, obj_id=i) for i in range(n_obj)]
    if origin is None:
        origin = np.array([rng.random() * 1000.0, rng.random() * 1000.0, 0.0])
    bbox = np.concatenate([origin, origin + cell_size])
    return SynthCell(idx, scene, objects, cell_size, bbox)


def batch_object_points(objects: Sequence[SynthObject3d], rng: np.random.Generator, num: int = NUM_POINTS) -> PointBatch:
    """Seeded restatement of ``batch_object_points`` (dataloading/kitti360pose/utils.py:89-110) for one cell."""
    xs, ps = [], []
    for obj in objects:
        pos, col = fixed_points_normalize(obj.xyz, obj.rgb, rng, num)
        ps.append(pos)
        xs.append(col)
    batch = np.repeat(np.arange(len(objects)), num)
    return PointBatch(
        torch.from_numpy(np.concatenate(xs)), torch.from_numpy(np.concatenate(ps)), torch.from_numpy(batch)
    )


@dataclass
class PackedCells:
    """The packed (tensor) form of a list of cells that the fast path consumes.

    pos/rgb [n_obj_total, P, 3] float32 (post-transform), centers/mean_rgb [n_obj_total, 3] float32
    (from the RAW points, models/object_encoder.py:121-131), cell_offsets [n_cells+1] int32.
    """

    pos: torch.Tensor
    rgb: torch.Tensor
    centers: torch.Tensor
    mean_rgb: torch.Tensor
    cell_offsets: torch.Tensor
    offsets_host: Optional[List[int]] = None  # the same offsets on the host (kept across .to(): no device sync to read them)

    @property
    def num_cells(self) -> int:
        return self.cell_offsets.numel() - 1

    def host_offsets(self) -> List[int]:
        if self.offsets_host is None:
            self.offsets_host = [int(x) for x in self.cell_offsets.tolist()]
        return self.offsets_host

    def to(self, device):
        host = self.host_offsets() if not self.cell_offsets.is_cuda else self.offsets_host
        return PackedCells(*(t.to(device) for t in (self.pos, self.rgb, self.centers, self.mean_rgb, self.cell_offsets)), host)

    def cell_slices(self):
        off = self.host_offsets()
        return [(off[i], off[i + 1]) for i in range(len(off) - 1)]


def pack_cells(objects: Sequence[Sequence], object_points: Sequence) -> PackedCells:
    """(List[List[Object3d]], List[Batch]) -> PackedCells.  Requires every object to hold the same point count."""
    counts = [len(o) for o in objects]
    pos = torch.cat([b.pos for b in object_points]).float()
    rgb = torch.cat([b.x for b in object_points]).float()
    n_obj = sum(counts)
    if n_obj == 0 or pos.shape[0] % n_obj != 0:
        raise ValueError("object_points must hold the same number of points for every object (FixedPoints)")
    P = pos.shape[0] // n_obj
    for b, c in zip(object_points, counts):
        if b.pos.shape[0] != c * P:
            raise ValueError("object_points must hold the same number of points for every object (FixedPoints)")
    centers = np.array([obj.get_center() for cell in objects for obj in cell], dtype=np.float64)
    colors = np.array([obj.get_color_rgb() for cell in objects for obj in cell], dtype=np.float64)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return PackedCells(
        pos.reshape(n_obj, P, 3).contiguous(),
        rgb.reshape(n_obj, P, 3).contiguous(),
        torch.tensor(centers, dtype=torch.float),
        torch.tensor(colors, dtype=torch.float),
        torch.from_numpy(off),
    )


def synth_cells(seed: int, n_cells: int, n_obj: Optional[int] = None):
    """-> (cells, objects List[List], object_points List[PointBatch]) like a cell-dataset batch."""
    rng = np.random.default_rng(seed)
    cells = [synth_cell(rng, i, n_obj) for i in range(n_cells)]
    objects = [c.objects for c in cells]
    points = [batch_object_points(c.objects, rng) for c in cells]
    return cells, objects, points


def synth_packed_cells(seed: int, n_cells: int, n_obj: Optional[int] = None) -> PackedCells:
    _, objects, points = synth_cells(seed, n_cells, n_obj)
    return pack_cells(objects, points)


# ----------------------------------------------------------------------------------------------
# text
# ----------------------------------------------------------------------------------------------
class SynthDescription:
    """Duck type of ``DescriptionBestCell`` (imports.py:125-175): the three fields the hint template reads."""

    def __init__(self, direction: str, object_color_text: str, object_label: str):
        self.direction, self.object_color_text, self.object_label = direction, object_color_text, object_label

    def hint(self) -> str:
        """``Kitti360BaseDataset.create_hint_description`` template (dataloading/kitti360pose/base.py:63-65)."""
        return f"The pose is {self.direction} of a {self.object_color_text} {self.object_label}."


def synth_description(rng: np.random.Generator) -> SynthDescription:
    d = DIRECTIONS[int(rng.integers(0, len(DIRECTIONS)))]
    c = COLOR_NAMES[int(rng.integers(0, len(COLOR_NAMES)))]
    k = KNOWN_CLASSES[int(rng.integers(0, len(KNOWN_CLASSES) - 1))]
    return SynthDescription(d, c, k)


def synth_hint(rng: np.random.Generator) -> str:
    return synth_description(rng).hint()


def synth_hints(seed: int, n_queries: int, n_hints: int = NUM_HINTS) -> List[List[str]]:
    rng = np.random.default_rng(seed)
    return [[synth_hint(rng) for _ in range(n_hints)] for _ in range(n_queries)]


def synth_queries(seed: int, n_queries: int, n_hints: int = NUM_HINTS) -> List[str]:
    """Coarse query texts: the hints joined by " " (dataloading/kitti360pose/cells.py:82)."""
    return [" ".join(h) for h in synth_hints(seed, n_queries, n_hints)]


# ----------------------------------------------------------------------------------------------
# a KITTI360Pose-shaped coarse dataset (what training.coarse.eval_epoch / evaluation.pipeline.run_coarse consume)
# ----------------------------------------------------------------------------------------------
class SynthPose:
    """Duck type of the reference ``Pose`` (imports.py:178-218): world position, its cell, six hint descriptions."""

    def __init__(self, pose_w: np.ndarray, cell_id: str, descriptions: List[SynthDescription]):
        self.pose_w = pose_w
        self.cell_id = cell_id
        self.descriptions = descriptions

    @property
    def hints(self) -> List[str]:
        return [d.hint() for d in self.descriptions]


class SynthCellDataset:
    """``Kitti360CoarseCellOnlyDataset`` stand-in (dataloading/kitti360pose/cells.py:163-213)."""

    def __init__(self, cells: List[SynthCell], seed: int):
        self.cells = cells
        self._points = {}
        self._seed = seed

    def __len__(self):
        return len(self.cells)

    def __getitem__(self, idx: int):
        cell = self.cells[idx]
        if idx not in self._points:  # FixedPoints resampling is random in the reference; seeded per cell here
            self._points[idx] = batch_object_points(cell.objects, np.random.default_rng(self._seed + 7919 * idx))
        return {"cells": cell, "cell_ids": cell.id, "objects": cell.objects, "object_points": self._points[idx]}


class SynthCoarseDataset:
    """``Kitti360CoarseDatasetMulti`` stand-in: ``n_poses`` poses spread over ``n_cells`` cells of one scene; the text of a
    pose is its six template hints joined by a space (dataloading/kitti360pose/cells.py:82)."""

    def __init__(self, seed: int, n_cells: int, n_poses: int, scenes: Sequence[str] = ("0010",), grid_stride: Optional[float] = None,
                 max_objects: Optional[int] = None):
        """``scenes``: cells are dealt round-robin to these scene names (ids ``"<scene>_<idx>"``); ``grid_stride``: lay the
        cells of a scene out on a square grid with this stride in metres (KITTI360Pose cells overlap: stride 10, size 30)
        instead of scattering them; ``max_objects``: upper bound of the objects per cell (default 16; more exercises the
        cut-off of the top-k dataset)."""
        rng = np.random.default_rng(seed)
        self.all_cells = []
        side = int(np.ceil(np.sqrt(max(1, n_cells / max(1, len(scenes))))))
        for i in range(n_cells):
            scene, j = scenes[i % len(scenes)], i // len(scenes)
            origin = None if grid_stride is None else np.array([(j % side) * grid_stride, (j // side) * grid_stride, 0.0])
            n_obj = None if max_objects is None else int(rng.integers(6, max_objects + 1))
            self.all_cells.append(synth_cell(rng, i, n_obj, scene=scene, origin=origin))
        self.all_poses = []
        for _ in range(n_poses):
            c = self.all_cells[int(rng.integers(0, n_cells))]
            pose_w = c.bbox_w[0:3] + rng.random(3) * c.cell_size
            self.all_poses.append(SynthPose(pose_w, c.id, [synth_description(rng) for _ in range(NUM_HINTS)]))
        self._cell_dataset = SynthCellDataset(self.all_cells, seed)

    def __len__(self):
        return len(self.all_poses)

    def __getitem__(self, idx: int):
        pose = self.all_poses[idx]
        return {"poses": pose, "texts": " ".join(pose.hints), "cell_ids": pose.cell_id}

    def get_cell_dataset(self):
        return self._cell_dataset


class SynthLoader:
    """``DataLoader(dataset, batch_size, collate_fn=dict-of-lists, shuffle=False)`` stand-in with the ``.dataset`` attribute
    the reference's evaluation code reads."""

    def __init__(self, dataset, batch_size: int):
        self.dataset, self.batch_size = dataset, int(batch_size)

    def __iter__(self):
        n = len(self.dataset)
        for i0 in range(0, n, self.batch_size):
            items = [self.dataset[i] for i in range(i0, min(i0 + self.batch_size, n))]
            yield {k: [it[k] for it in items] for k in items[0]}

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size


# ----------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------
def synth_state_dict(spec: Sequence[Tuple[str, Sequence[int]]], seed: int, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random-but-reproducible weights for a list of (state_dict key, shape).

    BatchNorm statistics are randomised (mean~N(0,.1), var~U(.5,1.5), gamma~U(.5,1.5), beta~N(0,.1)) so
    that BN folding is exercised; matrices ~ N(0, gain^2/fan_in); embedding row 0 (padding_idx) is zero.
    """
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in spec:
        shape = tuple(int(s) for s in shape)
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros(shape, dtype=torch.long)
            continue
        if name.endswith("running_var"):
            a = rng.uniform(0.5, 1.5, size=shape)
        elif name.endswith("running_mean"):
            a = rng.normal(size=shape) * 0.1
        elif name.endswith("embedding.weight"):
            a = rng.normal(size=shape) * 0.5
            a[0] = 0.0
        elif len(shape) == 0:
            a = np.asarray(1.0 + 0.1 * rng.normal())
        elif len(shape) == 1 and name.endswith("weight"):
            a = rng.uniform(0.5, 1.5, size=shape)
        elif len(shape) == 1:
            a = rng.normal(size=shape) * 0.1
        else:
            fan_in = int(np.prod(shape[1:]))
            a = rng.normal(size=shape) * (gain / np.sqrt(fan_in))
        sd[name] = torch.tensor(np.asarray(a), dtype=torch.float32)
    return sd


def randomize_module_(module: torch.nn.Module, seed: int, gain: float = 1.0) -> torch.nn.Module:
    """Overwrite every parameter/buffer of ``module`` with ``synth_state_dict`` values (in place)."""
    cur = module.state_dict()
    sd = synth_state_dict([(k, tuple(v.shape)) for k, v in cur.items()], seed, gain)
    module.load_state_dict(sd)
    return module


def synth_db_embeddings(seed: int, n: int, d: int = 256) -> torch.Tensor:
    """Retrieval micro-benchmark DB: unit-norm NON-NEGATIVE rows (cell embeddings are post-ReLU)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g).abs_()
    return torch.nn.functional.normalize(x)


def synth_query_embeddings(seed: int, n: int, d: int = 256) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n, d, generator=g))


def superglue_peaky_(sd: Dict[str, torch.Tensor], prefix: str = "", scale: float = 9.0) -> Dict[str, torch.Tensor]:
    """Reshape random SuperGlue weights so that planted correspondences survive to the score matrix.

    Purely random weights collapse all descriptors onto a common direction (uniform scores, no matches);
    shrinking each layer's residual delta and making ``final_proj`` a scaled identity plus noise keeps
    ``scores ~ scale^2/sqrt(D) * <desc0_i, desc1_j>`` so that the matching logic is exercised.
    """
    for k in list(sd.keys()):
        if k.startswith(prefix) and k.endswith("mlp.3.weight"):
            sd[k] = sd[k] * 0.1
    w = sd[prefix + "final_proj.weight"]
    D = w.shape[0]
    sd[prefix + "final_proj.weight"] = (scale * torch.eye(D) + 0.3 * w.squeeze(-1)).reshape(w.shape).contiguous()
    return sd


def synth_descriptor_pairs(seed: int, B: int, M: int, N: int, D: int, noise: float = 0.05):
    """Unit-norm desc0 [B,M,D], desc1 [B,N,D] with planted correspondences desc1[j] ~ desc0[perm[j]] (+ some outliers)."""
    g = torch.Generator().manual_seed(seed)
    d0 = torch.nn.functional.normalize(torch.randn(B, M, D, generator=g), dim=-1)
    d1 = torch.empty(B, N, D)
    for b in range(B):
        perm = torch.randperm(M, generator=g)[:N]
        d1[b] = d0[b, perm] + noise * torch.randn(N, D, generator=g)
        n_out = int(torch.randint(0, max(N // 3, 1) + 1, (1,), generator=g))
        if n_out:
            d1[b, :n_out] = torch.randn(n_out, D, generator=g)
    return d0, torch.nn.functional.normalize(d1, dim=-1)
