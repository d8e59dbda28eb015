"""``torch_geometric.data.{Data, Batch}`` stand-ins: attribute containers of per-node tensors (``x`` = rgb, ``pos`` = xyz)
and the concatenation of a list of them with a ``batch`` vector (``Batch.from_data_list``), as used by
``dataloading/kitti360pose/utils.py:99-109`` of the reference."""
import torch


class Data:
    def __init__(self, x=None, pos=None, y=None, **kwargs):
        self.x, self.pos, self.y = x, pos, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        for t in (self.pos, self.x):
            if t is not None:
                return int(t.shape[0])
        return 0

    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def __iter__(self):
        for k in self.keys():
            yield k, getattr(self, k)

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def to(self, device, *a, **kw):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, *a, **kw))
        return self

    def clone(self):
        return type(self)(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})

    def __repr__(self):
        return f"{type(self).__name__}(" + ", ".join(f"{k}={list(v.shape) if torch.is_tensor(v) else v}" for k, v in self) + ")"


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        assert len(data_list) >= 1
        out = cls()
        keys = data_list[0].keys()
        for k in keys:
            vals = [getattr(d, k) for d in data_list]
            setattr(out, k, torch.cat(vals, dim=0) if torch.is_tensor(vals[0]) else vals)
        out.batch = torch.cat([torch.full((d.num_nodes,), i, dtype=torch.long) for i, d in enumerate(data_list)])
        out.ptr = torch.tensor([0] + [d.num_nodes for d in data_list]).cumsum(0)
        return out

    @property
    def num_graphs(self):
        return int(self.ptr.numel() - 1) if getattr(self, "ptr", None) is not None else 0


class DataLoader(torch.utils.data.DataLoader):
    """``torch_geometric.data.DataLoader`` (imported by ``dataloading/kitti360pose/objects.py:12``): collates ``Data`` lists."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=Batch.from_data_list, **kwargs)
