"""Stand-in for the ``easydict`` package (``requirements.txt:1`` of the reference; not installed in this image).

Only found when the real package is absent: ``text2pos_cvpr2022_b200.compat.install()`` appends ``shims/`` to the END of
``sys.path``.  ``EasyDict`` here covers what the reference uses on the hot path (``models/superglue_matcher.py:119-126``,
``evaluation/pipeline.py:6``, ``dataloading/kitti360pose/eval.py:12``): construction from a dict / kwargs, key AND
attribute access, key iteration, nested dicts converted recursively."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = {} if d is None else dict(d)
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setattr__(self, k, v):
        dict.__setitem__(self, k, self._wrap(v))

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k, v in d.items():
            setattr(self, k, v)


__all__ = ["EasyDict"]
