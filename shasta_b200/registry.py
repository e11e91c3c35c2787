"""det3d registry / builder interface kept for the affinity path (drop-in boundary).

Semantics restated from the reference:
  * ``Registry`` keys classes by ``__name__``; duplicate -> KeyError, non-class -> TypeError
    (det3d/utils/registry.py:28-46).
  * ``build_from_cfg(cfg, registry, default_args)`` pops ``type`` (a registered name or a class), fills missing
    keyword arguments from ``default_args`` and calls the class (det3d/utils/registry.py:49-78).
  * ``build`` turns a list of cfgs into ``nn.Sequential`` (det3d/models/builder.py:20-25); ``build_track`` /
    ``build_simp_track`` inject ``train_cfg`` / ``test_cfg`` (det3d/models/builder.py:70-75).
"""
import inspect

from torch import nn


class Registry:
    """Name -> class table. ``register_module`` is used as a class decorator."""

    __slots__ = ("name", "_table")

    def __init__(self, name):
        self.name = name
        self._table = {}

    def __repr__(self):
        return "Registry(name=%s, items=%s)" % (self.name, sorted(self._table))

    def __contains__(self, key):
        return key in self._table

    @property
    def module_dict(self):
        return self._table

    def get(self, key):
        """Registered class or None (callers turn None into their own KeyError)."""
        return self._table.get(key)

    def register_module(self, cls):
        if not inspect.isclass(cls):
            raise TypeError("module must be a class, but got {}".format(type(cls)))
        if cls.__name__ in self._table:
            raise KeyError("{} is already registered in {}".format(cls.__name__, self.name))
        self._table[cls.__name__] = cls
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    """cfg: dict with "type" (registered name or a class) + constructor kwargs; default_args fill the gaps."""
    if not (isinstance(cfg, dict) and "type" in cfg):
        raise AssertionError("cfg must be a dict with a 'type' key")
    if not (default_args is None or isinstance(default_args, dict)):
        raise AssertionError("default_args must be a dict or None")
    kwargs = dict(cfg)
    kind = kwargs.pop("type")
    if inspect.isclass(kind):
        cls = kind
    elif isinstance(kind, str):
        cls = registry.get(kind)
        if cls is None:
            raise KeyError("{} is not in the {} registry".format(kind, registry.name))
    else:
        raise TypeError("type must be a str or valid type, but got {}".format(type(kind)))
    for key, value in (default_args or {}).items():
        kwargs.setdefault(key, value)
    return cls(**kwargs)


# the registries the affinity path touches (det3d/models/registry.py:3-14)
READERS = Registry("reader")
BACKBONES = Registry("backbone")
NECKS = Registry("neck")
TRACK = Registry("track")
SECOND_STAGE = Registry("second_stage")


def build(cfg, registry, default_args=None):
    if isinstance(cfg, list):
        return nn.Sequential(*[build_from_cfg(c, registry, default_args) for c in cfg])
    return build_from_cfg(cfg, registry, default_args)


def build_second_stage_module(cfg):
    return build(cfg, SECOND_STAGE)


def build_reader(cfg):
    return build(cfg, READERS)


def build_backbone(cfg):
    return build(cfg, BACKBONES)


def build_neck(cfg):
    return build(cfg, NECKS)


def build_track(cfg, train_cfg=None, test_cfg=None):
    return build(cfg, TRACK, dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_simp_track(cfg, train_cfg=None, test_cfg=None):
    return build(cfg, TRACK, dict(train_cfg=train_cfg, test_cfg=test_cfg))
