"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference ``NuScenesDataset`` class
(det3d/datasets/nuscenes/nuscenes.py) with stub modules for what this container lacks, so that its
``get_sensor_data`` can be executed on temporary files. Stubbed: ``pyquaternion.Quaternion`` (only ``rotation_matrix``
is used by the path: the standard unit-quaternion formula), ``det3d.datasets.custom.PointCloudDataset``,
``det3d.datasets.nuscenes.nusc_common`` (names only), ``det3d.datasets.registry.DATASETS``."""
import importlib.util
import os
import sys
import types

import numpy as np


class Quaternion:
    def __init__(self, q):
        self.q = np.asarray(q, dtype=np.float64)

    @property
    def rotation_matrix(self):
        q = self.q / np.linalg.norm(self.q)
        w, x, y, z = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def load(ref_root="/root/reference"):
    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("pyquaternion", Quaternion=Quaternion)

    class _Reg:
        def register_module(self, cls):
            return cls

    for pkg in ("det3d", "det3d.datasets", "det3d.datasets.nuscenes"):
        mod(pkg)
    mod("det3d.datasets.custom", PointCloudDataset=type("PointCloudDataset", (), {}))
    mod("det3d.datasets.nuscenes.nusc_common", general_to_detection={}, cls_attr_dist={},
        _second_det_to_nusc_box=None, _lidar_nusc_box_to_global=None, eval_main=None)
    mod("det3d.datasets.registry", DATASETS=_Reg())
    path = os.path.join(ref_root, "det3d", "datasets", "nuscenes", "nuscenes.py")
    spec = importlib.util.spec_from_file_location("ref_nuscenes_dataset", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m
