"""Stand-ins for the GUI / transport modules the reference imports at module level but that are not on the numeric
path (open3d windows, matplotlib figures, mayavi, rclpy, sensor_msgs, easydict), so that its entry scripts can run
head-less on a machine that has only NumPy + PyTorch + the engine.  Only modules that are really absent are stubbed;
an installed open3d / matplotlib is left alone.  Every stub attribute is a permissive no-op object.
"""
from __future__ import annotations

import importlib.util
import sys
import types


class _Anything:
    """Callable, attribute-able, iterable-of-nothing placeholder (figures, axes, windows, QoS profiles ...)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _StubModule(types.ModuleType):
    _dsc_stub = True

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _PointCloudHolder:
    """o3d.geometry.PointCloud: an object with settable ``points`` / ``colors`` (scaner_version2.py:122-123)."""
    points = None
    colors = None

    @staticmethod
    def voxel_down_sample(pcd, voxel_size):
        # scaner_version2.py:127 / scan_context_node.py:98: the one numeric Open3D call of the reference -> K7
        from .engine import default_engine
        out = _PointCloudHolder()
        import numpy as np
        out.points = default_engine(0).voxel_downsample_host(np.asarray(pcd.points), float(voxel_size))[0]
        return out


def _absent(name: str) -> bool:
    if name in sys.modules:
        return bool(getattr(sys.modules[name], "_dsc_stub", False))
    try:
        return importlib.util.find_spec(name) is None
    except (ValueError, ImportError):
        return True


def _stub(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if not isinstance(m, _StubModule):               # absent, or somebody else's bare placeholder
        m = _StubModule(name)
        m.__path__ = []            # importable as a package
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def install_gui_stubs() -> list:
    """Stub every absent GUI / transport dependency of the reference.  Returns the names that were stubbed."""
    done = []
    if _absent("open3d"):
        _stub("open3d")
        _stub("open3d.geometry", PointCloud=_PointCloudHolder)
        _stub("open3d.utility", Vector3dVector=lambda a: a)
        _stub("open3d.visualization")
        done.append("open3d")
    if _absent("matplotlib"):
        _stub("matplotlib", use=lambda *a, **k: None)
        for sub in ("pyplot", "cm", "colors", "patches", "image"):
            _stub("matplotlib." + sub)
        _stub("mpl_toolkits")
        _stub("mpl_toolkits.mplot3d", Axes3D=_Anything)
        done.append("matplotlib")
    if _absent("mayavi"):
        _stub("mayavi")
        _stub("mayavi.mlab")
        done.append("mayavi")
    if _absent("rclpy"):
        _stub("rclpy")
        _stub("rclpy.node", Node=object)
        _stub("rclpy.qos")
        done.append("rclpy")
    if _absent("sensor_msgs"):
        from . import ros
        _stub("sensor_msgs")
        _stub("sensor_msgs.msg", PointCloud2=ros.PointCloud2, PointField=ros.PointField)
        done.append("sensor_msgs")
    if _absent("sensor_msgs_py"):
        _stub("sensor_msgs_py")
        _stub("sensor_msgs_py.point_cloud2")
        done.append("sensor_msgs_py")
    if _absent("easydict"):
        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                for k, v in dict(d or {}, **kw).items():
                    self[k] = v

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v
        _stub("easydict", EasyDict=EasyDict)
        done.append("easydict")
    return done
