"""Run one of the reference's entry scripts, unmodified, on the CUDA engine:

    python -m distropointclouds_b200.run /path/to/DistroPointclouds/scaner_version2.py
    python -m distropointclouds_b200.run /path/to/DistroPointclouds/reconstruct_pointcloud.py
    python -m distropointclouds_b200.run /path/to/DistroPointclouds/distro_model/data_preparation.py [its flags]
    python -m distropointclouds_b200.run /path/to/DistroPointclouds/scan_context_node.py

What the runner does, in this order:
  1. stubs GUI / transport modules that are not installed (open3d, matplotlib, mayavi, rclpy, ...: headless.py);
  2. executes the script's module body -- everything except its ``if __name__ == "__main__":`` block -- so that its
     classes exist, then ``compat.install()`` patches the hot-path methods of every reference class now imported
     (the script's own and e.g. ``data_process.kitti_dataset.ScanContext``) and turns the blocking viewers off;
  3. makes ``torch.utils.data.DataLoader`` run its workers in the parent process: the reference builds its loader with
     ``num_workers=4`` fork workers (config/train_config.py), and a CUDA context cannot be used in a forked child --
     ``KittiDataset.__getitem__`` (which generates the scan contexts as a side effect) therefore has to run where the
     engine lives;
  4. executes the ``__main__`` block in the module's namespace, with the script's directory as working directory and
     ``sys.argv`` set to the remaining arguments.
"""
from __future__ import annotations

import ast
import os
import sys
import types
from typing import List, Optional


def split_main_guard(source: str, filename: str):
    """-> (code of the module body without its main guard, code of the guard's body or None)."""
    tree = ast.parse(source, filename)
    body, main = [], []
    for node in tree.body:
        t = node.test if isinstance(node, ast.If) else None
        is_guard = (isinstance(t, ast.Compare) and isinstance(t.left, ast.Name) and t.left.id == "__name__"
                    and len(t.ops) == 1 and isinstance(t.ops[0], ast.Eq) and len(t.comparators) == 1
                    and isinstance(t.comparators[0], ast.Constant) and t.comparators[0].value == "__main__")
        if is_guard:
            main.extend(node.body)
        else:
            body.append(node)
    mod = ast.Module(body=body, type_ignores=[])
    mn = ast.Module(body=main, type_ignores=[]) if main else None
    return compile(mod, filename, "exec"), (compile(mn, filename, "exec") if mn else None)


def force_inprocess_dataloader() -> bool:
    """DataLoader(num_workers>0) forks; CUDA cannot be used after fork -> run the workers' job in the parent."""
    try:
        import torch.utils.data as tud
    except Exception:
        return False
    if getattr(tud.DataLoader, "_dsc_patched", False):
        return True
    orig = tud.DataLoader.__init__

    def __init__(self, *a, **kw):
        if kw.get("num_workers", 0):
            print(f"[dsc.run] DataLoader(num_workers={kw['num_workers']}) -> 0: the dataset calls the CUDA engine, "
                  "which cannot be used from forked workers", file=sys.stderr)
            kw["num_workers"] = 0
            kw.pop("prefetch_factor", None)
            kw.pop("persistent_workers", None)
        orig(self, *a, **kw)
    tud.DataLoader.__init__ = __init__
    tud.DataLoader._dsc_patched = True
    return True


def run_script(path: str, argv: Optional[List[str]] = None, device: int = 0, stubs: bool = True, patch_voxel: bool = True):
    from . import compat, headless
    path = os.path.abspath(path)
    if stubs:
        headless.install_gui_stubs()
    with open(path, "r") as f:
        body, main = split_main_guard(f.read(), path)
    mod = types.ModuleType("__dsc_main__")
    mod.__file__ = path
    here = os.path.dirname(path)
    old_argv, old_cwd, old_path0 = sys.argv, os.getcwd(), list(sys.path)
    sys.modules[mod.__name__] = mod
    sys.argv = [path] + list(argv or [])
    sys.path.insert(0, here)
    os.chdir(here)
    try:
        exec(body, mod.__dict__)                         # classes and functions exist now; nothing has run yet
        for cls in (getattr(mod, "ScanContext", None),):
            if isinstance(cls, type) and hasattr(cls, "viz"):
                cls.viz = 0                              # scaner_version2.py:28 defaults to blocking Open3D windows
        patched = compat.install(device_index=device, patch_voxel=patch_voxel)
        force_inprocess_dataloader()
        print(f"[dsc.run] patched: {patched}", file=sys.stderr)
        mod.__dict__["__name__"] = "__main__"
        if main is not None:
            exec(main, mod.__dict__)
        return mod
    finally:
        compat.uninstall()
        sys.argv = old_argv
        os.chdir(old_cwd)
        sys.path[:] = old_path0
        sys.modules.pop("__dsc_main__", None)


def main():
    import argparse
    ap = argparse.ArgumentParser(prog="python -m distropointclouds_b200.run", description=__doc__.split("\n\n")[0])
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--no-stubs", action="store_true", help="do not stub absent GUI modules")
    ap.add_argument("--keep-open3d-voxel", action="store_true", help="leave _downsample_voxel to an installed Open3D")
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    run_script(a.script, a.script_args, a.device, not a.no_stubs, not a.keep_open3d_voxel)


if __name__ == "__main__":
    main()
