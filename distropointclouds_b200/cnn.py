"""Config 4: scan contexts born on the GPU -> K4 (to_tensor01) -> the reference's TinyCNN -> masks.

The network itself is the reference's (distro_model/tiny_sc_ped.py:134-154, weights
``tinycnn_best.pth``) and stays a stock PyTorch/cuDNN module -- PyTorch is plumbing here, the
engine's job is to hand it device-resident, already normalised grids so that the host round trip of
``run_inference`` (tiny_sc_ped.py:217-234: np.load -> to_tensor01 -> .to(device)) disappears.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _ffi
from .engine import Engine, ScanContextParams, default_engine, pack_scans

# (kind, in, out) per layer of tiny_sc_ped.TinyCNN; 'P' = MaxPool2d(2).  Index positions inside the two
# nn.Sequential containers match the reference so that its state_dict loads with strict=True.
_ENC = (("c3", 1, 16), ("c3", 16, 32), ("P",), ("c3", 32, 64), ("P",), ("c3", 64, 128), ("P",))
_DEC = (("t2", 128, 64), ("t2", 64, 32), ("t2", 32, 16), ("c1", 16, 1))


def build_tiny_cnn(in_ch: int = 1):
    """An nn.Module with the parameter names/shapes of tiny_sc_ped.TinyCNN (enc.*, dec.*)."""
    import torch.nn as nn

    def stack(spec, first_in):
        mods = []
        for item in spec:
            if item[0] == "P":
                mods.append(nn.MaxPool2d(2))
                continue
            kind, cin, cout = item
            cin = first_in if (first_in is not None and not mods) else cin
            if kind == "c3":
                mods += [nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
            elif kind == "t2":
                mods += [nn.ConvTranspose2d(cin, cout, 2, stride=2), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
            else:
                mods.append(nn.Conv2d(cin, cout, 1))
        return nn.Sequential(*mods)

    class TinyCNN(nn.Module):
        def __init__(self):
            super().__init__()
            self.enc = stack(_ENC, in_ch)
            self.dec = stack(_DEC, None)

        def forward(self, x):
            return self.dec(self.enc(x))

    return TinyCNN()


def fold_batchnorm(model):
    """Eval-mode rewrite: every (Conv2d | ConvTranspose2d) -> BatchNorm2d pair becomes one convolution with
    the BN affine folded into its weights and bias (same function up to rounding; 7 fewer kernels and 7 fewer
    passes over the activations per forward).  Returns a new module; the input is left untouched."""
    import copy
    import torch
    import torch.nn as nn
    m = copy.deepcopy(model).eval()
    for seq_name in ("enc", "dec"):
        seq = getattr(m, seq_name)
        mods = list(seq.children())
        out = []
        i = 0
        while i < len(mods):
            cur = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(cur, (nn.Conv2d, nn.ConvTranspose2d)) and isinstance(nxt, nn.BatchNorm2d):
                with torch.no_grad():
                    scale = nxt.weight / torch.sqrt(nxt.running_var + nxt.eps)
                    shape = (-1, 1, 1, 1) if isinstance(cur, nn.Conv2d) else (1, -1, 1, 1)   # out-channel axis
                    cur.weight.mul_(scale.reshape(shape))
                    bias = cur.bias if cur.bias is not None else torch.zeros_like(nxt.running_mean)
                    cur.bias = nn.Parameter((bias - nxt.running_mean) * scale + nxt.bias)
                out.append(cur)
                i += 2
            else:
                out.append(cur)
                i += 1
        setattr(m, seq_name, nn.Sequential(*out))
    return m


class ScanContextCnnPipeline:
    """points (device or host) -> grids -> normalised grids -> logits -> probabilities/masks, all on one GPU.

    precision: "fp32" (cuDNN fp32, TF32 off, BatchNorm layers as they are: the parity mode), "tf32" (PyTorch's
    default for convolutions), "fp16" or "bf16" (weights and activations in half precision).  The three fast
    modes fold every BatchNorm into its convolution (eval mode, same function up to rounding).  Parity is
    defined on fp32 logits; the faster modes are opt-in.  Measured on B200
    (profiles/microbench/cnn_modes.py): fp32 0.185, tf32+fold 0.103, bf16+fold 0.093, fp16+fold 0.086 ms/scan.
    """

    def __init__(self, weights_path: Optional[str] = None, device: int = 0, precision: str = "fp32",
                 engine: Optional[Engine] = None, chunk: int = 16, model=None):
        import torch
        if precision not in ("fp32", "tf32", "bf16", "fp16"):
            raise ValueError("precision must be fp32, tf32, fp16 or bf16")
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.engine = engine if engine is not None else default_engine(device)
        self.precision = precision
        self.chunk = int(chunk)
        self.model = model if model is not None else build_tiny_cnn()
        if weights_path is not None:
            sd = torch.load(weights_path, map_location="cpu", weights_only=True)      # tiny_sc_ped.py:221
            self.model.load_state_dict(sd, strict=True)
        self.model.eval()
        if precision != "fp32":
            self.model = fold_batchnorm(self.model)
        self._dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}.get(precision, torch.float32)
        self.model = self.model.to(self.device).to(self._dtype).to(memory_format=torch.channels_last)

    # -- stages ------------------------------------------------------------------------------------
    def build_grids(self, pts, offsets: np.ndarray, params: ScanContextParams, want_norm: bool = True,
                    stride_bytes: int = 16, strict: bool = False):
        """K1 (+K4) on a device-resident packed point tensor.  Returns (grids, norm|None) CUDA tensors.
        strict: per-scan counters are kept and read back (one small D2H), and an IndexError is raised where the
        reference's variant A would raise one -- NaN / Inf coordinates or an angle that rounds to 360.0
        (scaner_version2.py:93-107); without it such points are dropped silently."""
        torch = self.torch
        n = len(offsets) - 1
        grids = torch.empty((n, params.num_ring, params.num_sector), dtype=torch.float32, device=self.device)
        norm = torch.empty_like(grids) if want_norm else None
        stats = torch.zeros((n, len(_ffi.STATS_FIELDS)), dtype=torch.int64, device=self.device) if strict else None
        self.engine.build_device(pts, offsets, params, grids, norm, stats, stride_bytes)
        if strict and params.variant == _ffi.VARIANT_A:
            st = stats.cpu().numpy()
            bad = np.flatnonzero((st[:, 1] + st[:, 2]) > 0)
            if bad.size:
                k = int(bad[0])
                raise IndexError(f"scan {k}: {int(st[k, 1])} point(s) with non-finite polar coordinates, "
                                 f"{int(st[k, 2])} whose angle rounds to 360.0 (the reference raises IndexError here)")
        return grids, norm

    def logits_from_norm(self, norm):
        """norm: (n,R,S) float32 CUDA tensor in [0,1] -> (n,R,S) float32 logits."""
        torch = self.torch
        out = torch.empty_like(norm)
        tf32 = self.precision != "fp32"
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            with torch.no_grad():
                for a in range(0, norm.shape[0], self.chunk):
                    # C == 1: NCHW and NHWC are the same bytes, the format call only retags the strides
                    x = norm[a:a + self.chunk].unsqueeze(1).to(self._dtype).contiguous(memory_format=torch.channels_last)
                    out[a:a + self.chunk] = self.model(x)[:, 0].float()
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        return out

    # -- whole pipeline ----------------------------------------------------------------------------
    def infer_device(self, pts, offsets: np.ndarray, params: ScanContextParams = ScanContextParams(),
                     thr: float = 0.5, stride_bytes: int = 16, strict: bool = False):
        """Returns (grids, logits, mask uint8) as CUDA tensors; nothing touches the host (unless strict)."""
        grids, norm = self.build_grids(pts, offsets, params, True, stride_bytes, strict)
        logits = self.logits_from_norm(norm)
        mask = (self.torch.sigmoid(logits) > thr).to(self.torch.uint8)                 # tiny_sc_ped.py:227-228
        return grids, logits, mask

    def infer_scans(self, scans: Sequence[np.ndarray], params: ScanContextParams = ScanContextParams(),
                    thr: float = 0.5, strict: bool = True):
        """Host scans (list of (N,3|4) float32) -> (grids, probs, masks) NumPy arrays, like run_inference
        (tiny_sc_ped.py:217-234) for a batch; strict: raise where the reference's grid builder would."""
        torch = self.torch
        pts, offsets, stride = pack_scans(list(scans))
        d_pts = torch.from_numpy(np.ascontiguousarray(pts)).to(self.device)
        grids, logits, mask = self.infer_device(d_pts, offsets, params, thr, stride, strict)
        return grids.cpu().numpy(), torch.sigmoid(logits).cpu().numpy(), mask.cpu().numpy()


def mask_to_detections(mask: np.ndarray, min_pixels: int = 20):
    """tiny_sc_ped.py:196-215: connected components of the binary mask -> centre cells (host, scipy)."""
    from scipy import ndimage as ndi
    lab, n = ndi.label(np.asarray(mask) > 0)
    dets = []
    for i in range(1, n + 1):
        coords = np.argwhere(lab == i)
        if coords.shape[0] < min_pixels:
            continue
        r_mean, s_mean = coords[:, 0].mean(), coords[:, 1].mean()
        dets.append({"center_ring_idx": float(r_mean), "center_sector_idx": float(s_mean),
                     "area_pixels": int(coords.shape[0])})
    return dets
