"""Hand-written tensor-core leaf evaluator: the reference's `ResNet(blocks, 256)`
(model/resnet.py:44-94, eval mode) as ONE native object of the C ABI (`bk_evaluator_*`, csrc/bk_eval.cu): input
packing, the 2*blocks+1 tcgen05 convolutions and the fused heads all run on this library's kernels — PyTorch is only
used here, once, to read the model's parameters and fold the BatchNorms:

    y = gamma * (conv(x) + b - mean) / sqrt(var + eps) + beta  =  conv_{w * s}(x) + ((b - mean) * s + beta),  s = gamma / sqrt(var + eps)

`TensorCoreLeafEvaluator` is callable like `resnet.LeafEvaluator` (planes -> policy, value on CUDA tensors) so it plugs
into `SelfPlay.run_evaluator`; `SelfPlay.run_network(evaluator)` runs whole games with no Python in the round at all.
The stand-alone operators (`conv3x3`, padded-layout helpers) remain for tests and probes."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .resnet import ResNet

PAD = 21


def to_padded_nhwc(x: torch.Tensor) -> torch.Tensor:
    """[B, C, 20, 20] -> bf16 [B*441, C] (row 20 / col 20 of every image are zeros)."""
    b, c = x.shape[0], x.shape[1]
    return F.pad(x, (0, 1, 0, 1)).permute(0, 2, 3, 1).reshape(b * PAD * PAD, c).to(torch.bfloat16).contiguous()


def from_padded_nhwc(y: torch.Tensor, batch: int) -> torch.Tensor:
    """bf16 [B*441, C] -> float32 [B, C, 20, 20]."""
    return y.reshape(batch, PAD, PAD, -1)[:, :20, :20, :].permute(0, 3, 1, 2).float().contiguous()


def fold_conv_bn(conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d):
    """-> (w bf16 [9][out][in], bias f32 [out]) with the eval-mode BatchNorm folded in."""
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = conv.weight * s[:, None, None, None]
    b = (conv.bias - bn.running_mean) * s + bn.bias
    w9 = w.permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1])           # tap = ky*3 + kx
    return w9.to(torch.bfloat16).contiguous(), b.float().contiguous()


def conv3x3(x: torch.Tensor, w9: torch.Tensor, bias: torch.Tensor, residual: Optional[torch.Tensor], relu: bool,
            batch: int, out: Optional[torch.Tensor] = None, lib=None) -> torch.Tensor:
    lib = lib or _lib.default_lib()
    y = out if out is not None else torch.empty_like(x)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    lib.check(lib.bk_conv3x3_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(w9.data_ptr()), C.c_void_p(bias.data_ptr()),
                                  C.c_void_p(residual.data_ptr()) if residual is not None else None,
                                  C.c_void_p(y.data_ptr()), batch, 1 if relu else 0, C.c_void_p(stream)))
    return y


def _bf16_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.bfloat16).contiguous().view(torch.int16).cpu().numpy()


def _f32_host(t: torch.Tensor) -> np.ndarray:
    return np.ascontiguousarray(t.detach().float().cpu().numpy())


def _head_affine(conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d):
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return s, conv.bias * s + bn.bias - bn.running_mean * s


class TensorCoreLeafEvaluator:
    """The network on the device as a `bk_evaluator`; callable on CUDA planes [R,5,20,20] -> (policy [R,400], value [R,4]).
    max_rows = the largest batch it will be asked to evaluate (activation buffers are sized for it; it grows on demand
    when called with a larger batch)."""

    def __init__(self, model: ResNet, lib=None, max_rows: int = 1024, device: Optional[int] = None):
        if model.width != 256:
            raise ValueError("the tcgen05 trunk kernel is built for width 256 (BASELINE.json config 4)")
        self.lib = lib or _lib.default_lib()
        self.model = model.eval()
        p = next(model.parameters())
        self.device = int(device if device is not None else (p.device.index or 0) if p.is_cuda else 0)
        with torch.no_grad():
            w_in = torch.zeros((256, 64, 3, 3), dtype=torch.float32, device=p.device)
            w_in[:, :5] = model.input.weight
            folded = [fold_conv_bn(c, b) for blk in model.res_blocks for c, b in ((blk.conv1, blk.bn1), (blk.conv2, blk.bn2))]
            ps, pb = _head_affine(model.policy_head[0], model.policy_head[1])
            vs, vb = _head_affine(model.value_head[0], model.value_head[1])
            lin = model.value_head[4]
            self._params = dict(
                w_in=_bf16_host(w_in.permute(2, 3, 0, 1).reshape(9, 256, 64)), b_in=_f32_host(model.input.bias),
                w_blk=np.stack([_bf16_host(w) for w, _ in folded]) if folded else np.zeros(1, dtype=np.int16),
                b_blk=np.stack([_f32_host(b) for _, b in folded]) if folded else np.zeros(1, dtype=np.float32),
                head_w=_f32_host(torch.cat([model.policy_head[0].weight.reshape(1, 256), model.value_head[0].weight.reshape(1, 256)])),
                head_affine=_f32_host(torch.cat([ps, pb, vs, vb])), lin_w=_f32_host(lin.weight), lin_b=_f32_host(lin.bias))
        self.blocks = len(model.res_blocks)
        self._h = None
        self._create(max_rows)

    def _create(self, max_rows: int) -> None:
        self.close()
        q = self._params
        h = C.c_void_p()
        ptr = lambda a: C.c_void_p(a.ctypes.data)
        self.lib.check(self.lib.bk_evaluator_create(self.device, self.blocks, int(max_rows), ptr(q["w_in"]), ptr(q["b_in"]),
                                                    ptr(q["w_blk"]), ptr(q["b_blk"]), ptr(q["head_w"]), ptr(q["head_affine"]),
                                                    ptr(q["lin_w"]), ptr(q["lin_b"]), C.byref(h)))
        self._h = h
        self.max_rows = int(max_rows)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self.lib.bk_evaluator_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def reserve(self, rows: int) -> None:
        if rows > self.max_rows:
            self._create(rows)

    def forward(self, planes: torch.Tensor, debug: bool = False):
        """`model(boards)`; with debug=True also the pre-softmax head outputs (policy_head(x) [R,400], value_head(x) [R,4])."""
        rows = int(planes.shape[0])
        self.reserve(rows)
        planes = planes.to(dtype=torch.float32).contiguous()
        policy = torch.empty((rows, 400), dtype=torch.float32, device=planes.device)
        value = torch.empty((rows, 4), dtype=torch.float32, device=planes.device)
        logits = torch.empty((rows, 400), dtype=torch.float32, device=planes.device) if debug else None
        vtanh = torch.empty((rows, 4), dtype=torch.float32, device=planes.device) if debug else None
        stream = torch.cuda.current_stream(planes.device).cuda_stream
        opt = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self.lib.check(self.lib.bk_evaluator_forward(self._h, C.c_void_p(planes.data_ptr()), rows, opt(policy), opt(value),
                                                     opt(logits), opt(vtanh), C.c_void_p(stream)))
        return (policy, value, logits, vtanh) if debug else (policy, value)

    def __call__(self, planes: torch.Tensor):
        return self.forward(planes)
