"""Hand-written tensor-core leaf evaluator: the ResNet trunk's 3x3 convolutions run on
`bk_conv3x3_bf16` (tcgen05 / TMEM / TMA, csrc/bk_conv.cu); the 5-channel input convolution and the two tiny heads
stay in PyTorch.  BatchNorm (eval mode) is folded into the convolution weights and bias:
    y = gamma * (conv(x) + b - mean) / sqrt(var + eps) + beta  =  conv_{w * s}(x) + ((b - mean) * s + beta),  s = gamma / sqrt(var + eps)
Activations stay in the kernel's zero-padded NHWC bf16 layout [batch*441][256] for the whole trunk."""
from __future__ import annotations

import ctypes as C
from typing import Optional

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


def conv3x3_in(x: torch.Tensor, w9: torch.Tensor, bias: torch.Tensor, relu: bool, batch: int, out: torch.Tensor, lib=None):
    """Narrow-input form (in_channels = x.shape[1] in {64,128,192,256}), no residual."""
    lib = lib or _lib.default_lib()
    stream = torch.cuda.current_stream(x.device).cuda_stream
    lib.check(lib.bk_conv3x3_bf16_in(C.c_void_p(x.data_ptr()), C.c_void_p(w9.data_ptr()), C.c_void_p(bias.data_ptr()),
                                     C.c_void_p(out.data_ptr()), batch, x.shape[1], 1 if relu else 0, C.c_void_p(stream)))
    return out


def _bn_scalar(bn: torch.nn.BatchNorm2d):
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return s, bn.bias - bn.running_mean * s


class TensorCoreLeafEvaluator:
    """Callable evaluator for SelfPlay.run_evaluator; same contract as resnet.LeafEvaluator (eval mode).

    Every 3x3 convolution — the 5->256 input layer (planes zero-extended to 64 channels) and the 2*blocks
    trunk layers — runs on the hand-written tcgen05 kernel, activations never leave the padded NHWC bf16
    layout; the two 1x1 head convolutions are one [B*441, 256] x [256, 2] product on that layout, and the
    remaining head arithmetic (scalar BN, ReLU, masked softmax, Linear(400, 4), tanh, softmax) is on B x 400."""

    def __init__(self, model: ResNet, lib=None):
        if model.width != 256:
            raise ValueError("the tcgen05 trunk kernel is built for width 256 (BASELINE.json config 4)")
        self.model = model.eval()
        self.lib = lib or _lib.default_lib()
        with torch.no_grad():
            self.blocks = [(fold_conv_bn(b.conv1, b.bn1), fold_conv_bn(b.conv2, b.bn2)) for b in model.res_blocks]
            w_in = torch.zeros((256, 64, 3, 3), dtype=model.input.weight.dtype, device=model.input.weight.device)
            w_in[:, :5] = model.input.weight
            self.w_in = w_in.permute(2, 3, 0, 1).reshape(9, 256, 64).to(torch.bfloat16).contiguous()
            self.b_in = model.input.bias.float().contiguous()
            pc, pbn = model.policy_head[0], model.policy_head[1]
            vc, vbn = model.value_head[0], model.value_head[1]
            self.w_heads = torch.cat([pc.weight.reshape(1, 256), vc.weight.reshape(1, 256)], 0).t().to(torch.bfloat16).contiguous()
            ps, pb = _bn_scalar(pbn)
            vs, vb = _bn_scalar(vbn)
            self.head_scale = torch.cat([ps, vs]).float()
            self.head_shift = torch.cat([pc.bias * ps + pb, vc.bias * vs + vb]).float()
            self.lin = model.value_head[4]
        self._bufs = {}

    @torch.no_grad()
    def __call__(self, planes: torch.Tensor):
        batch = planes.shape[0]
        key = planes.device
        if key not in self._bufs or self._bufs[key][2] < batch:      # grow-only buffers: batch sizes vary round to round
            rows = batch * PAD * PAD
            self._bufs[key] = ([torch.zeros((rows, 256), dtype=torch.bfloat16, device=planes.device) for _ in range(3)],
                               torch.zeros((rows, 64), dtype=torch.bfloat16, device=planes.device), batch)
        (a, t, b), x64, _cap = self._bufs[key]
        rows = batch * PAD * PAD
        a, t, b, x64 = a[:rows], t[:rows], b[:rows], x64[:rows]
        x64.view(batch, PAD, PAD, 64)[:, :20, :20, :5] = planes.permute(0, 2, 3, 1).to(torch.bfloat16)
        conv3x3_in(x64, self.w_in, self.b_in, False, batch, out=a, lib=self.lib)      # model.input (no BN / ReLU, resnet.py:79)
        for (w1, b1), (w2, b2) in self.blocks:
            conv3x3(a, w1, b1, None, True, batch, out=t, lib=self.lib)                 # relu(bn1(conv1(x)))
            conv3x3(t, w2, b2, a, True, batch, out=b, lib=self.lib)                    # relu(bn2(conv2(.)) + x)
            a, b = b, a
        heads = (a @ self.w_heads).float().view(batch, PAD, PAD, 2)[:, :20, :20, :]     # both 1x1 head convolutions
        heads = torch.relu(heads * self.head_scale + self.head_shift).reshape(batch, 400, 2)
        legal = planes[:, 4].reshape(batch, -1)
        logits = heads[:, :, 0]
        policy = torch.softmax(logits * legal + (1 - legal) * -1e9, dim=1) * legal
        value = torch.softmax(torch.tanh(self.lin(heads[:, :, 1])), dim=1)
        return policy, value
