"""GPU check + timing of bk_conv3x3_bf16 against torch conv2d (bf16-rounded operands, fp32 reference)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
import torch
import torch.nn.functional as F
from blokus_self_play.tc_resnet import to_padded_nhwc, from_padded_nhwc, conv3x3

torch.manual_seed(0)
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x = torch.relu(torch.randn(B, 256, 20, 20, device=dev))
w = torch.randn(256, 256, 3, 3, device=dev) * 0.02
bias = torch.randn(256, device=dev) * 0.1
xb = x.to(torch.bfloat16).float()
wb = w.to(torch.bfloat16).float()
ref = F.conv2d(xb, wb, bias, padding=1)
xp = to_padded_nhwc(x)
w9 = wb.permute(2, 3, 0, 1).reshape(9, 256, 256).to(torch.bfloat16).contiguous()
y = conv3x3(xp, w9, bias.float().contiguous(), None, False, B)
torch.cuda.synchronize()
got = from_padded_nhwc(y, B)
err = (got - ref).abs().max().item()
scale = ref.abs().max().item()
print(f"B={B} max|err|={err:.4e} max|ref|={scale:.3f} rel={err/scale:.3e}")
pad_ok = bool((y.reshape(B, 21, 21, 256)[:, 20].abs().max() == 0) and (y.reshape(B, 21, 21, 256)[:, :, 20].abs().max() == 0))
print("pad rows zero:", pad_ok)
# fused residual + relu
r = torch.relu(torch.randn(B, 256, 20, 20, device=dev))
rp = to_padded_nhwc(r)
y2 = conv3x3(xp, w9, bias.float().contiguous(), rp, True, B)
ref2 = torch.relu(ref + r.to(torch.bfloat16).float())
err2 = (from_padded_nhwc(y2, B) - ref2).abs().max().item()
print(f"residual+relu max|err|={err2:.4e}")
ok = err / scale < 2e-2 and err2 / scale < 2e-2 and pad_ok
print("CONV_OK" if ok else "CONV_FAIL")
if ok and len(sys.argv) > 2:
    B2 = int(sys.argv[2])
    xp = to_padded_nhwc(torch.relu(torch.randn(B2, 256, 20, 20, device=dev)))
    out = torch.empty_like(xp)
    for _ in range(3):
        conv3x3(xp, w9, bias.float().contiguous(), None, True, B2, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for _ in range(n):
        conv3x3(xp, w9, bias.float().contiguous(), None, True, B2, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2 * B2 * 400 * 256 * 256 * 9
    print(f"B={B2} conv {ms:.3f} ms  {flops/ms/1e9:.1f} TFLOP/s (useful, 400 positions/image)")
    xb = torch.relu(torch.randn(B2, 256, 20, 20, device=dev)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(256, 256, 3, padding=1).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(3): conv(xb)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): conv(xb)
        e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / n
    print(f"cuDNN bf16 channels_last conv {ms2:.3f} ms  {flops/ms2/1e9:.1f} TFLOP/s")
