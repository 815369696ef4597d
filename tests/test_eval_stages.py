"""The leaf evaluator's non-convolution stages (csrc/bk_eval_kernels.cuh) against numpy: the planes written in the
first convolution's input layout (from a game state and from float planes) and the fused policy/value heads
(model/resnet.py:84-92).  Same kernel source on the CPU warp emulator here and on the device (-m gpu)."""
import ctypes as C

import numpy as np
import pytest

PAD, IMG = 21, 441


def bf16_bits(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def bf16_val(bits):
    return (bits.astype(np.uint32) << 16).view(np.float32)


def expected_x64(planes):
    n = planes.shape[0]
    x = np.zeros((n, PAD, PAD, 64), dtype=np.uint16)
    x[:, :20, :20, :5] = bf16_bits(planes.transpose(0, 2, 3, 1))
    return x.reshape(n * IMG, 64)


def _stages(lib, mem):
    """mem(array) -> (device pointer, read-back function): host memory on the emulator, torch CUDA tensors on the GPU."""
    from blokus_self_play import GameBatch
    b = GameBatch(3, lib=lib)
    b.playout(seed=4, max_plies=61)                                  # three different seats to move, mid-game, mid-turn
    planes = b.board_state().astype(np.float32)
    assert len(set(b.current_player().tolist())) >= 1 and planes[:, 4].sum() > 0
    want = expected_x64(planes)
    p, back = mem(np.zeros_like(want))
    lib.check(lib.bk_env_board_state_nhwc(b._h, C.c_void_p(p)))
    assert np.array_equal(back(), want), "kb_planes_nhwc differs from get_board_state"
    pp, _ = mem(planes)
    q, back2 = mem(np.zeros_like(want))
    lib.check(lib.bk_eval_pack_planes(C.c_void_p(pp), 3, C.c_void_p(q), None))
    if hasattr(lib, "bk_env_sync"):
        lib.check(lib.bk_env_sync(b._h))
    assert np.array_equal(back2(), want), "kb_pack_planes differs"
    # heads
    rng = np.random.default_rng(3)
    act_bits = bf16_bits(np.maximum(rng.normal(0.3, 0.6, size=(3 * IMG, 256)), 0).astype(np.float32))
    act = bf16_val(act_bits).reshape(3, PAD, PAD, 256)[:, :20, :20, :].reshape(3, 400, 256).astype(np.float64)
    head_w = rng.uniform(-1 / 16, 1 / 16, size=(2, 256)).astype(np.float32)
    affine = np.array([2.0, 0.4, 1.1, 0.3], dtype=np.float32)
    lin_w = rng.uniform(-0.05, 0.05, size=(4, 400)).astype(np.float32)
    lin_b = rng.uniform(-0.05, 0.05, size=4).astype(np.float32)
    params = np.concatenate([head_w.reshape(-1), affine, lin_w.reshape(-1), lin_b]).astype(np.float32)
    a_p, _ = mem(act_bits)
    h_p, _ = mem(params)
    outs = [mem(np.zeros(s, dtype=np.float32)) for s in ((3, 400), (3, 4), (3, 400), (3, 4))]
    lib.check(lib.bk_eval_heads(C.c_void_p(a_p), C.c_void_p(q), C.c_void_p(h_p), 3, *[C.c_void_p(o[0]) for o in outs], None))
    lib.check(lib.bk_env_sync(b._h))
    policy, value, logits, vtanh = [o[1]() for o in outs]
    lg = np.maximum(act @ head_w[0].astype(np.float64) * affine[0] + affine[1], 0)
    hv = np.maximum(act @ head_w[1].astype(np.float64) * affine[2] + affine[3], 0)
    mask = planes[:, 4].reshape(3, 400)
    z = np.where(mask > 0, lg, -np.inf)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    pol = e / e.sum(axis=1, keepdims=True)
    vt = np.tanh(hv @ lin_w.T.astype(np.float64) + lin_b)
    ev = np.exp(vt - vt.max(axis=1, keepdims=True))
    val = ev / ev.sum(axis=1, keepdims=True)
    assert np.allclose(logits, lg, atol=2e-5) and np.allclose(vtanh, vt, atol=2e-5)
    assert np.allclose(policy, pol, atol=2e-6) and np.allclose(value, val, atol=2e-6)
    assert np.all(policy[mask == 0] == 0) and np.allclose(policy.sum(axis=1), 1, atol=1e-5)
    b.close()


def test_emu_eval_stages(emu_lib):
    def mem(a):
        a = np.ascontiguousarray(a).copy()
        _keep.append(a)
        return a.ctypes.data, (lambda: a)
    _keep = []
    _stages(emu_lib, mem)


@pytest.mark.gpu
def test_gpu_eval_stages(cuda_lib):
    import torch

    def mem(a):
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).cuda()
        _keep.append(t)
        return t.data_ptr(), (lambda: (torch.cuda.synchronize(), t.cpu().numpy().view(a.dtype))[1])
    _keep = []
    _stages(cuda_lib, mem)
