"""Fused BatchNorm(+residual)+ReLU and global average pool kernels vs plain PyTorch fp32
references of the same ops, and a ResNet-18 step with the fused path on/off."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("shape", [(8, 64, 16, 16), (4, 128, 7, 9), (2, 24, 5, 5), (3, 512, 4, 4)])
@pytest.mark.parametrize("with_res,relu", [(False, True), (True, True), (False, False)])
def test_bn_act_matches_fp32_reference(shape, with_res, relu):
    from hefl_b200.ops import resnet_ops
    torch.manual_seed(0)
    B, C, H, W = shape
    dev = "cuda"
    x = _cl((torch.randn(shape, device=dev) * 1.5 + 0.3).to(torch.bfloat16))
    res = _cl(torch.randn(shape, device=dev).to(torch.bfloat16)) if with_res else None
    bn = nn.BatchNorm2d(C).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    bn_ref = nn.BatchNorm2d(C).to(dev)
    bn_ref.load_state_dict(bn.state_dict())
    g = _cl(torch.randn(shape, device=dev).to(torch.bfloat16))

    xa = x.clone().requires_grad_(True)
    ra = res.clone().requires_grad_(True) if with_res else None
    y = resnet_ops.bn_act(bn, xa, ra, relu)
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    y.backward(g)

    xb = x.float().requires_grad_(True)
    rb = res.float().requires_grad_(True) if with_res else None
    z = bn_ref(xb)
    if with_res:
        z = z + rb
    if relu:
        # the kernel masks with the bf16-rounded output; emulate that rounding for the mask
        z = torch.where(z.to(torch.bfloat16) > 0, z, torch.zeros_like(z))
    z.backward(g.float())

    assert torch.allclose(y.float(), z, atol=3e-2, rtol=2e-2)
    def close(a, b, tol):
        return (a.float() - b).norm() / (b.norm() + 1e-12) < tol
    assert close(xa.grad, xb.grad, 2e-2)
    if with_res:
        assert close(ra.grad, rb.grad, 1e-2)
    assert close(bn.weight.grad, bn_ref.weight.grad, 1e-2)
    assert close(bn.bias.grad, bn_ref.bias.grad, 1e-2)
    assert torch.allclose(bn.running_mean, bn_ref.running_mean, atol=1e-3)
    assert torch.allclose(bn.running_var, bn_ref.running_var, atol=2e-3, rtol=1e-3)
    assert int(bn.num_batches_tracked) == 1


def test_global_avgpool_matches_reference():
    from hefl_b200.ops import resnet_ops
    x = _cl(torch.randn(5, 96, 7, 7, device="cuda").to(torch.bfloat16)).requires_grad_(True)
    out = resnet_ops.global_avgpool(x)
    ref = x.detach().float().mean((2, 3))
    assert out.dtype == torch.float32 and torch.allclose(out, ref, atol=1e-5)
    g = torch.randn_like(out)
    out.backward(g)
    assert torch.allclose(x.grad.float(), (g / 49)[:, :, None, None].expand(5, 96, 7, 7), atol=1e-3, rtol=1e-2)


def test_resnet18_step_fused_vs_aten():
    """bf16 noise through 20 BN layers makes a fused-vs-ATen comparison loose, so both are
    scored against an fp32 run of the same model: the fused path must be as close as ATen's."""
    from hefl_b200 import _ext
    from hefl_b200.models import create_model
    from hefl_b200.ops import resnet_ops
    ops = _ext.ops()
    torch.manual_seed(1)
    ms = [create_model("resnet18", num_classes=10).cuda().train() for _ in range(3)]
    for m in ms[1:]:
        m.load_state_dict(ms[0].state_dict())
    x = torch.randn(16, 96, 96, 3, device="cuda").permute(0, 3, 1, 2)     # NHWC storage
    y = torch.randint(0, 10, (16,), device="cuda")

    def step(m, amp):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            logits = m(x).float()
        loss = F.cross_entropy(logits, y)
        loss.backward()
        return loss.item(), torch.cat([p.grad.flatten() for p in m.parameters()])

    n0 = int(ops.launch_count())
    l_f, g_f = step(ms[0], True)
    assert int(ops.launch_count()) - n0 >= 20 * 4        # every BN layer ran on our kernels
    resnet_ops.ENABLE = False
    try:
        l_a, g_a = step(ms[1], True)
    finally:
        resnet_ops.ENABLE = True
    l_r, g_r = step(ms[2], False)
    cos_f = F.cosine_similarity(g_f, g_r, dim=0).item()
    cos_a = F.cosine_similarity(g_a, g_r, dim=0).item()
    assert abs(l_f - l_r) < 5e-2 * max(1.0, abs(l_r))
    assert cos_f > cos_a - 0.03 and cos_f > 0.9, (cos_f, cos_a)
    for (n, b1), (_, b2) in zip(ms[0].named_buffers(), ms[2].named_buffers()):
        if b1.dtype.is_floating_point:
            assert torch.allclose(b1, b2, atol=3e-2, rtol=3e-2), n


def test_fp8_conv1x1_matches_fp32_within_e4m3_error():
    """e4m3 GEMM with delayed scaling vs an fp32 1x1 convolution: relative error of a few percent
    (3 mantissa bits), gradients exact up to bf16 (they do not go through fp8)."""
    from hefl_b200.ops import fp8
    torch.manual_seed(3)
    conv = fp8.Conv1x1(64, 128).cuda()
    x = _cl(torch.randn(8, 64, 16, 16, device="cuda").to(torch.bfloat16)).requires_grad_(True)
    ref = F.conv2d(x.detach().float(), conv.weight.detach().float())
    fp8.ENABLE = True
    try:
        for _ in range(3):                               # delayed scaling settles after the first call
            y = conv(x)
        assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
        rel = (y.float() - ref).norm() / ref.norm()
        assert rel < 0.06, rel
        g = _cl(torch.randn_like(y))
        y.backward(g)
        gx_ref = F.conv_transpose2d(g.float(), conv.weight.detach().float())
        assert (x.grad.float() - gx_ref).norm() / gx_ref.norm() < 2e-2
        gw_ref = torch.einsum("bohw,bihw->oi", g.float(), x.detach().float())
        assert (conv.weight.grad.view(128, 64) - gw_ref).norm() / gw_ref.norm() < 2e-2
        # strided variant (ResNet down-sampling path)
        down = fp8.Conv1x1(64, 128, stride=2).cuda()
        yd = down(x.detach())
        refd = F.conv2d(x.detach().float(), down.weight.detach().float(), stride=2)
        assert yd.shape == refd.shape and (yd.float() - refd).norm() / refd.norm() < 0.06
    finally:
        fp8.ENABLE = False


def test_resnet50_fp8_training_tracks_bf16():
    """ResNet-50 with e4m3 1x1 convolutions trains like the bf16 model from the same initialisation."""
    from hefl_b200.models import create_model
    from hefl_b200.ops import fp8
    x = torch.randn(16, 64, 64, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)).permute(0, 3, 1, 2)
    y = torch.randint(0, 4, (16,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))

    def run(use_fp8):
        torch.manual_seed(5)
        m = create_model("resnet50", num_classes=4).cuda().train()
        opt = torch.optim.SGD(m.parameters(), lr=0.002)
        fp8.ENABLE = use_fp8
        try:
            losses = []
            for _ in range(6):
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = F.cross_entropy(m(x).float(), y)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                losses.append(float(loss.detach()))
            used = sum(1 for mod in m.modules() if isinstance(mod, fp8.Conv1x1) and mod._fp8_state is not None)
            return losses, used
        finally:
            fp8.ENABLE = False

    l8, used8 = run(True)
    l16, used16 = run(False)
    assert used8 >= 30 and used16 == 0
    assert all(v == v and v < 1e4 for v in l8)
    # same trajectory as bf16, step by step (the loss itself need not fall on 6 steps of a random batch)
    for a8, a16 in zip(l8, l16):
        assert abs(a8 - a16) < 0.35 * max(a16, 0.1), (l8, l16)
