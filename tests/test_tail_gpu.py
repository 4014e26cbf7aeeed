"""Fused train tail (csrc/tail_fused.cu: final x4 upsample + weighted CE + their backward without full-resolution tensors,
SURVEY kernels K11 + K12 + K13) against the unfused kernel chain, against torch autograd, and inside a train step."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from iswm_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _st():
    return torch.cuda.current_stream().cuda_stream


def _unfused(lo, labels, weight, ignore=255, ldp=8):
    """iswm_logits_up_fwd -> iswm_class_hist -> iswm_wce_fwd_bwd -> iswm_logits_up_bwd (what the engine runs without the fusion)."""
    L = _lib.lib()
    B, h, w, C = lo.shape
    H, W = 4 * h, 4 * w
    logits = torch.empty((B, C, H, W), dtype=torch.float32, device=DEV)
    _lib.check(L.iswm_logits_up_fwd(lo.data_ptr(), B, h, w, C, H, W, logits.data_ptr(), _st()))
    hist = ops.class_hist(labels, C)
    loss, grad = ops.wce_fwd_bwd(logits, labels, weight, hist, ignore, 1.0, True)
    dlo = torch.empty((B, h, w, ldp), dtype=torch.bfloat16, device=DEV)
    bias = torch.zeros(C, dtype=torch.float32, device=DEV)
    _lib.check(L.iswm_logits_up_bwd(grad.data_ptr(), B, h, w, C, H, W, dlo.data_ptr(), ldp, bias.data_ptr(), _st()))
    return loss, hist, dlo, bias, logits


def _fused(lo, labels, weight, ignore=255, ldp=8, g=None):
    dlo_acc, hist, num = ops.tail_fwd(lo, labels, weight, ignore)
    loss = ops.tail_loss(num, weight, hist, ignore)
    B, h, w, _ = lo.shape
    dlo = torch.full((B, h, w, ldp), float("nan"), dtype=torch.bfloat16, device=DEV)     # every channel must be written
    bias = torch.zeros(2, dtype=torch.float32, device=DEV)
    scratch = torch.zeros(8200, dtype=torch.uint8, device=DEV)
    ops.tail_bwd(dlo_acc, weight, hist, ignore, g, dlo, bias, scratch)
    return loss, hist, dlo, bias, dlo_acc, scratch


def _case(B, h, w, dtype, seed, p_fg=0.25, p_ign=0.04):
    g = torch.Generator().manual_seed(seed)
    lo = (torch.randn((B, h, w, 2), generator=g) * 2.0).to(DEV)
    y = (torch.rand((B, 4 * h, 4 * w), generator=g) < p_fg).long()
    y[torch.rand((B, 4 * h, 4 * w), generator=g) < p_ign] = 255
    return lo, y.to(dtype).to(DEV)


@pytest.mark.parametrize("B,h,w,dtype,weighted", [(2, 32, 32, torch.int64, True), (3, 12, 20, torch.uint8, True), (1, 5, 17, torch.int32, False),
                                                  (2, 8, 8, torch.int64, True), (4, 128, 128, torch.int64, True)])
def test_fused_tail_equals_the_unfused_chain(B, h, w, dtype, weighted):
    lo, y = _case(B, h, w, dtype, B * h + w)
    wt = torch.tensor([1.0, 6.5], device=DEV) if weighted else None
    l0, h0, d0, b0, _ = _unfused(lo, y, wt)
    l1, h1, d1, b1, _, _ = _fused(lo, y, wt)
    assert torch.equal(h0, h1)                                         # integer counts: exact
    assert abs(float(l1) - float(l0)) <= 2e-6 * abs(float(l0)), (float(l0), float(l1))
    a, b = d0.float(), d1.float()
    assert torch.equal(a[..., 2:], torch.zeros_like(a[..., 2:])) and torch.equal(b[..., 2:], torch.zeros_like(b[..., 2:]))
    # the normaliser is applied after the adjoint instead of before it (and the two adjoints add their <= 64 terms in different
    # orders): fp32 rounding of the LARGEST term, visible as rare 1-ulp bf16 flips - more where positive and negative terms cancel
    diff = (a[..., :2] - b[..., :2]).abs()
    ulp = a[..., :2].abs().clamp_min(1e-30) * 2.0 ** -7 + 1e-6 * float(a.abs().max())
    assert bool((diff <= ulp).all()), float((diff / ulp).max())
    assert float((diff > 0).float().mean()) < 0.02, float((diff > 0).float().mean())
    assert torch.allclose(b0, b1, rtol=2e-4, atol=1e-6), (b0, b1)
    print(f"fused tail {B}x{h}x{w}: loss rel {abs(float(l1) - float(l0)) / abs(float(l0)):.2e}, bf16 flips {float((diff > 0).float().mean()):.2e}")


def test_fused_tail_equals_torch_autograd_and_is_reproducible():
    lo, y = _case(2, 16, 24, torch.int64, 7)
    wt = torch.tensor([1.0, 3.0], device=DEV)
    t = lo.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    up = F.interpolate(t, size=(64, 96), mode="bilinear", align_corners=False)
    loss = F.cross_entropy(up, y, weight=wt, ignore_index=255)
    loss.backward()
    gscale = torch.tensor(0.5, device=DEV)
    l1, h1, d1, b1, acc, _ = _fused(lo, y, wt, g=gscale)
    assert abs(float(l1) - float(loss)) <= 2e-6 * abs(float(loss))
    D = float((wt.double() * h1.double()).sum())
    ref = t.grad.permute(0, 2, 3, 1)
    assert torch.allclose(acc / D, ref, rtol=1e-4, atol=1e-9)
    assert torch.allclose(d1[..., :2].float(), 0.5 * ref, rtol=1e-2, atol=1e-9)          # bf16 operand, upstream gradient 0.5
    assert torch.allclose(b1, 0.5 * t.grad.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-7)
    assert torch.equal(acc[..., 1], -acc[..., 0])
    for _ in range(3):                                                   # no atomics on the gradient: bit-reproducible
        l2, h2, d2, b2, acc2, _ = _fused(lo, y, wt, g=gscale)
        assert torch.equal(acc2, acc) and torch.equal(d2, d1) and torch.equal(b2, b1) and torch.equal(h2, h1)


def test_fused_tail_all_ignored_batch_and_scratch_reuse():
    lo, y = _case(1, 8, 8, torch.int64, 3)
    y.fill_(255)
    l1, h1, d1, b1, acc, scratch = _fused(lo, y, None)
    assert torch.isnan(l1).item() and h1.tolist() == [0, 0]             # torch: nan for an all-ignored batch
    assert torch.count_nonzero(d1.float()).item() == 0 and torch.count_nonzero(b1).item() == 0
    assert int(scratch.view(torch.int32)[2048]) == 0                     # the block counter re-armed itself
    with pytest.raises(ValueError):
        ops.tail_fwd(lo, y[:, :16], None)                               # labels must be the x4 grid
    with pytest.raises(ValueError):
        ops.tail_fwd(lo, y[:, :16, :16].contiguous(), None)


def _model():
    from iswm_b200.network import modeling
    torch.manual_seed(0)
    return modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False).to(DEV).train()


def test_forward_loss_equals_model_then_criterion():
    """model.forward_loss (fused tail) against criterion(model(x), y) from the same weights: same loss, same gradients up to the
    bf16 flips of the classifier-output gradient; the fallback conditions give the unfused result exactly."""
    from iswm_b200.utils.loss import CrossEntropyLoss
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, 3, 96, 96), generator=g).to(DEV)
    y = (torch.rand((2, 96, 96), generator=g) < 0.2).long()
    y[torch.rand((2, 96, 96), generator=g) < 0.03] = 255
    y = y.to(DEV)
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 4.0])).to(DEV)
    ma, mb = _model(), _model()
    for m in (ma, mb):
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
    la = crit(ma(x), y)
    la.backward()
    lb = mb.forward_loss(x, y, crit)
    lb.backward()
    assert abs(float(la) - float(lb)) <= 2e-6 * abs(float(la))
    ga, gb = ma.engine().flat_g, mb.engine().flat_g
    rel = float((ga - gb).norm() / ga.norm())
    print(f"forward_loss vs unfused: loss {float(la):.6f} / {float(lb):.6f}, rel grad diff {rel:.2e}")
    assert rel <= 2e-3, rel
    for u, v in zip(ma.buffers(), mb.buffers()):                         # BatchNorm statistics: the forward is the same
        assert torch.equal(u, v)
    # fallbacks: eval mode and odd sizes go through the plain composition
    mb.eval()
    with torch.no_grad():
        assert float(mb.forward_loss(x, y, crit)) == float(crit(mb(x), y))
    mb.train()
    xo, yo = x[:, :, :94, :94].contiguous(), y[:, :94, :94].contiguous()
    lo_ = mb.forward_loss(xo, yo, crit)
    lo_.backward()
    assert torch.isfinite(lo_).item()


def test_graphed_step_with_the_fused_tail_tracks_the_unfused_step():
    from iswm_b200.graphs import GraphedTrainStep
    from iswm_b200.optim import FusedSGD
    from iswm_b200.utils.loss import CrossEntropyLoss
    g = torch.Generator().manual_seed(11)
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])).to(DEV)
    ma, mb = _model(), _model()
    oa, ob = FusedSGD(ma, lr=1e-2, momentum=0.9, weight_decay=1e-4), FusedSGD(mb, lr=1e-2, momentum=0.9, weight_decay=1e-4)
    sa, sb = GraphedTrainStep(ma, crit, oa, fused_tail=False), GraphedTrainStep(mb, crit, ob, fused_tail=True)
    assert sb.fused_tail and not sa.fused_tail
    for i in range(3):
        x = torch.randn((2, 3, 96, 96), generator=g).to(DEV)
        y = (torch.rand((2, 96, 96), generator=g) < 0.2).long().to(DEV)
        if i > 0:
            with torch.no_grad():
                mb.engine().flat_w.copy_(ma.engine().flat_w)
                for u, v in zip(ma.buffers(), mb.buffers()):
                    v.copy_(u)
                ob._mom.copy_(oa._mom)
            mb.engine().invalidate_packed()
        la, lb = float(sa(x, y)), float(sb(x, y))
        assert abs(la - lb) <= 2e-6 * max(1.0, abs(la)), (i, la, lb)
        wa, wb = ma.engine().flat_w, mb.engine().flat_w
        assert float((wa - wb).norm() / wa.norm()) <= 1e-5, i
    assert sb.launches_per_replay < sa.launches_per_replay               # 3 tail launches instead of 6
