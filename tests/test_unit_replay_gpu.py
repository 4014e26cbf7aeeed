"""Unit-level replay: every conv+BN(+ReLU,+residual) unit of a real DeepLabV3+ train step is
re-computed from the ENGINE'S OWN tensors with stock torch ops in fp32 (forward: conv, batch
statistics, normalise; backward: BN gradient formula, weight gradient, data gradient) and compared
with what the CUDA kernels produced in place. Because each unit is checked on identical inputs,
this isolates kernel + wiring errors from the (large, chaotic) drift that bf16 rounding itself
causes across 50+ batch-normalised layers. Tolerances are one bf16 rounding of the stored result
(RMS 2^-9/sqrt(3) ~ 1.1e-3; we allow 4e-3 relative L2) and fp32 accumulation noise for fp32 outputs."""
import pytest
import torch
import torch.nn.functional as F

from iswm_b200.network import modeling
from iswm_b200.utils.loss import CrossEntropyLoss
from oracle.gen_golden import seeded_state_dict, synth_labels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


@pytest.mark.parametrize("backbone,os_,B,H,W", [("resnet50", 16, 4, 96, 80), ("resnet50", 8, 2, 64, 64)])
def test_every_unit_matches_torch_on_the_engines_own_tensors(backbone, os_, B, H, W):
    ctor = modeling.deeplabv3plus_resnet50 if backbone == "resnet50" else modeling.deeplabv3plus_resnet101
    m = ctor(num_classes=2, output_stride=os_, pretrained_backbone=False)
    m.load_state_dict(seeded_state_dict(m.state_dict(), 99))
    m.to(DEV).train()
    eng = m.engine()
    eng.dropout_p = 0.0
    eng.debug_units = []
    g = torch.Generator().manual_seed(3)
    x = torch.randn((B, 3, H, W), generator=g).to(DEV)
    y = synth_labels((B, H, W), seed=4, fg=0.2, ign=0.05).to(DEV)
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])).to(DEV)
    logits = m(x)
    loss = crit(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    recs = eng.debug_units
    n_units = len(eng.specs)
    assert len(recs) == n_units, (len(recs), n_units)
    bad = []

    def chk(name, what, got, ref, tol):
        e = rel(got, ref)
        if not (e <= tol):
            bad.append((name, what, round(e, 5)))

    for r in recs:
        name = r["name"]
        wq = r["w"].to(torch.bfloat16).float()
        if r.get("kind") == "cls":
            xin = nchw(r["x"])
            dlo = nchw(r["dlo"])[:, :2]
            # adjoint of the final bilinear upsample, via autograd on the stock op
            lo = torch.zeros((B, 2, xin.shape[2], xin.shape[3]), device=DEV, requires_grad=True)
            F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False).backward(r["dlogits"])
            chk(name, "dlo", dlo, lo.grad, 4e-3)
            chk(name, "dW", r["dW"], torch.nn.grad.conv2d_weight(xin, r["w"].shape, dlo), 2e-3)
            chk(name, "dbias", r["dbias"], r["dlogits"].sum((0, 2, 3)), 1e-3)
            chk(name, "dx", nchw(r["xgrad_after"]), torch.nn.grad.conv2d_input(xin.shape, wq, dlo), 4e-3)
            continue
        k, stride, dil = r["k"], r["stride"], r["dilation"]
        pad = dil * (k // 2)
        xin = r["image"].to(torch.bfloat16).float() if r["x"] is None else nchw(r["x"])
        raw, out, dout, dy = nchw(r["raw"]), nchw(r["out"]), nchw(r["dout"]), nchw(r["dy"])
        Mn = raw.numel() // raw.shape[1]
        # ---- forward
        y32 = F.conv2d(xin, wq, None, stride, pad, dil)
        chk(name, "conv", raw, y32, 4e-3)
        # batch statistics are those of the STORED bf16 conv output (the conv epilogue sums its staged tile)
        mean = raw.mean((0, 2, 3))
        var = raw.var((0, 2, 3), unbiased=False)
        chk(name, "mean", r["mean"], mean, 2e-3 if Mn > 8 else 2e-2)
        chk(name, "invstd", r["invstd"], torch.rsqrt(var + 1e-5), 2e-3 if Mn > 8 else 5e-2)
        if Mn > 8:   # ... and within bf16 rounding noise of the unrounded accumulator's statistics
            chk(name, "invstd(fp32 acc)", r["invstd"], torch.rsqrt(y32.var((0, 2, 3), unbiased=False) + 1e-5), 1e-2)
        mu, istd = r["mean"][None, :, None, None], r["invstd"][None, :, None, None]
        z = (raw - mu) * istd * r["gamma"][None, :, None, None] + r["beta"][None, :, None, None]
        if r["residual"] is not None:
            z = z + nchw(r["residual"])
        if r["relu"]:
            z = F.relu(z)
        chk(name, "bn_out", out, z, 4e-3)
        # ---- backward
        dz = dout * (out > 0) if r["relu"] else dout
        xhat = (raw - mu) * istd
        s1, s2 = dz.sum((0, 2, 3)), (dz * xhat).sum((0, 2, 3))
        dy_ref = (r["gamma"] * r["invstd"])[None, :, None, None] * (dz - (s1 / Mn)[None, :, None, None] - xhat * (s2 / Mn)[None, :, None, None])
        chk(name, "dy", dy, dy_ref, 4e-3 if Mn > 8 else 5e-2)
        scale = float(max(s1.norm(), s2.norm())) + 1e-12
        assert float((r["dbeta"] - s1).norm()) <= 2e-3 * scale + 1e-6, (name, "dbeta")
        assert float((r["dgamma"] - s2).norm()) <= 2e-3 * scale + 1e-6, (name, "dgamma")
        chk(name, "dW", r["dW"], torch.nn.grad.conv2d_weight(xin, r["w"].shape, dy, stride, pad, dil), 2e-3)
        if r["xgrad_after"] is not None:
            dx_ref = torch.nn.grad.conv2d_input(xin.shape, wq, dy, stride, pad, dil)
            before = 0 if r["xgrad_before"] is None else nchw(r["xgrad_before"])
            total = dx_ref + before
            chk(name, "dx(acc)", nchw(r["xgrad_after"]), total, 6e-3)
        if r["residual"] is not None:
            before = 0 if r["resgrad_before"] is None else nchw(r["resgrad_before"])
            chk(name, "dres", nchw(r["resgrad_after"]), dz + before, 4e-3)
    assert not bad, f"{len(bad)} mismatches, first: {bad[:15]}"
