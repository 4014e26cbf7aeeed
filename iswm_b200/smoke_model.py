"""One tiny DeepLabV3+ invocation for `__graft_entry__.smoke()`: eval forward checked against the fp32 torch oracle
(the checker, oracle/torch_model.py), then one full train step (forward + weighted CE + backward + fused SGD)."""
from __future__ import annotations

import torch


def run(dev) -> None:
    from oracle import torch_model as TM                      # checker only
    from oracle.gen_golden import seeded_state_dict, synth_labels

    from .network import modeling
    from .optim import FusedSGD
    from .utils.loss import CrossEntropyLoss

    m = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False)
    sd = seeded_state_dict(m.state_dict(), 1234)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(7)
    x = torch.randn((2, 3, 64, 64), generator=g)
    y = synth_labels((2, 64, 64), seed=8, fg=0.2, ign=0.05)
    ref = TM.oracle_model("resnet50", 2, 16)
    ref.load_state_dict(sd)
    ref.eval()
    with torch.no_grad():
        want = ref(x)
    m.to(dev).eval()
    got = m(x.to(dev)).float().cpu()
    rel = float((got - want).norm() / want.norm())
    assert rel <= 2e-2, f"eval logits differ from the fp32 oracle: rel L2 {rel:.3e} (bar 2e-2, bf16 activations)"
    m.train()
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 2.5]), ignore_index=255).to(dev)
    opt = FusedSGD(m, lr=1e-3, momentum=0.9, weight_decay=1e-4)
    loss = crit(m(x.to(dev)), y.to(dev))
    opt.zero_grad()
    loss.backward()
    gn = float(m.engine().flat_g.norm())
    opt.step()
    torch.cuda.synchronize()
    assert torch.isfinite(loss).item() and gn > 0 and gn == gn, (float(loss), gn)
    from . import ops
    assert ops.abort_code() == 0, "a tensor-core kernel timed out on an mbarrier"
    print(f"model smoke ok: eval rel-L2 {rel:.2e}, train loss {float(loss):.4f}, |grad| {gn:.3e}")
