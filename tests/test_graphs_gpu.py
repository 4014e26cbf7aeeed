"""CUDA-graph train step (iswm_b200.graphs) and the device-resident step state it relies on (dropout step counter,
learning rate / Adam step count read by the optimiser kernels from device memory)."""
import numpy as np
import pytest
import torch

from iswm_b200 import _lib
from iswm_b200.graphs import GraphedTrainStep
from iswm_b200.network import modeling
from iswm_b200.optim import CosineAnnealingLR, FusedAdamW, FusedSGD
from iswm_b200.utils.loss import CrossEntropyLoss

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batches(n, B=2, S=96):
    g = torch.Generator().manual_seed(3)
    out = []
    for _ in range(n):
        x = torch.randn((B, 3, S, S), generator=g)
        y = (torch.rand((B, S, S), generator=g) < 0.1).long()
        y[torch.rand((B, S, S), generator=g) < 0.02] = 255
        out.append((x.to(DEV), y.to(DEV)))
    return out


def _model():
    torch.manual_seed(0)
    return modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False).to(DEV).train()


def _state(m, opt):
    eng = m.engine()
    st = {"w": eng.flat_w, "buffers": list(m.buffers())}
    for k in ("_mom", "_m", "_v"):
        if getattr(opt, k, None) is not None:
            st[k] = getattr(opt, k)
    return st


@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_graphed_step_equals_eager(opt_name):
    """Every replay against the eager step FROM THE SAME STATE (the eager model's state is copied into the graphed one
    before each step, so that the chaotic amplification of the ~1e-7 weight-gradient atomics noise over several steps -
    BatchNorm over the 2-sample pooled ASPP branch flips signs - stays out of the comparison): Dropout ON, a new batch
    every step, cosine LR stepped every iteration. The forward is bit-reproducible, so the losses must agree to fp32
    noise and the updated weights to the atomics noise."""
    batches = _batches(4)
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])).to(DEV)

    def make():
        m = _model()
        opt = FusedSGD(m, lr=1e-2, momentum=0.9, weight_decay=1e-4) if opt_name == "sgd" else FusedAdamW(m, lr=1e-3, weight_decay=1e-2)
        return m, opt, CosineAnnealingLR(opt, T_max=8, eta_min=1e-4)

    me, oe, se = make()
    mg, og, sg = make()
    stepper = GraphedTrainStep(mg, crit, og, fused_tail=False)     # the same kernels as the eager step (the fused tail: tests/test_tail_gpu.py)
    for i, (x, y) in enumerate(batches):
        if i > 0:                                   # same starting state for this step
            a, b = _state(me, oe), _state(mg, og)
            assert a.keys() == b.keys()
            with torch.no_grad():
                for k in a:
                    if k == "buffers":
                        for u, v in zip(a[k], b[k]):
                            v.copy_(u)
                    else:
                        b[k].copy_(a[k])
            mg.engine().invalidate_packed()
        loss = crit(me(x), y)
        oe.zero_grad()
        loss.backward()
        oe.step()
        lg = float(stepper(x, y))
        le = float(loss.detach())
        assert abs(lg - le) <= 1e-6 * max(1.0, abs(le)), (i, le, lg)
        we, wg = me.engine().flat_w, mg.engine().flat_w
        assert float((wg - we).norm() / we.norm()) <= 1e-6, i
        for u, v in zip(me.buffers(), mg.buffers()):      # BatchNorm running statistics / num_batches_tracked
            assert torch.allclose(u.float(), v.float(), rtol=1e-6, atol=1e-7)
        assert abs(oe.param_groups[0]["lr"] - og.param_groups[0]["lr"]) == 0.0
        se.step()
        sg.step()
    assert oe._steps == og._steps == 4 and me.engine().step == mg.engine().step == 4


def test_graphed_step_then_eval_uses_fresh_weights():
    m = _model()
    crit = CrossEntropyLoss().to(DEV)
    opt = FusedSGD(m, lr=5e-2, momentum=0.9)
    stepper = GraphedTrainStep(m, crit, opt)
    (x, y), = _batches(1)
    m.eval()
    with pytest.raises(RuntimeError):
        stepper(x, y)
    m.train()
    for _ in range(2):
        stepper(x, y)
    m.eval()
    with torch.no_grad():
        a = m(x).clone()
    m.engine().invalidate_packed()                          # force a repack: must change nothing if the cache was fresh
    with torch.no_grad():
        b = m(x)
    assert torch.equal(a, b)
    assert stepper.launches_per_replay > 300
    with pytest.raises(ValueError):
        stepper(x[:1], y[:1])
    # integer labels of another width (the reference's loader yields uint8; train.py:1040 widens them) are widened by the copy into the
    # captured step's buffer; floating-point "labels" are refused
    stepper(x, y.to(torch.uint8))
    assert stepper.labels.dtype == y.dtype and torch.equal(stepper.labels, y)
    with pytest.raises(ValueError):
        stepper(x, y.float())


def test_dropout_mask_follows_device_step_counter():
    """bn_train_apply with Dropout: seed_eff = seed + 1000003 * *d_step - equal to passing that seed by value, a
    different mask for another step, keep fraction ~ 1 - p, survivors scaled by 1/(1-p)."""
    M, C, p = 4096, 64, 0.25
    g = torch.Generator().manual_seed(1)
    x = torch.randn((M, C), generator=g).to(torch.bfloat16).to(DEV)
    stats = torch.stack([x.double().sum(0), (x.double() ** 2).sum(0)]).reshape(-1).contiguous()
    gm, bt = torch.ones(C, device=DEV), torch.full((C,), 3.0, device=DEV)     # shift keeps every output away from 0
    save = torch.empty(2 * C, device=DEV)
    st = torch.cuda.current_stream().cuda_stream

    def run(seed, step):
        out = torch.empty_like(x)
        sp = None if step is None else torch.tensor([step], dtype=torch.int64, device=DEV)
        _lib.check(_lib.lib().iswm_bn_train_apply(x.data_ptr(), C, stats.data_ptr(), 1, M, C, gm.data_ptr(), bt.data_ptr(), 1e-5, 0.1, None, None, None,
                                                  save.data_ptr(), save[C:].data_ptr(), None, C, 0, p, seed, None if sp is None else sp.data_ptr(),
                                                  out.data_ptr(), C, None, st), "bn_train_apply")
        torch.cuda.synchronize()
        return out.float()

    base = run(77, None)
    assert torch.equal(run(77, 0), base)
    s5 = run(77, 5)
    assert torch.equal(s5, run(77 + 5 * 1000003, None)) and not torch.equal(s5, base)
    keep = (s5 != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.01
    nodrop = run(77, None) * 0 + (x.float() - x.float().mean(0)) / x.float().var(0, unbiased=False).add(1e-5).sqrt() + 3.0
    kept = s5 != 0
    assert torch.allclose(s5[kept], (nodrop / (1 - p))[kept], rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("mode", ["eager_fused", "graphed", "torch_optim"])
def test_eval_after_training_refolds_batchnorm(mode):
    """validate_and_save / predict between training iterations (train.py:620-745): eval -> train 2 steps -> eval. The
    second eval must see the NEW running statistics and affine parameters, i.e. equal a fresh model loaded from the same
    state_dict. (The folded eval-mode scale / shift are cached per unit; the kernels update running_mean / var through raw
    pointers and the fused optimisers update gamma / beta through the flat master buffer, so no tensor version moves.)"""
    batches = _batches(2)
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])).to(DEV)
    m = _model()
    xe = batches[0][0]
    m.eval()
    with torch.no_grad():
        before = m(xe).clone()
    m.train()
    if mode == "torch_optim":
        opt = torch.optim.SGD(m.parameters(), lr=1e-2, momentum=0.9)
    else:
        opt = FusedSGD(m, lr=1e-2, momentum=0.9)
    stepper = GraphedTrainStep(m, crit, opt) if mode == "graphed" else None
    for x, y in batches:
        if stepper is not None:
            stepper(x, y)
        else:
            loss = crit(m(x), y)
            opt.zero_grad()
            loss.backward()
            opt.step()
    m.eval()
    with torch.no_grad():
        after = m(xe).clone()
    fresh = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False)
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in m.state_dict().items()})
    fresh.to(DEV).eval()
    with torch.no_grad():
        want = fresh(xe)
    assert torch.equal(after, want), float((after - want).abs().max())
    assert not torch.equal(after, before)
    assert int(m.state_dict()["backbone.bn1.num_batches_tracked"]) == 2


def test_eval_forward_between_forward_and_backward_keeps_the_tape():
    """An eval / predict forward issued between a train forward and its loss.backward() (periodic visualisation hooks do
    that) must not destroy the pending tape; a second TRAIN forward supersedes the first one and its stale backward raises."""
    (x, y), (x2, _) = _batches(2)
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])).to(DEV)
    ma, mb = _model(), _model()
    for m in (ma, mb):
        m.engine().dropout_p = 0.0
    la = crit(ma(x), y)
    la.backward()
    lb = crit(mb(x), y)
    mb.eval()
    with torch.no_grad():
        mb(x2)
        mb.forward_lowres(x2)
    mb.train()
    lb.backward()
    torch.cuda.synchronize()
    assert float(la.detach()) == float(lb.detach())
    ga, gb = ma.engine().flat_g, mb.engine().flat_g
    assert float((ga - gb).norm() / ga.norm()) <= 1e-5
    stale = crit(mb(x), y)
    crit(mb(x2), y)
    with pytest.raises(RuntimeError, match="superseded"):
        stale.backward()


def test_fused_sgd_repack_is_bit_identical_to_step_then_pack():
    """FusedSGD with iswm_sgd_pack_batched (update + both bf16 operand packings in one pass, csrc/sgd_pack.cu) against
    iswm_sgd_step followed by iswm_pack_weights_batched: master weights, momentum and every packed operand bit-identical after
    three steps (weight decay, nesterov, the ASPP K-concatenated dgrad operand, the stem's row-tap operand, Cout = 2 / 48 tails)."""
    batches = _batches(3)
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])).to(DEV)
    res = []
    for fuse in (True, False):
        m = _model()
        m.engine().dropout_p = 0.0
        opt = FusedSGD(m, lr=5e-2, momentum=0.9, weight_decay=1e-4, nesterov=True)
        opt.fuse_pack = fuse
        n_pack = 0
        for x, y in batches:
            loss = crit(m(x), y)
            opt.zero_grad()
            loss.backward()
            n0 = _lib.launch_count()
            opt.step()
            n_pack += _lib.launch_count() - n0
        eng = m.engine()
        if not fuse:
            eng.pack_all(True)
        assert eng.packed_is_fresh()
        torch.cuda.synchronize()
        packed = {s.name: (s.packed_fwd.clone(), None if s.packed_dgrad is None else s.packed_dgrad.clone()) for s in eng.specs}
        res.append((eng.flat_w.clone(), opt._mom.clone(), packed, eng.aspp_wcat.clone(), float(loss.detach()), n_pack))
    a, b = res
    assert a[5] == 3 and b[5] == 3                              # one launch per step either way (the repack is the unfused path's extra one)
    # (three steps at this learning rate amplify the weight-gradient kernels' fp32 atomics order chaotically - DESIGN 2 - so the two
    # RUNS are not compared with each other; the update arithmetic is pinned by the test below, the packings here)
    # the fused path's operands equal a fresh pack of ITS OWN weights, bit for bit
    m2 = _model()
    e2 = m2.engine()
    crit(m2(batches[0][0]), batches[0][1]).backward()            # engine bound to the device, operand buffers allocated
    with torch.no_grad():
        e2.flatten_parameters().copy_(a[0])
    e2.invalidate_packed()
    e2.pack_all(True)
    torch.cuda.synchronize()
    for s in e2.specs:
        f, d = a[2][s.name]
        assert torch.equal(s.packed_fwd, f), s.name
        if d is not None and s.packed_dgrad is not None:
            assert torch.equal(s.packed_dgrad, d), s.name
    assert torch.equal(e2.aspp_wcat, a[3])


def test_fused_sgd_update_matches_sgd_step_kernel_exactly():
    """Same gradients in, same weights / momentum out: the fused kernel's update arithmetic is iswm_sgd_step's."""
    m = _model()
    crit = CrossEntropyLoss().to(DEV)
    (x, y), = _batches(1)
    opt = FusedSGD(m, lr=5e-2, momentum=0.9, weight_decay=1e-4, nesterov=True)
    eng = m.engine()
    eng.flatten_parameters()                                     # before the first pack: the very first step is fused already
    loss = crit(m(x), y)
    opt.zero_grad()
    loss.backward()
    w0, g0 = eng.flat_w.clone(), eng.flat_g.clone()
    n0 = _lib.launch_count()
    opt.step()                                                   # fused (first step: momentum buffer = gradient)
    assert _lib.launch_count() - n0 == 1 and eng.packed_is_fresh()
    loss = crit(m(x), y)
    opt.zero_grad()
    loss.backward()
    w1, g1, m1 = eng.flat_w.clone(), eng.flat_g.clone(), opt._mom.clone()
    opt.step()                                                   # fused, momentum in play
    torch.cuda.synchronize()
    # replay both updates with the plain kernel on copies
    L, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    wa, ma = w0.clone(), torch.zeros_like(w0)
    _lib.check(L.iswm_sgd_step(wa.data_ptr(), g0.data_ptr(), ma.data_ptr(), wa.numel(), 5e-2, 0.9, 1e-4, 1, 1, None, st))
    assert torch.equal(wa, w1) and torch.equal(ma, m1)
    _lib.check(L.iswm_sgd_step(wa.data_ptr(), g1.data_ptr(), ma.data_ptr(), wa.numel(), 5e-2, 0.9, 1e-4, 1, 0, None, st))
    torch.cuda.synchronize()
    assert torch.equal(wa, eng.flat_w) and torch.equal(ma, opt._mom)


def test_graph_replay_repacks_after_weights_were_loaded_from_outside():
    """The captured step holds no repack at its start when FusedSGD repacks inside its own kernel; weights written from outside
    between replays (load_state_dict) must still reach the bf16 operands before the next replay."""
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])).to(DEV)
    (x, y), (x2, y2) = _batches(2)
    mg = _model()
    mg.engine().dropout_p = 0.0
    og = FusedSGD(mg, lr=1e-2, momentum=0.0, nesterov=False)
    stepper = GraphedTrainStep(mg, crit, og)
    stepper(x, y)
    assert stepper._fused_pack
    torch.manual_seed(123)
    donor = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False).to(DEV).train()
    sd = {k: v.clone() for k, v in donor.state_dict().items()}
    mg.load_state_dict(sd)
    lg = float(stepper(x2, y2))
    me = _model()
    me.engine().dropout_p = 0.0
    me.load_state_dict(sd)
    le = float(crit(me(x2), y2).detach())
    assert abs(lg - le) <= 1e-6 * max(1.0, abs(le)), (lg, le)
