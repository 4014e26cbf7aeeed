"""Host logic of iswm_b200.optim that needs no GPU: CosineAnnealingLR step for step against torch's scheduler as
train.py:446-452 builds it (T_max = total_itrs, eta_min = lr * 0.01, stepped every iteration, train.py:1103), and the
state_dict round trip the checkpoint code relies on (train.py:570, :1016)."""
import types

import torch

from iswm_b200.optim import CosineAnnealingLR, setup_scheduler


class _Opt:
    def __init__(self, lr):
        self.param_groups = [{"lr": lr}]


def test_cosine_lr_matches_torch_step_for_step():
    for base_lr, T, eta in ((1e-3, 100, 1e-6), (1e-3, 30000, 1e-4 * 0.01), (0.01, 7, 0.0)):
        p = torch.nn.Parameter(torch.zeros(1))
        topt = torch.optim.SGD([p], lr=base_lr, momentum=0.9, nesterov=True)
        tsch = torch.optim.lr_scheduler.CosineAnnealingLR(topt, T_max=T, eta_min=eta)
        ours = CosineAnnealingLR(_Opt(base_lr), T_max=T, eta_min=eta)
        n = min(2 * T + 3, 2500)                     # past T_max too: torch keeps following the cosine, so do we
        for i in range(n):
            topt.step()
            tsch.step()
            ours.step()
            a, b = tsch.get_last_lr()[0], ours.get_last_lr()[0]
            assert abs(a - b) <= 1e-9 * base_lr + 1e-15, (T, i, a, b)   # torch uses the chained (recursive) form: fp64 noise only


def test_cosine_lr_state_dict_round_trip_and_torch_interchange():
    opts = types.SimpleNamespace(total_itrs=50, lr=1e-4)
    o1 = _Opt(1e-3)
    s1 = setup_scheduler(o1, opts)
    for _ in range(17):
        s1.step()
    sd = s1.state_dict()
    o2 = _Opt(1e-3)
    s2 = setup_scheduler(o2, opts)
    s2.load_state_dict(sd)
    assert s2.last_epoch == 17 and o2.param_groups[0]["lr"] == o1.param_groups[0]["lr"]
    for _ in range(5):
        s1.step(); s2.step()
    assert s1.get_last_lr() == s2.get_last_lr()
    # a checkpoint written by the reference's torch scheduler loads (same keys)
    p = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.SGD([p], lr=1e-3)
    tsch = torch.optim.lr_scheduler.CosineAnnealingLR(topt, T_max=50, eta_min=1e-6)
    for _ in range(9):
        topt.step(); tsch.step()
    o3 = _Opt(1e-3)
    s3 = CosineAnnealingLR(o3, T_max=1, eta_min=0.0)
    s3.load_state_dict(tsch.state_dict())
    assert s3.T_max == 50 and s3.last_epoch == 9 and abs(o3.param_groups[0]["lr"] - tsch.get_last_lr()[0]) < 1e-18
    tsch.step(); topt.step(); s3.step()
    assert abs(s3.get_last_lr()[0] - tsch.get_last_lr()[0]) <= 1e-12
