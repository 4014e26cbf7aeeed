"""Generate tests/golden/*.npz from the REAL reference code (run in the build container).

    python oracle/gen_golden.py

Imports /root/reference through oracle/ref_import.py and records, for fixed seeds:
  loss_metric.npz   torch CrossEntropyLoss(weight, ignore_index=255) loss+grad (the criterion of
                    train.py:454-459), StreamMetrics._fast_hist and derived metrics
                    (metrics/stream_metrics.py:24-63), calculate_class_weights arithmetic
                    (train.py:401-410), argmax / softmax-threshold class maps;
  model_r50_os16.npz  reference deeplabv3plus_resnet50(num_classes=2, output_stride=16) in eval
                    and train mode on a 2x3x64x64 input with seeded weights: logits, loss, a few
                    gradients, BN running stats after one step; plus the seed recipe so the test
                    can rebuild the same weights without the reference;
  model_r50_os8.npz   deeplabv3plus_resnet50(output_stride=8) (rates 12/24/36, layers 3 and 4 dilated), eval only;
  model_r101_os8.npz  same for _load_model('deeplabv3plus','resnet101',2,output_stride=8), eval only.
Weights themselves are not stored (160 MB); they are regenerated from the seed through the
state_dict key order, which the fixture also pins (names + shapes).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def seeded_state_dict(ref_sd, seed):
    """Deterministic weights for a given key order/shape list (shared with the tests)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in ref_sd.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros_like(v)
        elif k.endswith("running_var"):
            sd[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(v.shape, generator=g) * 0.1
        elif v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            sd[k] = torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif k.endswith("weight"):          # BN gamma
            sd[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
            if k.endswith("bn3.weight"):
                # residual branches enter at 1/4 gain, as in a trained / zero-init-residual network
                # (resnet.py:165-172): with unit gain and batch-statistics BN the 16-33 stacked blocks
                # amplify ANY perturbation (fp32 summation order included) ~x1.25 per block, which
                # makes a bf16-vs-fp32 comparison meaningless on random weights
                sd[k] = sd[k] * 0.25
        else:                               # BN beta / conv bias
            sd[k] = torch.randn(v.shape, generator=g) * 0.1
    return sd


def synth_labels(shape, seed, fg=0.02, ign=0.01):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(shape, generator=g)
    y = (u < fg).long()
    y[torch.rand(shape, generator=g) < ign] = 255
    return y


def gen_loss_metric(ref_metrics):
    rng = np.random.default_rng(0)
    out = {}
    # KAT-2 style + random cases for the criterion
    for name, (B, C, H, W) in {"small": (2, 2, 5, 7), "c3": (2, 3, 4, 6), "mid": (3, 2, 32, 48)}.items():
        x = torch.tensor(rng.standard_normal((B, C, H, W)) * 3.0, dtype=torch.float32, requires_grad=True)
        y = torch.tensor(rng.integers(0, C, (B, H, W)), dtype=torch.long)
        y[torch.tensor(rng.random((B, H, W)) < 0.1)] = 255
        w = torch.tensor(rng.random(C) * 4 + 0.5, dtype=torch.float32)
        for wname, ww in (("w", w), ("nw", None)):
            crit = torch.nn.CrossEntropyLoss(weight=ww, ignore_index=255, reduction="mean")
            x.grad = None
            loss = crit(x, y)
            loss.backward()
            out[f"ce_{name}_{wname}_logits"] = x.detach().numpy()
            out[f"ce_{name}_{wname}_labels"] = y.numpy()
            out[f"ce_{name}_{wname}_weight"] = (ww if ww is not None else torch.ones(C)).numpy()
            out[f"ce_{name}_{wname}_loss"] = np.float64(loss.item())
            out[f"ce_{name}_{wname}_grad"] = x.grad.numpy().copy()
    # KAT-2 of SURVEY §8c
    x = torch.tensor([[[[2.0, -1.0], [0.5, 0.0]], [[0.0, 1.0], [0.5, 3.0]]]], requires_grad=True)
    y = torch.tensor([[[0, 1], [255, 1]]])
    loss = torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 3.0]), ignore_index=255)(x, y)
    loss.backward()
    out["kat2_loss"] = np.float64(loss.item())
    out["kat2_grad"] = x.grad.numpy().copy()
    # StreamMetrics._fast_hist + derived metrics, through the reference class
    with contextlib.redirect_stdout(io.StringIO()):
        sm = ref_metrics.StreamMetrics(2)
        gt = np.array([[0, 0, 1, 1], [1, 0, 255, 1]])
        pr = np.array([[0, 1, 1, 0], [1, 0, 1, 1]])
        out["kat1_hist"] = sm._fast_hist(gt.flatten(), pr.flatten())
        out["kat1_metrics"] = np.array(sm._calculate_foreground_metrics(out["kat1_hist"]), dtype=np.float64)
        for name, n, size in (("h2", 2, 5000), ("h5", 5, 4000)):
            t = rng.integers(0, n, size)
            t[rng.random(size) < 0.05] = 255
            p = rng.integers(0, n, size)
            smn = ref_metrics.StreamMetrics(n)
            out[f"{name}_true"] = t
            out[f"{name}_pred"] = p
            out[f"{name}_hist"] = smn._fast_hist(t, p)
        # update()/get_results() accumulation path, sequence_data=False and True
        sm2 = ref_metrics.StreamMetrics(2)
        t = rng.integers(0, 2, (3, 16, 16)); p = rng.integers(0, 2, (3, 16, 16))
        sm2.update(t, p, sequence_data=True)
        sm2.update(t[0], p[0], sequence_data=False)
        res = sm2.get_results()
        out["upd_true"] = t; out["upd_pred"] = p
        out["upd_cm"] = sm2.confusion_matrix.copy()
        out["upd_vals"] = np.array([res["MIoU"], res["Foreground IoU"], res["Foreground F1"], res["Precision"], res["Recall"]], dtype=np.float64)
    # class weights (train.py:401-410 arithmetic)
    out["kat3_w"] = torch.FloatTensor([1.0, np.sqrt(1000 / 37)]).numpy()
    # argmax / threshold
    lg = torch.tensor(rng.standard_normal((2, 2, 9, 11)), dtype=torch.float32)
    lg[0, :, 0, 0] = 0.0  # tie
    out["am_logits"] = lg.numpy()
    out["am_argmax"] = lg.max(1)[1].numpy()
    prob = torch.softmax(lg, dim=1)
    out["am_thresh"] = (prob[:, 1] > 0.5).long().numpy()
    out["am_conf"] = (prob[:, 1].numpy() * 255).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "loss_metric.npz"), **out)
    print("wrote loss_metric.npz", len(out), "arrays")


def gen_model(modeling, name, builder, H, W, do_train):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = builder()
    ref_sd = model.state_dict()
    sd = seeded_state_dict(ref_sd, seed=1234)
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(7)
    x = torch.randn((2, 3, H, W), generator=g)
    y = synth_labels((2, H, W), seed=8, fg=0.2, ign=0.05)
    w = torch.tensor([1.0, 2.5])
    out = {"keys": np.array(list(ref_sd.keys())), "shapes": np.array([str(tuple(v.shape)) for v in ref_sd.values()]),
           "n_params": np.int64(sum(p.numel() for p in model.parameters())), "x": x.numpy(), "y": y.numpy(), "w": w.numpy()}
    model.eval()
    with torch.no_grad():
        out["eval_logits"] = model(x).numpy()
    if do_train:
        model.train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0                      # parity runs: no dropout noise
        crit = torch.nn.CrossEntropyLoss(weight=w, ignore_index=255, reduction="mean")
        logits = model(x)
        loss = crit(logits, y)
        loss.backward()
        out["train_logits"] = logits.detach().numpy()
        out["train_loss"] = np.float64(loss.item())
        named = dict(model.named_parameters())
        for k in ("backbone.conv1.weight", "backbone.layer1.0.conv2.weight", "backbone.layer2.0.downsample.0.weight",
                  "backbone.layer4.2.conv2.weight", "backbone.layer4.2.bn3.weight", "backbone.layer4.2.bn3.bias",
                  "classifier.aspp.convs.1.0.weight", "classifier.aspp.convs.4.1.weight", "classifier.aspp.project.0.weight",
                  "classifier.project.0.weight", "classifier.classifier.0.weight", "classifier.classifier.3.weight",
                  "classifier.classifier.6.weight", "classifier.classifier.6.bias", "backbone.bn1.weight"):
            gk = named[k].grad
            out["grad:" + k + ":norm"] = np.float64(gk.norm().item())
            out["grad:" + k + ":head"] = gk.flatten()[:64].numpy().copy()
        out["bn1_running_mean_after"] = model.state_dict()["backbone.bn1.running_mean"].numpy().copy()
        out["bn1_running_var_after"] = model.state_dict()["backbone.bn1.running_var"].numpy().copy()
        out["aspp_pool_bn_running_var_after"] = model.state_dict()["classifier.aspp.convs.4.2.running_var"].numpy().copy()
    np.savez_compressed(os.path.join(OUT, name), **out)
    print("wrote", name)


def main():
    os.makedirs(OUT, exist_ok=True)
    modeling, ref_metrics = ref_import.reference_modules()
    gen_loss_metric(ref_metrics)
    gen_model(modeling, "model_r50_os16.npz",
              lambda: modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False), 96, 96, True)
    gen_model(modeling, "model_r50_os8.npz",
              lambda: modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=8, pretrained_backbone=False), 56, 72, False)
    gen_model(modeling, "model_r101_os8.npz",
              lambda: modeling._load_model("deeplabv3plus", "resnet101", 2, output_stride=8, pretrained_backbone=False), 48, 40, False)


if __name__ == "__main__":
    main()
