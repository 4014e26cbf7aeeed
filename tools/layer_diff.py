"""Debug tool (GPU box): per-unit forward error and per-parameter gradient error of the CUDA engine
against the fp32 torch oracle on the host CPU.   python tools/layer_diff.py [train|eval] [backbone] [os] [H] [W]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iswm_b200.network import modeling  # noqa: E402
from iswm_b200.utils.loss import CrossEntropyLoss  # noqa: E402
from oracle import torch_model as TM  # noqa: E402
from oracle.gen_golden import seeded_state_dict, synth_labels  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "train"
    matched = mode.endswith("q")          # trainq / evalq: reference = precision-matched oracle
    mode = mode.rstrip("q")
    backbone = sys.argv[2] if len(sys.argv) > 2 else "resnet50"
    os_ = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 96
    W = int(sys.argv[5]) if len(sys.argv) > 5 else 96
    B = int(sys.argv[6]) if len(sys.argv) > 6 else 2
    ctor = modeling.deeplabv3plus_resnet50 if backbone == "resnet50" else modeling.deeplabv3plus_resnet101
    m = ctor(num_classes=2, output_stride=os_, pretrained_backbone=False)
    sd = seeded_state_dict(m.state_dict(), 1234)
    m.load_state_dict(sd)
    o = TM.oracle_model(backbone, 2, os_)
    o.load_state_dict(sd)
    for mod in o.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    g = torch.Generator().manual_seed(7)
    x = torch.randn((B, 3, H, W), generator=g)
    y = synth_labels((B, H, W), seed=8, fg=0.2, ign=0.05)
    w = torch.tensor([1.0, 2.5])
    caps = {}

    def hook(name):
        def f(mod, inp, out):
            if isinstance(out, torch.Tensor):
                caps[name] = out.detach()
        return f
    for name, mod in o.named_modules():
        mod.register_forward_hook(hook(name))
    if matched:
        from oracle import torch_model_q as TQ
        TQ.TRACE = []
        if mode == "train":
            o.train()
            ref_logits, ref_loss = TQ.train_step_q(o, x, y, w)
        else:
            o.eval()
            with torch.no_grad():
                ref_logits = TQ.forward_q(o, x, False)
        caps.clear()
    elif mode == "train":
        o.train()
        ref_logits, ref_loss = TM.train_step(o, x, y, w)
    else:
        o.eval()
        with torch.no_grad():
            ref_logits = o(x)
    m.to("cuda:0")
    m.train(mode == "train")
    eng = m.engine()
    eng.dropout_p = 0.0
    eng.debug_taps = {}
    if mode == "train":
        crit = CrossEntropyLoss(weight=w).cuda()
        logits = m(x.cuda())
        loss = crit(logits, y.cuda())
        loss.backward()
        print("loss", loss.item(), "ref", ref_loss.item())
    else:
        logits = m(x.cuda())
    torch.cuda.synchronize()

    def err(a, b):
        a, b = a.float().cpu(), b.float().cpu()
        return float((a - b).norm() / (b.norm() + 1e-20)), float((a - b).abs().max() / (b.abs().max() + 1e-20))

    def ref_for(unit):
        # unit name = conv module name; post-BN(-ReLU)(-residual) reference
        if unit == "backbone.conv1":
            return torch.relu(caps["backbone.bn1"])
        parts = unit.split(".")
        if unit.startswith("backbone.layer"):
            blk = ".".join(parts[:3])
            if parts[3] == "conv3":
                return caps[blk]
            if parts[3] == "downsample":
                return caps[blk + ".downsample.1"]
            return torch.relu(caps[blk + ".bn" + parts[3][-1]])
        if unit == "classifier.project.0":
            return caps["classifier.project.2"] if "classifier.project.2" in caps else torch.relu(caps["classifier.project.1"])
        if unit.startswith("classifier.aspp.convs.4"):
            return torch.relu(caps["classifier.aspp.convs.4.2"])
        if unit.startswith("classifier.aspp.convs."):
            i = parts[3]
            return torch.relu(caps[f"classifier.aspp.convs.{i}.1"])
        if unit == "classifier.aspp.project.0":
            return torch.relu(caps["classifier.aspp.project.1"])
        if unit == "classifier.classifier.0":
            return torch.relu(caps["classifier.classifier.1"])
        if unit == "classifier.classifier.3":
            return torch.relu(caps["classifier.classifier.4"])
        return None

    print(f"{'unit':45s} {'relL2':>9s} {'max/rng':>9s}")
    if matched:
        from oracle import torch_model_q as TQ
        names = [n for n in eng.debug_taps if not n.endswith(":raw")]
        assert len(names) == len(TQ.TRACE), (len(names), len(TQ.TRACE))
        for n, r in zip(names, TQ.TRACE):
            t = eng.debug_taps[n]
            e = err(t, r)
            nz = float((t.float().cpu() != r).float().mean())
            print(f"{n:45s} {e[0]:9.5f} {e[1]:9.5f}  frac_diff={nz:.4f}")
    for name, t in ({} if matched else eng.debug_taps).items():
        if name.endswith(":raw"):
            base = name[:-4]
            refname = None
            if base.startswith("backbone.layer") or base.startswith("classifier") or base == "backbone.conv1":
                refname = base
            r = caps.get(refname)
            if r is not None and r.shape == t.shape:
                e = err(t, r)
                print(f"{name:45s} {e[0]:9.4f} {e[1]:9.4f}")
            continue
        r = ref_for(name)
        if r is None or r.shape != t.shape:
            print(f"{name:45s} (no ref / shape {tuple(t.shape)} vs {None if r is None else tuple(r.shape)})")
            continue
        e = err(t, r)
        print(f"{name:45s} {e[0]:9.4f} {e[1]:9.4f}")
    e = err(logits.detach(), ref_logits)
    print(f"{'LOGITS':45s} {e[0]:9.4f} {e[1]:9.4f}")
    if mode == "train":
        refg = dict(o.named_parameters())
        print(f"\n{'param grad':55s} {'relL2':>9s} {'cos':>9s} {'|ref|':>10s}")
        for name, p in m.named_parameters():
            a, b = p.grad.float().cpu().flatten(), refg[name].grad.flatten()
            l2 = float((a - b).norm() / (b.norm() + 1e-20))
            c = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
            flag = "  <<<" if (l2 > 5e-2 and b.norm() > 1e-6) else ""
            print(f"{name:55s} {l2:9.4f} {c:9.5f} {float(b.norm()):10.3e}{flag}")


if __name__ == "__main__":
    main()
