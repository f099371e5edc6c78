"""Worst gradient rel-L2 of the CUDA training step against autograd through the CPU oracle, per A3GC_BWD_MMA mode.
python tests/diag_train_parity.py   (diagnostic; the bound the tests enforce is 1e-4)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import a3gc_ip_b200 as A
from oracle import net_oracle as O
from util import NET_CLS_NAMES


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    nira = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nira_template_15_norm.pt")).float()
    for variant, hidden, B, T in (("A3GC", 256, 4, 12), ("A3GC", 128, 6, 12), ("A3GC", 64, 8, 12), ("AAGC", 256, 3, 8)):
        f0, out = 15, 9
        sd = O.random_state_dict(variant, f0, out, hidden, nira, seed=21)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(B, T, 15, f0, generator=g)
        target = torch.randn(B, T, 15 * out, generator=g)
        sdr = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
        xr = x.clone().requires_grad_(True)
        y, _ = O.net_forward(variant, xr, sdr, None)
        torch.mean(torch.sum(torch.square(target - y.reshape(target.shape)), -1)).backward()
        for mode in ("0", "1", "2", "4"):
            os.environ["A3GC_BWD_MMA"] = mode
            net = getattr(A, NET_CLS_NAMES[variant])(f0, out, hidden, nira, linear_dropout=0.0, dropout=0.0, recurrent_dropout=0.0)
            net.load_state_dict(sd, strict=True)
            net = net.cuda().train()
            xc = x.cuda().requires_grad_(True)
            yc, _ = net(xc)
            A.pose_loss()(yc.view(B, T, 15 * out), target.cuda()).backward()
            worst = max(((rel_l2(p.grad.cpu(), sdr[n].grad), n) for n, p in net.named_parameters() if p.grad is not None))
            print(f"{variant} H={hidden} B={B} T={T} A3GC_BWD_MMA={mode}: dx {rel_l2(xc.grad.cpu(), xr.grad):.2e}  "
                  f"worst parameter gradient {worst[0]:.2e} ({worst[1]})", flush=True)


if __name__ == "__main__":
    main()
