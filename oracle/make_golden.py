"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.npz from the REAL reference.

Run in the authoring container (needs /root/reference):  python -m oracle.make_golden
Inputs and weights are reproducible from the seeds below (torch CPU generator), so the
fixtures hold only compact views of the reference's outputs: strided samples, sums and
per-parameter gradient summaries.
"""
import os
from collections import OrderedDict

import numpy as np
import torch

from oracle import reference_shim
from oracle import ref_torch as R
from vae_segmentation_b200.synthetic import synth_image, synth_label

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SEG_CASE = dict(seed=100, batch=2, patch=32)
VAE_CASE = dict(seed=101, batch=1, patch=128)
JOINT_CASE = dict(seed=102, batch=1, patch=128)
LOSS_CASE = dict(seed=103)
COND_CASE = dict(steps=4, patch_seg=32, patch_vae=64, lr=0.1)     # pinned prefix of the conditioned recipes


def grad_summary(grads):
    """per-parameter [sum, l2, first 4 flattened values] -- compact gradient fingerprint."""
    rows = []
    for k, g in grads.items():
        f = g.detach().double().flatten()
        head = torch.zeros(4, dtype=torch.float64)
        head[: min(4, f.numel())] = f[:4]
        rows.append(torch.cat([f.sum()[None], f.norm()[None], head]))
    return torch.stack(rows).numpy()


def sample(t, step=3):
    return t.detach()[..., ::step, ::step, ::step].contiguous().numpy()


def case_inputs(case, vae=False, seg=False):
    """Deterministic weights + inputs for a case; the RNG draw order here is the contract
    the tests replay: seg weights, vae weights, image, label."""
    torch.manual_seed(case["seed"])
    seg_sd = R.init_seg_state() if seg else None
    vae_sd = R.init_vae_state(2, 128, case["patch"]) if vae else None
    img = synth_image(case["batch"], case["patch"])
    label = synth_label(case["batch"], case["patch"])
    return seg_sd, vae_sd, img, label


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    jm, ev = reference_shim.load()
    import importlib.util
    # main_source.py's eps=1e-4 avg_dsc twin cannot be imported (module-level argparse);
    # evaluation.avg_dsc differs only in eps, handled analytically below.

    # ---- Segmentation train step (main_source.py:415-437 semantics; evaluation eps) ----
    seg_sd, _, img, label = case_inputs(SEG_CASE, seg=True)
    seg = jm.Segmentation(1, 2, norm_type=1)
    seg.load_state_dict(seg_sd, strict=True)
    batch = {"img": img}
    batch = seg(batch, "img", "pred")
    batch["onehot"] = R.one_hot(label)
    loss = 1 - ev.avg_dsc(batch, source_key="pred", target_key="onehot", botindex=1, topindex=2)
    loss.backward()
    grads = OrderedDict((k, p.grad) for k, p in seg.named_parameters())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "seg_p32.npz"),
                        probs_sample=sample(batch["pred"]),
                        probs_sum=batch["pred"].detach().sum((2, 3, 4)).numpy(),
                        loss=loss.detach().numpy(), grad_summary=grad_summary(grads),
                        grad_names=np.array(list(grads.keys())))
    print("seg_p32 loss", float(loss))

    # ---- VAE train step at the reference's native 128^3 (main_source.py:389-406) ----
    _, vae_sd, _, label = case_inputs(VAE_CASE, vae=True)
    vae = jm.VAE(2, 2, norm_type=1, dim=128)
    vae.load_state_dict(vae_sd, strict=True)
    oh = R.one_hot(label)
    torch.manual_seed(VAE_CASE["seed"] + 1000)     # z comes from the CPU generator (F11)
    recon, mean, std = vae(oh, if_random=True, scale=0.35)
    d = {"recon": recon, "onehot": oh, "mean": mean, "std": std}
    kl = ev.KLloss(d)
    dsc = 1 - ev.avg_dsc(d, source_key="recon", target_key="onehot", botindex=1, topindex=2)
    loss = dsc + 0.00002 * kl
    loss.backward()
    grads = OrderedDict((k, p.grad) for k, p in vae.named_parameters())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "vae_p128.npz"),
                        recon_sample=sample(recon, 5), recon_sum=recon.detach().sum((2, 3, 4)).numpy(),
                        mean=mean.detach().numpy(), std=std.detach().numpy(),
                        kl=kl.detach().numpy(), dsc=dsc.detach().numpy(), loss=loss.detach().numpy(),
                        grad_summary=grad_summary(grads), grad_names=np.array(list(grads.keys())))
    print("vae_p128 loss", float(loss), "kl", float(kl))

    # ---- joint teacher-student step (main_target.py:520-592,734-736), type 0, lambda 1 ----
    seg_sd, vae_sd, img, label = case_inputs(JOINT_CASE, vae=True, seg=True)
    student = jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128)])
    student.Seg.load_state_dict(seg_sd)
    student.Vae.load_state_dict(vae_sd)
    for p in student.Vae.parameters():
        p.requires_grad = False
    student.Vae.eval()
    teacher = jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128)])
    teacher.load_state_dict(student.state_dict())
    for p in teacher.parameters():
        p.requires_grad = False
    b = {"img": img, "only": R.one_hot(label)}
    b = student(b, "img", "pred", "recon_pred", dropout=True)
    b = teacher(b, "img", "only_fake", "asdf")
    b["only_fake"] = ev.binarize(b["only_fake"])
    recon_loss = 1 - ev.avg_dsc(b, source_key="pred", target_key="recon_pred", botindex=1, topindex=2)
    klloss = ev.KLloss(b)
    dsc_loss = 1 - ev.avg_dsc(b, source_key="pred", target_key="only", botindex=1, topindex=2)
    dsc_fake = 1 - ev.avg_dsc(b, source_key="pred", target_key="only_fake", botindex=1, topindex=2)
    final = 1.0 * recon_loss + dsc_fake
    final.backward()
    grads = OrderedDict((k, p.grad) for k, p in student.Seg.named_parameters())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "joint_p128.npz"),
                        pred_sample=sample(b["pred"], 5), recon_sample=sample(b["recon_pred"], 5),
                        recon_loss=recon_loss.detach().numpy(), klloss=klloss.detach().numpy(),
                        dsc_loss=dsc_loss.detach().numpy(), dsc_loss_fake=dsc_fake.detach().numpy(),
                        final=final.detach().numpy(), mean=b["mean"].detach().numpy(),
                        std=b["std"].detach().numpy(),
                        grad_summary=grad_summary(grads), grad_names=np.array(list(grads.keys())))
    print("joint_p128 final", float(final), "recon", float(recon_loss))

    # ---- conditioned fixtures (oracle/conditioned.py): the same K-step recipes through the REAL reference modules and
    #      torch.optim.SGD; tests/test_oracle.py checks the oracle-trained weights against these checksums ----
    from oracle import conditioned as C

    def real_seg_step(sd, img, label):
        with torch.random.fork_rng():                 # module construction draws from the generator the recipe's data uses
            seg = jm.Segmentation(1, 2, norm_type=1)
        seg.load_state_dict(sd, strict=True)
        b = seg({"img": img}, "img", "pred")
        b["onehot"] = R.one_hot(label)
        loss = 1 - ev.avg_dsc(b, source_key="pred", target_key="onehot", botindex=1, topindex=2)
        loss.backward()
        return loss.detach(), OrderedDict((k, p.grad) for k, p in seg.named_parameters())

    def real_vae_step(sd, label, z):
        flat = sd["fc_mean.weight"].shape[1]
        with torch.random.fork_rng():
            vae = jm.VAE(2, 2, norm_type=1, dim=128)
            # the reference hard-codes the 128^3 flat dimension: re-size its three Linear layers for the fixture's patch
            vae.fc_mean = torch.nn.Linear(flat, 128)
            vae.fc_std = torch.nn.Linear(flat, 128)
            vae.fc2 = torch.nn.Linear(128, flat)
        vae.load_state_dict(sd, strict=True)
        oh = R.one_hot(label)
        side = round((flat // 256) ** (1.0 / 3.0))
        # VAE.forward with the generalised view (joint_model.py:233-266 verbatim otherwise)
        out = vae.down5(vae.down4(vae.down3(vae.down2(vae.down1(vae.in_block(oh))))))
        out = out.view(out.size(0), flat)
        mean, std = vae.fc_mean(out), torch.relu(vae.fc_std(out))
        lat = mean + z * std * 0.35
        h = vae.fc2(lat).view(out.size(0), 256, side, side, side)
        recon = vae.final(vae.out_block(vae.up5(vae.up4(vae.up3(vae.up2(vae.up1(h)))))))
        d = {"recon": recon, "onehot": oh, "mean": mean, "std": std}
        loss = 1 - ev.avg_dsc(d, source_key="recon", target_key="onehot", botindex=1, topindex=2) + 0.00002 * ev.KLloss(d)
        loss.backward()
        return loss.detach(), OrderedDict((k, p.grad) for k, p in vae.named_parameters())

    K = COND_CASE["steps"]
    sd_real, l_real = C.train_seg(K, patch=COND_CASE["patch_seg"], lr=COND_CASE["lr"], step_fn=real_seg_step)
    vd_real, lv_real = C.train_vae(K, patch=COND_CASE["patch_vae"], lr=COND_CASE["lr"], step_fn=real_vae_step)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "conditioned.npz"), seg_losses=np.array(l_real), vae_losses=np.array(lv_real),
                        seg_checksum=C.checksum(sd_real), vae_checksum=C.checksum(vd_real))
    print("conditioned seg losses", l_real, "vae", lv_real)

    # ---- loss functions on small random tensors (utils/evaluation.py) ----
    torch.manual_seed(LOSS_CASE["seed"])
    a = torch.rand(3, 2, 6, 7, 8)
    t = torch.rand(3, 2, 6, 7, 8)
    mean = torch.randn(3, 128)
    std = torch.rand(3, 128)
    d = {"a": a, "t": t, "mean": mean, "std": std}
    out = dict(
        a=a.numpy(), t=t.numpy(), mean=mean.numpy(), std=std.numpy(),
        dsc_full=ev.avg_dsc(d, "a", "t").numpy(),
        dsc_fg=ev.avg_dsc(d, "a", "t", botindex=1, topindex=2).numpy(),
        dsc_fg_vec=ev.avg_dsc(d, "a", "t", botindex=1, topindex=2, return_mean=False).numpy(),
        dsc_binary=ev.avg_dsc(d, "a", "t", binary=True, botindex=1, topindex=2).numpy(),
        kl=ev.KLloss(d).numpy(), binarize=ev.binarize(a).numpy(),
        confident=ev.confident_binarize(a).numpy(), dice=ev.dice(a, t).numpy())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "losses.npz"), **out)
    print("losses ok")


if __name__ == "__main__":
    main()
