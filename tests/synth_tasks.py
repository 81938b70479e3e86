"""Learnable synthetic segmentation task for the full-size parity tests (test infrastructure).

The benchmark workload is U(-1,1) noise with i.i.d. random labels: nothing can be learnt from it, so a
network "warmed" on it stays as ill-conditioned as at random init and the teacher never clears the 0.95 confidence
threshold (every consistency mask is empty).  The parity tests need the regime of a real run -- BatchNorm statistics of
structured activations, confident teacher, mask means of 0.3-0.6 -- so here the label IS a function of the image:
a smooth latent field z (low-resolution Gaussian noise, bicubic x16) is rendered into the image channels
(tanh(a_c z + b_c) + pixel noise) and quantised into the classes.  Shapes, dtypes, value ranges and the CutMix
inputs are those of ``oracle.ssl_step_ref.synthetic_batch``.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _latent(B, H, W, g, cell=16):
    z = torch.randn(B, 1, max(H // cell, 2) + 2, max(W // cell, 2) + 2, generator=g)
    z = F.interpolate(z, size=(H + 2 * cell, W + 2 * cell), mode="bicubic", align_corners=False)
    return z[:, :, cell:cell + H, cell:cell + W].contiguous()


def _render(z, n_channels, g, noise=0.15):
    a = torch.tensor([1.6, -1.1, 0.7, 2.2][:n_channels]).view(1, -1, 1, 1)
    b = torch.tensor([0.0, 0.3, -0.4, 0.1][:n_channels]).view(1, -1, 1, 1)
    x = torch.tanh(a * z + b) + noise * torch.randn(z.shape[0], n_channels, z.shape[2], z.shape[3], generator=g)
    return x.clamp(-1, 1)


def _labels(z, n_classes, branch):
    if branch == "softmax":
        # equal-mass class bands of a unit Gaussian field
        edges = torch.tensor([float(np.sqrt(2) * _erfinv(2 * q - 1)) for q in np.linspace(0, 1, n_classes + 1)[1:-1]])
        return torch.bucketize(z[:, 0], edges)                                  # int64 [B,H,W]
    # fundus-like nested structures: channel 0 = cup (z > 0.5), channel 1 = disc (z > -0.2)
    thr = [0.5, -0.2, 0.0, 0.8][:n_classes]
    return torch.cat([(z > t).float() for t in thr], dim=1)                     # float [B,C,H,W]


def _erfinv(y):
    return float(torch.erfinv(torch.tensor(float(y), dtype=torch.float64)))


def _box(H, W, rng, lo=0.02, hi=0.4, r_lo=0.3, r_hi=1 / 0.3):
    area = rng.uniform(lo, hi) * H * W
    while True:
        r = rng.uniform(r_lo, r_hi)
        bw, bh = int(np.sqrt(area / r)), int(np.sqrt(area * r))
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        if x0 + bw <= W and y0 + bh <= H:
            out = np.zeros((H, W), np.float32)
            out[y0:y0 + bh, x0:x0 + bw] = 1
            return out


def blob_batch(n_channels, n_classes, H, W, B_l, B_u, seed=0, branch="softmax"):
    """Same keys / dtypes as ``synthetic_batch``; the unlabelled images follow the same image<->label law."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    z_l, z_u = _latent(B_l, H, W, g), _latent(B_u, H, W, g)
    lb_x, ulb_w = _render(z_l, n_channels, g), _render(z_u, n_channels, g)
    ulb_s = (ulb_w * (1 + 0.2 * (torch.rand(B_u, 1, 1, 1, generator=g) - 0.5))
             + 0.1 * torch.randn(B_u, n_channels, H, W, generator=g)).clamp(-1, 1)
    lb_mask = _labels(z_l, n_classes, branch)
    cut_mask = torch.ones(B_l, 1 if branch == "softmax" else n_classes, H, W)
    box = torch.from_numpy(np.stack([_box(H, W, rng) for _ in range(B_u)]))
    choice = torch.from_numpy(rng.integers(0, B_l, B_u)).long()
    # the style-moved partner image (train.py:628-636: the CutMix partner re-styled towards the unlabelled image)
    move = (lb_x[choice] * (1 + 0.1 * (torch.rand(B_u, 1, 1, 1, generator=g) - 0.5)) + 0.05 * torch.randn(B_u, n_channels, H, W, generator=g)).clamp(-1, 1)
    return dict(lb_x=lb_x, lb_mask=lb_mask, ulb_w=ulb_w, ulb_s=ulb_s, move_transx=move, box=box, choice=choice,
                cut_img=lb_x.clone(), cut_label=lb_mask.clone(), cut_mask=cut_mask)


def to_device_batch(batch, device="cuda", compact=True):
    """Batch for ``SSLTrainer.step``: uint8 label/mask planes and int32 ``choice`` (what the kernels read)."""
    out = {k: v.to(device) for k, v in batch.items()}
    if compact:
        for k in ("lb_mask", "cut_label", "cut_mask", "box"):
            out[k] = out[k].to(torch.uint8)
        out["choice"] = out["choice"].to(torch.int32)
    return out
