"""Synthetic workload for bench.py and tools/: host-side batch of the shapes the reference's loaders yield
(train.py:560-600: labelled image+mask, weak/strong unlabelled views, the style-moved view, CutMix boxes and
the paste-bank draw) plus the per-image FLOP model used for the roofline.  Pure data generation -- no model,
no loss; kept in the product package so the measured arm never imports the test oracle."""
import math

import numpy as np
import torch


def _paste_box(H, W, rng, lo=0.02, hi=0.4, r_lo=0.3, r_hi=1 / 0.3):
    """One CutMix rectangle, area fraction in [lo, hi] and aspect in [r_lo, r_hi] (train.py:222-240, p=1)."""
    area = rng.uniform(lo, hi) * H * W
    while True:
        r = rng.uniform(r_lo, r_hi)
        bw, bh = int(math.sqrt(area / r)), int(math.sqrt(area * r))
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        if x0 + bw <= W and y0 + bh <= H:
            out = np.zeros((H, W), np.float32)
            out[y0:y0 + bh, x0:x0 + bw] = 1
            return out


def synthetic_batch(n_channels, n_classes, H, W, B_l, B_u, seed=1337, branch="softmax"):
    """dict of CPU tensors with the keys SSLTrainer.step expects (images in [-1, 1] fp32)."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)

    def img(b):
        return torch.rand(b, n_channels, H, W, generator=g) * 2 - 1

    lb_x, ulb_w, move = img(B_l), img(B_u), img(B_u)
    ulb_s = (ulb_w + 0.1 * torch.randn(B_u, n_channels, H, W, generator=g)).clamp(-1, 1)
    if branch == "softmax":
        lb_mask = torch.randint(0, n_classes, (B_l, H, W), generator=g)
        cut_mask = torch.ones(B_l, 1, H, W)
    else:
        lb_mask = torch.randint(0, 2, (B_l, n_classes, H, W), generator=g).float()
        cut_mask = torch.ones(B_l, n_classes, H, W)
    box = torch.from_numpy(np.stack([_paste_box(H, W, rng) for _ in range(B_u)]))
    choice = torch.from_numpy(rng.integers(0, B_l, B_u)).long()
    return dict(lb_x=lb_x, lb_mask=lb_mask, ulb_w=ulb_w, ulb_s=ulb_s, move_transx=move, box=box, choice=choice,
                cut_img=lb_x.clone(), cut_label=lb_mask.clone(), cut_mask=cut_mask)


def conv_flops_unet_a(n_channels, n_classes, H, W):
    """Algorithmic conv + transposed-conv FLOPs (2 x MAC) per image, one forward of UNet-A."""
    def double_conv(cin, cout, h, w):
        return 2 * 9 * h * w * (cin * cout + cout * cout)

    total = double_conv(n_channels, 64, H, W)
    ch, h, w = 64, H, W
    for _ in range(4):                      # encoder
        h, w = h // 2, w // 2
        total += double_conv(ch, 2 * ch, h, w)
        ch *= 2
    for _ in range(4):                      # decoder: convT k2 s2 then DoubleConv on the concat
        total += 2 * h * w * ch * (ch // 2) * 4
        h, w = 2 * h, 2 * w
        total += double_conv(ch, ch // 2, h, w)
        ch //= 2
    return total + 2 * H * W * 64 * n_classes
