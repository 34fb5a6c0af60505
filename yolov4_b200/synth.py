"""Seeded synthetic head outputs and labels.

There is no dataset or checkpoint in the build environment, so benchmarks and parity tests run on
synthetic raw head tensors with the "trained-like" mixture the survey calibrated against the
reference: per cell fg ~ Bernoulli(0.005); t_xy ~ N(0,1); t_wh ~ N(0,0.5^2); obj ~ N(0,2^2) for fg
cells else N(-10,1.5^2); cls_k ~ N(-6,2^2) with one random hot class ~ N(2,2^2) in fg cells.
At 608x608 this yields ~11.5k surviving (box,class) pairs per image at conf 1e-4 and ~110 at 0.2.
"""
import torch

STRIDES = (8, 16, 32)


def synth_head_outputs(batch, img_size, n_classes=80, seed=0, device="cpu", fg_prob=0.005,
                       clustered=False, dtype=torch.float32):
    """Returns [raw0, raw1, raw2], raw_l contiguous [B, 3*(5+C), F_l, F_l] (yolov4.py:235-251).

    clustered=True draws fg flag / hot class / t_wh on a half-resolution grid and repeats them 2x2, so
    neighbouring cells predict same-class boxes whose IoU lands on both sides of the NMS threshold.
    """
    g = torch.Generator(device=device).manual_seed(seed)
    C = n_classes
    outs = []

    def randn(*shape):
        return torch.randn(*shape, generator=g, device=device, dtype=dtype)

    for s in STRIDES:
        F = img_size // s
        t = torch.empty(batch, 3, 5 + C, F, F, device=device, dtype=dtype)
        if clustered:
            Fh = (F + 1) // 2
            up = lambda v: v.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1)[..., :F, :F]
            fg = up(torch.rand(batch, 3, 1, Fh, Fh, generator=g, device=device) < fg_prob * 4)
            hot = up(torch.randint(0, C, (batch, 3, 1, Fh, Fh), generator=g, device=device))
            twh = up(randn(batch, 3, 2, Fh, Fh) * 0.5)
        else:
            fg = torch.rand(batch, 3, 1, F, F, generator=g, device=device) < fg_prob
            hot = torch.randint(0, C, (batch, 3, 1, F, F), generator=g, device=device)
            twh = randn(batch, 3, 2, F, F) * 0.5
        t[:, :, 0:2] = randn(batch, 3, 2, F, F)
        t[:, :, 2:4] = twh
        t[:, :, 4:5] = torch.where(fg, randn(batch, 3, 1, F, F) * 2.0, randn(batch, 3, 1, F, F) * 1.5 - 10.0)
        cls = randn(batch, 3, C, F, F) * 2.0 - 6.0
        hotval = randn(batch, 3, 1, F, F) * 2.0 + 2.0
        is_hot = fg & (torch.arange(C, device=device).view(1, 1, C, 1, 1) == hot)
        t[:, :, 5:] = torch.where(is_hot, hotval, cls)
        outs.append(t.view(batch, 3 * (5 + C), F, F))
    return outs


def synth_labels(batch, img_size, n_valid=50, max_labels=60, n_classes=80, seed=0, device="cpu"):
    """Padded labels [B, max_labels, 5] = (xc, yc, w, h, cls) in input pixels, zero rows after n_valid
    (layout of yolo/data/transform.py:464-471; config 4 of BASELINE.json)."""
    g = torch.Generator(device=device).manual_seed(seed)
    lab = torch.zeros(batch, max_labels, 5, device=device)
    wh = torch.rand(batch, n_valid, 2, generator=g, device=device) * 300.0 + 8.0
    wh = torch.minimum(wh, torch.full_like(wh, float(img_size) - 1.0))
    c = torch.rand(batch, n_valid, 2, generator=g, device=device) * (img_size - wh) + wh / 2
    cls = torch.randint(0, n_classes, (batch, n_valid), generator=g, device=device)
    lab[:, :n_valid, 0:2] = c
    lab[:, :n_valid, 2:4] = wh
    lab[:, :n_valid, 4] = cls.to(lab.dtype)
    return lab
