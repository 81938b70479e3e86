"""Frequency-domain style mix on the device.

Replaces the per-sample host loop of the reference (train.py:628-636: D2H copy, numpy fft2 / ifft2 through
``extract_amp_spectrum`` train.py:158-165, ``low_freq_mutate_np`` :167-187, ``source_to_target_freq`` :189-207,
H2D copy) by four thin float64 DFT kernels that only touch the (2b+1)^2 low-frequency bins the mix changes.

    move_transx = amp_mix(mix_img, ulb_x_w, ratio, L=args.LB)      # train.py:628-636 in one call

``ratio`` is the host random number of train.py:180 (``random.uniform(0, iter_num / max_iterations)``), one per
sample -- host RNG stays a step input, as everywhere in this repo."""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L
from .engine import _call, _ptr, _stream


def amp_mix(src: torch.Tensor, trg: torch.Tensor, ratio, L_window: float = 0.01) -> torch.Tensor:
    """src (``mix_img``), trg (``ulb_x_w``): fp32 NCHW CUDA tensors in [-1,1]; ratio: float | sequence | tensor of
    one blend ratio per sample.  Returns ``move_transx`` (fp32 NCHW in [-1,1])."""
    L.require_device()
    if src.dim() != 4 or src.shape != trg.shape:
        raise ValueError("amp_mix: src and trg must be NCHW tensors of the same shape")
    if not src.is_cuda:
        raise RuntimeError("amp_mix has no CPU path")
    src = src.float().contiguous()
    trg = trg.float().contiguous()
    N, C, H, W = src.shape
    r = torch.as_tensor(ratio, dtype=torch.float64)
    if r.dim() == 0:
        r = r.repeat(N)
    if r.numel() != N:
        raise ValueError("amp_mix: one ratio per sample expected")
    r = r.to(src.device).contiguous()
    nbytes = int(L.lib.ustrun_fft_amp_mix_workspace_bytes(N, C, H, W, float(L_window)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=src.device)
    out = torch.empty_like(src)
    _call("ustrun_fft_amp_mix", _ptr(src), _ptr(trg), _ptr(r), ctypes.c_double(float(L_window)), _ptr(out), N, C, H, W, _ptr(ws), nbytes, _stream())
    return out
