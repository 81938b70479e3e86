"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's frequency-domain style mix.

Follows /root/reference/train.py line by line -- ``extract_amp_spectrum`` (train.py:158-165),
``low_freq_mutate_np`` (train.py:167-187), ``source_to_target_freq`` (train.py:189-207) and the call site
train.py:628-636 -- with the only source of randomness (``random.uniform(0, degree)``, train.py:180) turned into
the explicit ``ratio`` argument.  Pinned bit-for-bit against the reference's own functions by
``oracle/make_golden.py`` (fixtures: tests/golden/fft_mix.npz)."""
import numpy as np


def extract_amp_spectrum(img_np):                                     # train.py:158-165
    fft = np.fft.fft2(img_np, axes=(-2, -1))
    return np.abs(fft)


def low_freq_mutate(amp_src, amp_trg, L, ratio):                      # train.py:167-187
    a_src = np.fft.fftshift(amp_src, axes=(-2, -1))
    a_trg = np.fft.fftshift(amp_trg, axes=(-2, -1))
    _, h, w = a_src.shape
    b = (np.floor(np.amin((h, w)) * L)).astype(int)
    c_h = np.floor(h / 2.0).astype(int)
    c_w = np.floor(w / 2.0).astype(int)
    h1, h2, w1, w2 = c_h - b, c_h + b + 1, c_w - b, c_w + b + 1
    a_src[:, h1:h2, w1:w2] = a_src[:, h1:h2, w1:w2] * (1 - ratio) + a_trg[:, h1:h2, w1:w2] * ratio
    return np.fft.ifftshift(a_src, axes=(-2, -1))


def source_to_target_freq(src_img, amp_trg, L, ratio):               # train.py:189-207
    fft_src = np.fft.fft2(src_img, axes=(-2, -1))
    amp_src, pha_src = np.abs(fft_src), np.angle(fft_src)
    amp_src_ = low_freq_mutate(amp_src, amp_trg, L, ratio)
    fft_src_ = amp_src_ * np.exp(1j * pha_src)
    return np.real(np.fft.ifft2(fft_src_, axes=(-2, -1)))


def move_transx(mix_img, ulb_x_w, ratios, L=0.01, fft_dtype=None):   # train.py:628-636
    """mix_img, ulb_x_w: float32 arrays [N,C,H,W] in [-1,1]; ratios: one float per sample.  Returns float32.

    ``fft_dtype``: numpy >= 2.0 runs pocketfft in the input precision, so with the float32 images the reference
    hands it the transforms are SINGLE precision there (numpy 1.x up-cast to float64).  None reproduces the installed
    numpy exactly (what the fixtures pin); np.float64 evaluates the same formulas in double (the yardstick for the
    device kernels, which compute in float64)."""
    out = []
    for i in range(len(mix_img)):
        trg255 = (ulb_x_w[i] + 1) * 127.5
        src255 = (mix_img[i] + np.float32(1)) * np.float32(127.5)
        if fft_dtype is not None:
            trg255, src255 = trg255.astype(fft_dtype), src255.astype(fft_dtype)
        amp_trg = extract_amp_spectrum(trg255)
        img_freq = source_to_target_freq(src255, amp_trg, L, ratios[i])
        out.append(np.clip(img_freq, 0, 255).astype(np.float32))
    res = np.array(out, dtype=np.float32)
    return res / np.float32(127.5) - np.float32(1)
