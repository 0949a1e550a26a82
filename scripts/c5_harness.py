"""Stand-in for the reference's CausalVQAE conv stacks (BASELINE configs[4]: causal Conv1d encoder -> RVQ ->
decoder on 24 kHz audio), for machines where /root/reference does not exist (the GPU boxes).

Only the SHAPE PLAN is taken from the reference (probed from /root/reference/networks/vae.py
with config defaults): channels 1 -> 32 -> 64 -> 128 -> 256 -> 512 -> 1024 -> (k = 3) -> 512 latent, strides
(2, 3, 4, 4, 5) = 480 samples per latent frame, a k = 7 stride-1 transposed convolution opening the decoder and
nearest-neighbour upsampling + convolution for the strided decoder blocks; the quantizer is imported exactly as the
reference imports it (vae.py:6) and called on the "b c l -> b l c" VIEW (vae.py:313-318).  The layers themselves are a
generic SoundStream-style stack written here, not the reference's code: the conv stacks are out of this repo's scope
(SURVEY.md section 8), they only feed the quantizer latents of the real shape, strides and memory layout.
"""
import torch
from torch import nn
import torch.nn.functional as F

STRIDES = (2, 3, 4, 4, 5)
CHANNELS = (32, 64, 128, 256, 512, 1024)
LATENT = 512


class CausalConv(nn.Module):
    def __init__(self, cin, cout, k, stride=1, dilation=1):
        super().__init__()
        self.pad = (k - 1) * dilation - (stride - 1)
        self.conv = nn.Conv1d(cin, cout, k, stride=stride, dilation=dilation)

    def forward(self, x):
        return self.conv(F.pad(x, (max(self.pad, 0), 0)))


class ResUnit(nn.Module):
    def __init__(self, c, dilation):
        super().__init__()
        self.a = CausalConv(c, c, 7, dilation=dilation)
        self.b = CausalConv(c, c, 1)

    def forward(self, x):
        return x + self.b(F.elu(self.a(F.elu(x))))


class EncBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.res = nn.Sequential(ResUnit(cin, 1), ResUnit(cin, 3), ResUnit(cin, 9))
        self.down = CausalConv(cin, cout, 2 * stride + 1, stride=stride)

    def forward(self, x):
        return self.down(F.elu(self.res(x)))


class DecBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.stride = stride
        self.up = CausalConv(cin, cout, 2 * stride + 1)
        self.res = nn.Sequential(ResUnit(cout, 1), ResUnit(cout, 3), ResUnit(cout, 9))

    def forward(self, x):
        x = F.interpolate(F.elu(x), scale_factor=self.stride, mode="nearest")
        return self.res(self.up(x))


class SyntheticCausalVQAE(nn.Module):
    """forward(x, update_codebook, codebook_n) -> (y, commit_loss, index): the reference model's call contract
    (vae.py:293-322), return order included."""

    def __init__(self, num_quantizers=8, codebook_size=1024, vq_type="ema", vq_cutoff_freq=1, use_som=True,
                 som_kernel_type="hard", width=1.0):
        super().__init__()
        from som_quantizer import ResidualQuantizer          # the reference's import line (vae.py:6)
        ch = [max(8, int(c * width)) for c in CHANNELS]
        self.stem = CausalConv(1, ch[0], 7)
        self.enc = nn.Sequential(*[EncBlock(ch[i], ch[i + 1], s) for i, s in enumerate(STRIDES)])
        self.to_latent = CausalConv(ch[-1], LATENT, 3)
        self.quantizer = ResidualQuantizer(num_quantizers=num_quantizers, dim=LATENT, quantizer_class=vq_type,
                                           codebook_sizes=codebook_size, vq_cutoff_freq=vq_cutoff_freq, use_som=use_som,
                                           som_kernel_type=som_kernel_type)
        self.from_latent = nn.ConvTranspose1d(LATENT, ch[-1], 7, stride=1, padding=3)
        self.dec = nn.Sequential(*[DecBlock(ch[i + 1], ch[i], s) for i, s in reversed(list(enumerate(STRIDES)))])
        self.head = CausalConv(ch[0], 1, 7)

    def encode(self, x, update_codebook=False, codebook_n=None, prioritize_early=False):
        z = self.to_latent(F.elu(self.enc(self.stem(x))))
        z = z.permute(0, 2, 1)                                # "b c l -> b l c": a VIEW, strides (C L, 1, L)
        zq, index, commit = self.quantizer(z, codebook_n, update_codebook=update_codebook,
                                           prioritize_early=prioritize_early)
        return zq.permute(0, 2, 1), commit, index

    def forward(self, x, update_codebook=False, codebook_n=None, prioritize_early=False):
        zq, commit, index = self.encode(x, update_codebook, codebook_n, prioritize_early)
        y = self.head(F.elu(self.dec(self.from_latent(zq))))
        return y, commit, index
