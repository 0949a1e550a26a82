"""Exercise script (compute-sanitizer is closed on this pool; run it plainly, the asserts are the check) for the third-session kernels: a training step with the SOM neighbourhood and
stale-code re-seeding on (ragged codebook sizes, the reference's strided frame layout), the backward pass for both
quantizer classes, the wire format.  python scripts/exercise_maintenance.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_generation_b200 import ResidualQuantizer

torch.manual_seed(0)
for d, sizes, kernel in ((128, [512, 300, 64], "hard"), (256, [256, 256, 100], "gaussian")):
    m = ResidualQuantizer(len(sizes), d, "ema", sizes, vq_cutoff_freq=1.0, use_som=True, som_kernel_type=kernel)
    with torch.no_grad():
        m.ema_count[0, :7] = 0.1
        m.ema_count[2, 3] = 0.1
    m = m.cuda().train()
    xs = torch.randn(3, d, 171, device="cuda")            # odd L: ragged last tile, strided (B, L, d) view
    for _ in range(2):
        with torch.no_grad():
            _, idx, _ = m(xs.permute(0, 2, 1), None, update_codebook=True)
    torch.cuda.synchronize()
    print("update ok", d, m.n_replaced.tolist(), flush=True)
    packed = m.pack_indices(idx)
    assert torch.equal(m.unpack_indices(packed), idx)
    for cls in ("ema", "base"):
        b = ResidualQuantizer(2, d, cls, 128).cuda()
        for strided in (False, True):
            leaf = torch.randn(2, d, 77, device="cuda", requires_grad=True) if strided else \
                torch.randn(2, 77, d, device="cuda", requires_grad=True)
            x = leaf.permute(0, 2, 1) if strided else leaf
            out, _, commit = b(x)
            (out.sum() + commit).backward()
            torch.cuda.synchronize()
            assert leaf.grad is not None and torch.isfinite(leaf.grad).all()
    print("backward ok", d, flush=True)
print("done")
