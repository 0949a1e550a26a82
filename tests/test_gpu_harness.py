"""GPU test of the drop-in boundary inside a model shaped like the reference's CausalVQAE
(/root/reference/networks/vae.py:293-322): conv encoder -> einops "b c l -> b l c" VIEW -> quantizer ->
"b l c -> b c l" -> conv decoder, trained for a few steps the way vae.py:385-392 does
(mse + commit loss, Adam, update_codebook=True).  /root/reference does not exist on the GPU box, so the conv
stacks here are a small stand-in; the quantizer is imported exactly as the reference imports it."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class TinyCausalVQAE(torch.nn.Module):
    def __init__(self, d=128, nq=4, K=256, vq_type="ema"):
        super().__init__()
        from som_quantizer import ResidualQuantizer, tuple_checker      # the reference's import line (vae.py:6)
        import einops
        self.einops = einops
        self.codebook_size = tuple_checker(K, nq)
        self.quantizer = ResidualQuantizer(num_quantizers=nq, dim=d, quantizer_class=vq_type, codebook_sizes=K,
                                           vq_cutoff_freq=1, use_som=True, som_kernel_type="hard")
        self.enc = torch.nn.Sequential(torch.nn.Conv1d(1, 32, 7, stride=4, padding=3), torch.nn.ELU(),
                                       torch.nn.Conv1d(32, d, 7, stride=4, padding=3))
        self.dec = torch.nn.Sequential(torch.nn.ConvTranspose1d(d, 32, 8, stride=4, padding=2), torch.nn.ELU(),
                                       torch.nn.ConvTranspose1d(32, 1, 8, stride=4, padding=2))

    def forward(self, x, update_codebook=False, codebook_n=None, prioritize_early=False):
        z = self.enc(x)
        z = self.einops.rearrange(z, "b c l -> b l c")                  # non-contiguous view (vae.py:313)
        assert not z.is_contiguous()
        zq, index, commit = self.quantizer(z, codebook_n, update_codebook=update_codebook,
                                           prioritize_early=prioritize_early)
        zq = self.einops.rearrange(zq, "b l c -> b c l")
        return self.dec(zq), commit, index


@pytest.mark.parametrize("vq_type", ["ema", "base"])
def test_model_trains_through_the_drop_in(vq_type):
    torch.manual_seed(0)
    model = TinyCausalVQAE(vq_type=vq_type).cuda()
    x = torch.sin(torch.linspace(0, 400, 4096, device="cuda"))[None, None].repeat(4, 1, 1)
    x = x + 0.05 * torch.randn_like(x)
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    losses = []
    for step in range(30):
        opt.zero_grad()
        y, commit, index = model(x, update_codebook=True, codebook_n=None if step % 2 else 3)
        loss = torch.nn.functional.mse_loss(y, x) + commit
        loss.backward()
        assert model.enc[0].weight.grad is not None and torch.isfinite(model.enc[0].weight.grad).all()
        if vq_type == "base":
            assert model.quantizer.codebooks.grad is not None
        opt.step()          # "base": the optimiser writes the codebooks in place; the module notices by itself
        losses.append(float(loss))
    assert index.dtype == torch.int64 and index.shape[:2] == (4, 256)
    assert losses[-1] < losses[0]
    # inference path of utils.sound_to_codebooks (utils.py:246-253)
    model.eval()
    with torch.no_grad():
        _, _, idx = model(x)
    oh = torch.nn.functional.one_hot(idx[0], num_classes=model.codebook_size[0])
    assert oh.shape == (256, 4, 256)
    # CausalVQAE.sample (vae.py:329-334): sum of per-stage dequantize
    z = 0
    for i in range(model.quantizer.num_quantizers):
        z = z + model.quantizer.quantizers[i].dequantize(idx[:1, :, i])
    assert z.shape == (1, 256, 128)


def test_c5_shape_plan_model_with_the_references_training_config():
    """BASELINE configs[4] / config/training.yml:13-21: the conv stand-in with the reference's channel and stride plan
    (scripts/c5_harness.py), quantizer built from the yml's vae_args (10 x 512 codes, vq_type "base", use_som, cutoff
    0.1), one training step the way training.py:325-346 does it and the eval path, on the real latent layout."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from c5_harness import SyntheticCausalVQAE
    torch.manual_seed(0)
    model = SyntheticCausalVQAE(num_quantizers=10, codebook_size=512, vq_type="base", vq_cutoff_freq=0.1, use_som=True,
                                som_kernel_type="hard", width=0.25).cuda().train()
    x = torch.randn(4, 1, 72000, device="cuda") * 0.1                 # training.py:310-311: (B=4, 1, 72000) -> L = 150
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    for codebook_n in (10, 3):                                        # training.py:283-294: stage count varies per call
        opt.zero_grad()
        y, commit, index = model(x, update_codebook=True, codebook_n=codebook_n)
        assert y.shape == x.shape and index.shape == (4, 150, codebook_n) and index.dtype == torch.int64
        loss = torch.nn.functional.mse_loss(y, x) + commit
        loss.backward()
        assert torch.isfinite(loss) and model.quantizer.codebooks.grad is not None
        assert torch.isfinite(model.stem.conv.weight.grad).all()
        opt.step()
    assert len(model.quantizer.get_stale_clusters()) == 10            # training.py:435,461
    model.quantizer.update_cutoff(ratio=0.95)                         # training.py:454
    model.eval()
    with torch.no_grad():
        y, commit, index = model(x[:1])
    assert index.shape == (1, 150, 10)
