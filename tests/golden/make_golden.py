"""Generates tests/golden/*.npz -- run in the build container (where /root/reference exists).

Two kinds of vectors:
  rvq_small.npz      full tensors of a small seeded problem (x, codebooks -> idx, xq, residual, commit,
                     EMA counts/sums/codebooks after one update), produced by oracle/rvq_oracle.py.
  c1_reference.npz   BASELINE configs[0]: the UNMODIFIED reference model (/root/reference/networks/vae.py,
                     CausalVQAE(**config/training.yml vae_args)) run on the bundled networks/om.wav on CPU with the
                     oracle standing in for the absent third-party `som_quantizer`: the latent frames the reference's
                     encoder hands to the quantizer (vae.py:313-318), the quantizer's codebooks (seeded), and the
                     oracle's indices / commit loss for them.  Latents are stored as fp16-exact values so that the
                     fixture stays small and the test inputs are bit-identical everywhere.

PARITY UNPINNED: the reference ships no golden vectors for this path; these pin the ORACLE
(and the layout/shape contract of the reference's call site), not the upstream package.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import rvq_oracle as O  # noqa: E402


def small():
    torch.manual_seed(20261018)
    nq, K, d, N = 4, 64, 64, 200
    cbs = torch.randn(nq, K, d) * torch.tensor([0.7 ** q for q in range(nq)])[:, None, None]
    x = torch.randn(N, d)
    idx, xq, r, commit = O.rvq_encode_ref(x, list(cbs))
    res = O.stage_residuals_from_indices(x, list(cbs), idx)
    cnts, sums, ncbs = [], [], []
    for q in range(nq):
        cnt, sm = O.ema_stats_ref(res[q], idx[:, q], K)
        ncb, nc, ns = O.ema_finalize_ref(cbs[q], torch.ones(K), cbs[q].clone(), cnt, sm)
        cnts.append(cnt)
        sums.append(sm)
        ncbs.append(ncb)
    np.savez_compressed(os.path.join(HERE, "rvq_small.npz"), x=x.numpy(), codebooks=cbs.numpy(), idx=idx.numpy(),
                        xq=xq.numpy(), resid=r.numpy(), commit=np.array(commit), cnt=torch.stack(cnts).numpy(),
                        sum=torch.stack(sums).numpy(), new_codebooks=torch.stack(ncbs).numpy())
    print("rvq_small.npz", idx.shape, commit)


def maint():
    """codebook_maint.npz: SOM neighbourhood spreading and stale-code re-seeding on a small seeded problem
    (semantics ASSUMED, see oracle/rvq_oracle.py)."""
    torch.manual_seed(20261019)
    K, d, N, q, seed = 24, 64, 96, 2, 0x1234ABCD
    h, w = O.approximate_square_root(K)
    sm = torch.randn(K, d)
    cnt = torch.randint(0, 9, (K,)).float()
    out = {}
    for name, t in (("hard", 3), ("gaussian", 0), ("gaussian", 40)):
        radius, wts = O.som_weights(name, t)
        osm, ocnt = O.som_spread_ref(sm.numpy(), cnt.numpy(), h, w, radius, wts)
        out[f"som_{name}_{t}_sum"] = osm
        out[f"som_{name}_{t}_cnt"] = ocnt
        out[f"som_{name}_{t}_w"] = np.asarray(wts)
    r_q = torch.randn(N, d)
    rep = O.reseed_vectors_ref(r_q, q, K, seed)
    rep_rank1 = O.reseed_vectors_ref(r_q[N // 2:], q, K, seed, frame_offset=N // 2, frames_total=N)
    frames = np.array([O.reseed_frame_ref(seed, q, K, k, N) for k in range(K)])
    cb = torch.randn(K, d)
    ema_count = torch.rand(K) * 2.0
    ema_sum = torch.randn(K, d)
    ncb, nc, ns, n_rep = O.reseed_apply_ref(cb, ema_count, ema_sum, rep, 1.0, 1.0)
    np.savez_compressed(os.path.join(HERE, "codebook_maint.npz"), K=K, d=d, N=N, q=q, seed=seed, grid=np.array([h, w]),
                        sum=sm.numpy(), cnt=cnt.numpy(), r_q=r_q.numpy(), rep=rep.numpy(), rep_rank1=rep_rank1.numpy(),
                        frames=frames, cb=cb.numpy(), ema_count=ema_count.numpy(), ema_sum=ema_sum.numpy(),
                        new_cb=ncb.numpy(), new_count=nc.numpy(), new_sum=ns.numpy(), n_replaced=n_rep, **out)
    print("codebook_maint.npz", frames[:6], n_rep)


def c1():
    import yaml
    from scipy.io import wavfile
    ref = "/root/reference"
    sys.path.insert(0, os.path.join(ref, "networks"))
    # the reference imports matplotlib (absent here) and som_quantizer (third-party, absent): stand-ins
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sq = types.ModuleType("som_quantizer")
    sq.ResidualQuantizer = O.ResidualQuantizerRef
    sq.tuple_checker = O.tuple_checker
    sys.modules["som_quantizer"] = sq
    import vae  # the unmodified reference file
    cfg = yaml.safe_load(open(os.path.join(ref, "config", "training.yml")))
    torch.manual_seed(0)
    model = vae.CausalVQAE(**cfg["vae_args"]).eval()
    sr, wav = wavfile.read(os.path.join(ref, "networks", "om.wav"))
    om = torch.from_numpy(wav.astype(np.float32)).t().mean(0, keepdim=True).unsqueeze(0)[:, :, :65280]
    captured = {}
    q = model.quantizer
    orig = q.forward

    def spy(x, n=None, update_codebook=False, prioritize_early=False):
        captured["x_shape"] = tuple(x.shape)
        captured["x_stride"] = tuple(x.stride())
        captured["x"] = x.detach().clone()
        out = orig(x, n, update_codebook=update_codebook, prioritize_early=prioritize_early)
        captured["out"] = out
        return out

    q.forward = spy
    with torch.no_grad():
        y, commit, index = model(om)
    x = captured["x"]
    x16 = x.half().float()                      # fp16-exact copy keeps the fixture small (139 KB)
    torch.manual_seed(4321)
    nq, K, d = q.num_quantizers, q.codebook_sizes[0], q.dim
    cbs = torch.randn(nq, K, d) * x16.std() * torch.tensor([0.8 ** i for i in range(nq)])[:, None, None]
    idx, xq, r, commits = O.rvq_encode_ref(x16.reshape(-1, d), list(cbs))
    np.savez_compressed(os.path.join(HERE, "c1_reference.npz"), x_fp16=x16.half().numpy(),
                        x_shape=np.array(captured["x_shape"]), x_stride=np.array(captured["x_stride"]),
                        codebook_seed=np.array(4321), codebook_scale=np.array(float(x16.std())),
                        idx=idx.numpy().astype(np.int16), commit=np.array(commits), xq_checksum=np.array(float(xq.double().sum())),
                        ctor=np.array([nq, K, d]), ref_index_shape=np.array(tuple(index.shape)),
                        wav_samples=np.array(om.shape[-1]),
                        codebook_checksum=np.array(float(cbs.double().abs().sum())))
    print("c1_reference.npz", captured["x_shape"], captured["x_stride"], "index", tuple(index.shape), "commit", commits[:3])


if __name__ == "__main__":
    small()
    maint()
    c1()
