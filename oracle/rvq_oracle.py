"""CPU restatement of the reference's RVQ bottleneck -- TEST INFRASTRUCTURE, NOT PRODUCT.

PARITY UNPINNED.  The arithmetic of this path is not in the reference tree: the
reference imports it from the third-party package ``som_quantizer``
(LumenPallidium/quantization-maps; ``/root/reference/networks/vae.py:6``), which
is not vendored, not version-pinned (``/root/reference/environment.yml`` does not
list it) and not installable here (no network).  The reference ships no tests and
no golden vectors.  This file therefore restates

* the boundary contract that the reference's own call sites pin
  (``vae.py:245-251`` ctor kwargs, ``vae.py:315-318`` call and return order,
  ``vae.py:333`` ``quantizers[i].dequantize``, ``vae.py:350-351`` ``update_cutoff``,
  ``training.py:183,435,454,461`` attributes, ``utils.py:239-257`` index dtype/shape,
  ``utils.py:212-220`` ``tuple_checker``), and
* the published algorithm named by BASELINE.json's north star (distance to all
  codes -> argmin -> gather -> residual subtract -> EMA count/sum), with the EMA
  constants of the lineage the reference credits (``README.md:26-27``: Jukebox /
  rosinality VQ-VAE: decay 0.99, Laplace smoothing eps 1e-5).

* the codebook maintenance behind the constructor's ``use_som`` / ``som_kernel_type`` / ``vq_cutoff_freq``
  keywords: SOM neighbourhood of the statistics after the paper the reference cites
  (``README.md:10``, arXiv 2302.07950) and dead-code re-seeding of the Jukebox / lucidrains lineage - both
  entirely ASSUMED semantics, with a stated hash for the frame choice so that they are testable.

Every choice not pinned by a call site is marked ASSUMED.
Independent cross-check: ``tests/test_oracle.py`` compares the index chain with
HuggingFace ``EncodecResidualVectorQuantizer.encode`` (an unrelated implementation
of the same distance -> argmax(-dist) -> residual chain).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
``--impl reference``) may import this module.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

EMA_DECAY = 0.99      # ASSUMED (rosinality / Jukebox)
EMA_EPS = 1e-5        # ASSUMED (rosinality Laplace smoothing)


def tuple_checker(item, length):
    """Restates ``/root/reference/networks/utils.py:212-220``: scalar -> list of
    ``length`` copies; tuple/list is length-checked and returned unchanged."""
    if isinstance(item, (int, float, str)):
        item = [item] * length
    elif isinstance(item, (tuple, list)):
        assert len(item) == length, f"Expected tuple of length {length}, got {len(item)}"
    return item


def approximate_square_root(n: int) -> Tuple[int, int]:
    """(h, w) with h*w == n and h the largest divisor <= sqrt(n) (SOM grid shape,
    ``utils.py:244-245,257`` needs height*width == K)."""
    h = int(math.isqrt(n))
    while h > 1 and n % h:
        h -= 1
    return h, n // h


# --------------------------------------------------------------------------
# functional core (the arithmetic the CUDA path is checked against)
# --------------------------------------------------------------------------
def stage_scores(r: torch.Tensor, cb: torch.Tensor) -> torch.Tensor:
    """score[n,k] = ||c_k||^2 - 2 r_n.c_k  (fp32; ||r||^2 dropped, argmin-invariant).
    ASSUMED form."""
    return (cb * cb).sum(1)[None, :] - 2.0 * (r @ cb.t())


def rvq_encode_ref(x2d: torch.Tensor, codebooks: Sequence[torch.Tensor], nq_use: Optional[int] = None,
                   chunk: int = 65536):
    """Encode ``x2d`` (N, d) fp32 through the first ``nq_use`` stages.

    Returns ``(idx int64 (N, nq_use), xq fp32 (N, d), resid fp32 (N, d),
    commit fp64-accumulated python floats per stage [nq_use])``.
    Frames are processed in chunks so the N x K score matrix stays bounded.
    """
    nq = len(codebooks) if nq_use is None else int(nq_use)
    x2d = x2d.float()
    N, d = x2d.shape
    idx = torch.empty(N, nq, dtype=torch.int64)
    xq = torch.empty_like(x2d)
    resid = torch.empty_like(x2d)
    sq = [0.0] * nq
    for s in range(0, N, chunk):
        r = x2d[s:s + chunk].clone()
        acc = torch.zeros_like(r)
        for q in range(nq):
            cb = codebooks[q].float()
            i = stage_scores(r, cb).argmin(dim=1)      # lowest index wins exact ties (torch CPU)
            z = cb[i]
            r = r - z
            acc = acc + z
            sq[q] += float((r.double() ** 2).sum())
            idx[s:s + chunk, q] = i
        xq[s:s + chunk] = acc
        resid[s:s + chunk] = r
    commit = [v / (N * d) for v in sq]
    return idx, xq, resid, commit


def ema_stats_ref(r_in: torch.Tensor, idx_q: torch.Tensor, K: int):
    """Per-stage statistics: count[k] = #{n: idx=k}; sum[k] = sum of stage-input residuals."""
    cnt = torch.bincount(idx_q, minlength=K).float()
    sm = torch.zeros(K, r_in.shape[1], dtype=torch.float32).index_add_(0, idx_q, r_in.float())
    return cnt, sm


def ema_finalize_ref(cb, ema_count, ema_sum, cnt, sm, decay=EMA_DECAY, eps=EMA_EPS):
    """EMA + Laplace-smoothed codebook refresh (ASSUMED constants; rosinality form)."""
    K = cb.shape[0]
    ema_count = decay * ema_count + (1.0 - decay) * cnt
    ema_sum = decay * ema_sum + (1.0 - decay) * sm
    n_tot = ema_count.sum()
    smoothed = (ema_count + eps) / (n_tot + K * eps) * n_tot
    cb = ema_sum / smoothed[:, None]
    return cb, ema_count, ema_sum


def stage_residuals_from_indices(x2d: torch.Tensor, codebooks, idx: torch.Tensor) -> List[torch.Tensor]:
    """Residual entering each stage when the chain follows ``idx`` (teacher forcing)."""
    r = x2d.float().clone()
    out = []
    for q in range(idx.shape[1]):
        out.append(r)
        r = r - codebooks[q].float()[idx[:, q]]
    out.append(r)
    return out


def adjudicate_indices(x2d, codebooks, idx_test: torch.Tensor, eps_scale: float = 8.0, chunk: int = 32768):
    """Teacher-forced index parity check used by the GPU tests.

    For every stage the oracle argmin is recomputed on the residual obtained by
    following ``idx_test``'s own prefix.  A disagreement is *legitimate* only when
    the two candidate scores differ, in fp64, by at most
    ``eps_tie = eps_scale * d * 2**-24 * (||r||*||c||_max + ||c||_max^2)`` (fp32 dot-product
    reordering noise).  Returns ``dict(n_mismatch, n_illegal, max_gap_ratio)``.
    """
    x2d = x2d.float()
    N, d = x2d.shape
    nq = idx_test.shape[1]
    n_mis = 0
    n_bad = 0
    worst = 0.0
    for s in range(0, N, chunk):
        r = x2d[s:s + chunk].clone()
        it = idx_test[s:s + chunk]
        for q in range(nq):
            cb = codebooks[q].float()
            io = stage_scores(r, cb).argmin(dim=1)
            dif = (io != it[:, q]).nonzero().flatten()
            if dif.numel():
                n_mis += int(dif.numel())
                rr = r[dif].double()
                ca = cb[io[dif]].double()
                cbk = cb[it[dif, q]].double()
                sa = (ca * ca).sum(1) - 2.0 * (rr * ca).sum(1)
                sb = (cbk * cbk).sum(1) - 2.0 * (rr * cbk).sum(1)
                cmax = cb.double().norm(dim=1).max()
                eps_tie = eps_scale * d * 2.0 ** -24 * (rr.norm(dim=1) * cmax + cmax * cmax)
                ratio = ((sb - sa).abs() / eps_tie)
                worst = max(worst, float(ratio.max()))
                n_bad += int((ratio > 1.0).sum())
            r = r - cb[it[:, q]]
    return dict(n_mismatch=n_mis, n_illegal=n_bad, max_gap_ratio=worst)


# --------------------------------------------------------------------------
# codebook maintenance around the EMA update -- every choice ASSUMED
# --------------------------------------------------------------------------
SOM_SHRINK = 0.1      # ASSUMED: neighbourhood width sigma_t = 1 / (1 + SOM_SHRINK * t), t = update steps so far
SOM_MAX_RADIUS = 4


def som_weights(kernel_type: str, t: int, shrink: float = SOM_SHRINK):
    """(radius, float32 weights [(2r+1), (2r+1)]) of the SOM neighbourhood at update step ``t``.

    Restates the neighbourhood functions of Irie et al., "Self-organising neural discrete representation
    learning a la Kohonen" (arXiv 2302.07950, the SOM paper ``/root/reference/README.md:10`` cites), with the
    time-shrinking width sigma_t = 1 / (1 + shrink * t):
      "hard"      the code itself 1, its four grid neighbours (Manhattan distance 1) sigma_t, everything else 0;
      "gaussian"  exp(-(dy^2 + dx^2) / (2 sigma_t^2)) inside the window of radius min(4, max(1, ceil(3 sigma_t))).
    """
    import numpy as np
    sigma = 1.0 / (1.0 + float(shrink) * float(t))
    if kernel_type == "hard":
        w = np.zeros((3, 3), dtype=np.float32)
        w[1, 1] = 1.0
        w[0, 1] = w[2, 1] = w[1, 0] = w[1, 2] = np.float32(sigma)
        return 1, w
    if kernel_type == "gaussian":
        r = min(SOM_MAX_RADIUS, max(1, int(math.ceil(3.0 * sigma))))
        w = np.zeros((2 * r + 1, 2 * r + 1), dtype=np.float32)
        for dy in range(-r, r + 1):
            for dx in range(-r, r + 1):
                w[dy + r, dx + r] = np.float32(math.exp(-(dy * dy + dx * dx) / (2.0 * sigma * sigma)))
        return r, w
    raise ValueError(f"som_kernel_type must be 'hard' or 'gaussian', got {kernel_type!r}")


def som_spread_ref(sm, cnt, height: int, width: int, radius: int, weights):
    """Neighbourhood spreading of one stage's statistics on its ``height x width`` map (row-major codes):
    out[(y, x)] = sum_{dy, dx} w[dy, dx] * in[(y + dy, x + dx)], neighbours outside the grid skipped, zero weights
    skipped, terms added in row-major (dy, dx) order with separate float32 multiply and add.  Codes beyond
    height * width pass through.  ``sm`` (K, d), ``cnt`` (K,) -> float32 numpy arrays of the same shapes."""
    import numpy as np
    sm = np.asarray(sm, dtype=np.float32)
    cnt = np.asarray(cnt, dtype=np.float32)
    K, d = sm.shape
    n = height * width
    full = np.concatenate([sm, cnt[:, None]], axis=1)
    grid = full[:n].reshape(height, width, d + 1)
    acc = np.zeros_like(grid)
    for dy in range(-radius, radius + 1):
        for dx in range(-radius, radius + 1):
            w = np.float32(weights[dy + radius][dx + radius])
            if w == 0:
                continue
            y0, y1 = max(0, -dy), min(height, height - dy)
            x0, x1 = max(0, -dx), min(width, width - dx)
            if y0 >= y1 or x0 >= x1:
                continue
            acc[y0:y1, x0:x1] = acc[y0:y1, x0:x1] + w * grid[y0 + dy:y1 + dy, x0 + dx:x1 + dx]
    out = full.copy()
    out[:n] = acc.reshape(n, d + 1)
    return out[:, :d].copy(), out[:, d].copy()


_M64 = (1 << 64) - 1


def reseed_frame_ref(seed: int, q: int, K: int, k: int, frames_total: int) -> int:
    """Global frame whose stage-q residual re-seeds code (q, k): splitmix64 finaliser of
    ``seed + (q K + k + 1) * 0x9E3779B97F4A7C15`` (64-bit wrap-around), modulo ``frames_total``."""
    z = (seed + (q * K + k + 1) * 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z % frames_total


def reseed_vectors_ref(r_q, q: int, K: int, seed: int, frame_offset: int = 0, frames_total: Optional[int] = None):
    """Replacement vector of every code of stage ``q``: the stage-q residual ``r_q[n]`` of the frame
    ``n = reseed_frame_ref(seed, q, K, k, frames_total) - frame_offset``, zeros when another rank owns that frame
    (the caller sums over ranks).  (K, d) float32."""
    N, d = r_q.shape
    total = N if frames_total is None else int(frames_total)
    rep = torch.zeros(K, d, dtype=torch.float32)
    for k in range(K):
        n = reseed_frame_ref(seed, q, K, k, total) - frame_offset
        if 0 <= n < N:
            rep[k] = r_q[n].float()
    return rep


def reseed_apply_ref(cb, ema_count, ema_sum, rep, cutoff: float, reset_count: float):
    """Dead-code re-seeding of one stage after its EMA refresh (lineage: Jukebox ``restore_k`` / lucidrains
    ``expire_codes_``; ASSUMED): every code with ``ema_count < cutoff`` takes ``rep[k]``,
    ``ema_sum = rep[k] * reset_count``, ``ema_count = reset_count``.  Returns (cb, ema_count, ema_sum, n_replaced)."""
    stale = ema_count < cutoff
    rc = torch.tensor(reset_count, dtype=torch.float32)
    cb = torch.where(stale[:, None], rep, cb)
    ema_sum = torch.where(stale[:, None], rep * rc, ema_sum)
    ema_count = torch.where(stale, rc, ema_count)
    return cb, ema_count, ema_sum, int(stale.sum())


class _SomGradSpread(torch.autograd.Function):
    """Identity on a stage's codebook whose BACKWARD spreads the gradient over each code's map neighbourhood
    (``som_spread_ref``): the SOM coupling for gradient-trained (``quantizer_class="base"``) codebooks - the
    neighbours of a winning code are pulled along with it.  ASSUMED (the reference's config/training.yml:15-21 asks
    for ``vq_type: "base"`` with ``use_som: True``; the package that defines it is absent)."""

    @staticmethod
    def forward(ctx, cb, height, width, radius, weights):
        ctx.geom = (height, width, radius, weights)
        return cb.view_as(cb)

    @staticmethod
    def backward(ctx, g):
        import numpy as np
        h, w, radius, weights = ctx.geom
        gs, _ = som_spread_ref(g.detach().cpu().numpy(), np.zeros(g.shape[0], dtype=np.float32), h, w, radius, weights)
        return torch.from_numpy(gs).to(g.device), None, None, None, None


# --------------------------------------------------------------------------
# nn.Module with the reference's call-site contract
# --------------------------------------------------------------------------
class _SOMGrid:
    def __init__(self, K):
        self.height, self.width = approximate_square_root(K)


class _StageRef:
    """``quantizers[i]`` view: ``.dequantize`` (``vae.py:333``), ``.som`` (``utils.py:244-245``)."""

    def __init__(self, parent, q):
        self._p, self._q = parent, q
        if parent.use_som:
            self.som = _SOMGrid(parent.codebook_sizes[q])

    @property
    def codebook(self):
        return self._p.codebooks[self._q][: self._p.codebook_sizes[self._q]]

    def dequantize(self, idx):
        return self.codebook[idx]


class ResidualQuantizerRef(nn.Module):
    """Oracle ``ResidualQuantizer`` (ctor kwargs: ``vae.py:245-251``; call: ``vae.py:315-318``)."""

    def __init__(self, num_quantizers, dim, quantizer_class="ema", codebook_sizes=1024,
                 vq_cutoff_freq=1, use_som=True, som_kernel_type="hard",
                 decay=EMA_DECAY, eps=EMA_EPS, commitment_weight=1.0, som_shrink=SOM_SHRINK, reseed_seed=0):
        super().__init__()
        self.som_shrink, self.reseed_seed = float(som_shrink), int(reseed_seed)
        self.num_quantizers = int(num_quantizers)
        self.dim = int(dim)
        self.quantizer_class = quantizer_class
        self.codebook_sizes = [int(k) for k in tuple_checker(codebook_sizes, self.num_quantizers)]
        self.vq_cutoff_freq = float(vq_cutoff_freq)
        self.use_som = bool(use_som)
        self.som_kernel_type = som_kernel_type
        self.decay, self.eps, self.commitment_weight = float(decay), float(eps), float(commitment_weight)
        Kmax = max(self.codebook_sizes)
        cb = torch.randn(self.num_quantizers, Kmax, self.dim)          # ASSUMED init
        if quantizer_class == "base":
            self.codebooks = nn.Parameter(cb)
        else:
            self.register_buffer("codebooks", cb)
        self.register_buffer("ema_count", torch.ones(self.num_quantizers, Kmax))
        self.register_buffer("ema_sum", cb.detach().clone())
        self.register_buffer("update_steps", torch.zeros((), dtype=torch.int64))
        self.quantizers = [_StageRef(self, q) for q in range(self.num_quantizers)]
        self.n_replaced = [0] * self.num_quantizers          # codes re-seeded by the last update

    def step_seed(self):
        return (self.reseed_seed + int(self.update_steps) * 0xD1B54A32D192ED03) & _M64

    def forward(self, x, n=None, update_codebook=False, prioritize_early=False):
        if prioritize_early:
            raise NotImplementedError("prioritize_early=True: semantics unknown (never used by the reference)")
        nq = self.num_quantizers if n is None else int(n)
        shp = x.shape
        r = x.reshape(-1, self.dim).float()
        N = r.shape[0]
        xq = torch.zeros_like(r)
        commit = r.new_zeros(())
        idxs = []
        for q in range(nq):
            K = self.codebook_sizes[q]
            cb = self.codebooks[q, :K]
            if self.quantizer_class == "base" and self.use_som and cb.requires_grad:
                radius, w = som_weights(self.som_kernel_type, int(self.update_steps), self.som_shrink)
                cb = _SomGradSpread.apply(cb, *approximate_square_root(K), radius, w)
            with torch.no_grad():
                i = stage_scores(r.detach(), cb.detach()).argmin(dim=1)
            z = cb[i]
            commit = commit + self.commitment_weight * ((r - z.detach()) ** 2).mean()
            if self.quantizer_class == "base":
                commit = commit + ((r.detach() - z) ** 2).mean()       # codebook loss (VQ-VAE), ASSUMED
            if update_codebook and self.training and self.quantizer_class == "ema":
                with torch.no_grad():
                    cnt, sm = ema_stats_ref(r.detach(), i, K)
                    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
                    rank = torch.distributed.get_rank() if dist_on else 0
                    world = torch.distributed.get_world_size() if dist_on else 1
                    rep = None
                    if self.vq_cutoff_freq > 0:        # replacement vectors for stale codes (equal shards ASSUMED)
                        rep = reseed_vectors_ref(r.detach(), q, max(self.codebook_sizes), self.step_seed(),
                                                 rank * N, N * world)[:K]
                    if dist_on:
                        torch.distributed.all_reduce(cnt)
                        torch.distributed.all_reduce(sm)
                        if rep is not None:
                            torch.distributed.all_reduce(rep)
                    if self.use_som:                   # SOM neighbourhood of the (all-reduced) statistics
                        radius, w = som_weights(self.som_kernel_type, int(self.update_steps), self.som_shrink)
                        h, wd = approximate_square_root(K)
                        sm_, cnt_ = som_spread_ref(sm.numpy(), cnt.numpy(), h, wd, radius, w)
                        sm, cnt = torch.from_numpy(sm_), torch.from_numpy(cnt_)
                    ncb, nc, ns = ema_finalize_ref(cb, self.ema_count[q, :K], self.ema_sum[q, :K], cnt, sm,
                                                   self.decay, self.eps)
                    if rep is not None:
                        ncb, nc, ns, nrep = reseed_apply_ref(ncb, nc, ns, rep, self.vq_cutoff_freq,
                                                             self.vq_cutoff_freq)
                        self.n_replaced[q] = nrep
                    self._pending = getattr(self, "_pending", [])
                    self._pending.append((q, K, ncb, nc, ns))
            if update_codebook and self.training and self.quantizer_class == "base":
                # gradient-trained codebooks (training.py:305-308 passes update_codebook=True for them too): only the
                # usage counts are averaged and stale codes re-seeded; ASSUMED
                with torch.no_grad():
                    cnt, _ = ema_stats_ref(r.detach(), i, K)
                    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
                    rank = torch.distributed.get_rank() if dist_on else 0
                    world = torch.distributed.get_world_size() if dist_on else 1
                    rep = None
                    if self.vq_cutoff_freq > 0:
                        rep = reseed_vectors_ref(r.detach(), q, max(self.codebook_sizes), self.step_seed(),
                                                 rank * N, N * world)[:K]
                    if dist_on:
                        torch.distributed.all_reduce(cnt)
                        if rep is not None:
                            torch.distributed.all_reduce(rep)
                    nc = self.decay * self.ema_count[q, :K] + (1.0 - self.decay) * cnt
                    ncb, ns = self.codebooks.detach()[q, :K].clone(), self.ema_sum[q, :K].clone()
                    if rep is not None:
                        ncb, nc, ns, nrep = reseed_apply_ref(ncb, nc, ns, rep, self.vq_cutoff_freq,
                                                             self.vq_cutoff_freq)
                        self.n_replaced[q] = nrep
                    self._pending = getattr(self, "_pending", [])
                    self._pending.append((q, K, ncb, nc, ns))
            xq = xq + z.detach()
            r = r - z.detach()
            idxs.append(i)
        for (q, K, ncb, nc, ns) in getattr(self, "_pending", []):       # takes effect from the NEXT call
            self.codebooks.data[q, :K] = ncb
            self.ema_count[q, :K] = nc
            self.ema_sum[q, :K] = ns
        if getattr(self, "_pending", []):
            self.update_steps += 1
        self._pending = []
        xf = x.reshape(-1, self.dim)
        x_quantized = (xf + (xq - xf).detach()).reshape(shp)           # straight-through
        index = torch.stack(idxs, dim=-1).reshape(*shp[:-1], nq)
        return x_quantized, index, commit

    def get_stale_clusters(self):
        """Per stage, the number of codes whose EMA count (hits per call) is below ``vq_cutoff_freq`` (ASSUMED:
        the dead-code threshold of the Jukebox / lucidrains lineage, in count units)."""
        return [int((self.ema_count[q, :self.codebook_sizes[q]] < self.vq_cutoff_freq).sum())
                for q in range(self.num_quantizers)]

    def update_cutoff(self, new_cutoff=None, ratio=None):
        if new_cutoff is not None:
            self.vq_cutoff_freq = float(new_cutoff)
        if ratio is not None:
            self.vq_cutoff_freq *= float(ratio)


# ----------------------------------------------------------------------------------------- wire format
def pack_indices_ref(idx, bits):
    """CPU restatement of the code wire format (include/rvq_sm100a.h: rvq_pack_indices): the n codes of a frame
    LSB-first, `bits` bits each, frames byte-aligned.  bits per frame = n * log2 K as in the reference's
    bitrate_calculator (/root/reference/networks/utils.py:137-147).  idx: int64 array-like [N, n] -> uint8 [N, ceil(n*bits/8)]."""
    import numpy as np
    a = np.asarray(idx, dtype=np.uint64)
    N, n = a.shape
    bpf = (n * bits + 7) // 8
    out = np.zeros((N, bpf), dtype=np.uint8)
    for f in range(N):
        acc = 0
        for q in range(n):
            acc |= (int(a[f, q]) & ((1 << bits) - 1)) << (q * bits)
        out[f] = np.frombuffer(acc.to_bytes(bpf, "little"), dtype=np.uint8)
    return out


def unpack_indices_ref(packed, n, bits):
    import numpy as np
    p_ = np.asarray(packed, dtype=np.uint8)
    out = np.zeros((p_.shape[0], n), dtype=np.int64)
    for f in range(p_.shape[0]):
        acc = int.from_bytes(p_[f].tobytes(), "little")
        for q in range(n):
            out[f, q] = (acc >> (q * bits)) & ((1 << bits) - 1)
    return out
