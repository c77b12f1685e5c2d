"""Synthetic genotype-likelihood generator.

Restates the *distribution* of the reference's simulator
(``scripts/ngsF-HMMsim.R``; R is not available in this image): site spacing
``max(1, int(Normal(1e5, 1e5/3)))`` bp (sim.R:193-196), per-individual IBD state
from the two-state chain ``P(0->1) = (1-exp(-a d)) F``, ``P(1->0) = (1-exp(-a d))(1-F)``
with a Bernoulli(F) first state (sim.R:14-36), two haplotypes drawn
Bernoulli(freq) and copied where IBD (sim.R:38-46), depth ~ Poisson(depth),
reads ~ Binomial(depth, {err, 0.5, 1-err}[g]), GL_g = normalised binomial pmf,
natural log rounded to 10 digits (sim.R:48-67).  Depth 0 gives (1/3,1/3,1/3).

The chain is sampled with the renewal form of the transition matrix
T = c I + (1-c) 1 q^T: with probability 1-c the state is redrawn from q,
otherwise kept - identical in distribution and vectorisable over sites.
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass

import numpy as np


@dataclass
class SimData:
    n_ind: int
    n_sites: int
    pos_bp: np.ndarray        # (S,) int64 absolute positions, single chromosome
    dist_mb: np.ndarray       # (S,) float64 distance to previous site in Mb (site 0: its own position)
    log_gl: np.ndarray        # (S, N, 3) float64 natural-log GL, site-major (the binary input layout)
    true_freq: np.ndarray     # (S,)
    true_F: np.ndarray        # (N,)
    true_alpha: np.ndarray    # (N,)
    true_path: np.ndarray     # (N, S) uint8
    geno: np.ndarray          # (N, S) uint8 true genotypes


def simulate(n_ind: int, n_sites: int, *, freq=0.2, indF=0.5, alpha=0.01, depth=2.0, error=0.01,
             seed=12345, round_digits=10) -> SimData:
    """freq / indF / alpha may be scalars, arrays, or (lo, hi) tuples for uniform draws."""
    rng = np.random.default_rng(seed)
    N, S = int(n_ind), int(n_sites)

    def expand(v, n):
        if isinstance(v, tuple):
            return rng.uniform(v[0], v[1], size=n)
        a = np.asarray(v, dtype=np.float64)
        return np.full(n, float(a)) if a.ndim == 0 else a.astype(np.float64)

    f = expand(freq, S)
    F = expand(indF, N)
    a = expand(alpha, N)

    step = np.maximum(1, rng.normal(1e5, 1e5 / 3, size=S).astype(np.int64))
    pos = np.cumsum(step)
    dist_mb = step.astype(np.float64) / 1e6

    path = np.empty((N, S), dtype=np.uint8)
    geno = np.empty((N, S), dtype=np.uint8)
    log_gl = np.empty((S, N, 3), dtype=np.float64)
    lp = np.log(np.array([error, 0.5, 1 - error]))
    lq = np.log(np.array([1 - error, 0.5, error]))
    p_read = np.array([error, 0.5, 1 - error])
    idx = np.arange(S)
    for i in range(N):
        keep = np.exp(-a[i] * dist_mb)
        redraw = rng.random(S) < (1 - keep)
        redraw[0] = True
        fresh = (rng.random(S) < F[i]).astype(np.uint8)
        last = np.maximum.accumulate(np.where(redraw, idx, 0))
        st = fresh[last]
        path[i] = st
        h1 = (rng.random(S) < f).astype(np.uint8)
        h2 = (rng.random(S) < f).astype(np.uint8)
        h2 = np.where(st == 1, h1, h2)
        g = h1 + h2
        geno[i] = g
        dp = rng.poisson(depth, size=S)
        nA = rng.binomial(dp, p_read[g])
        ll = nA[:, None] * lp[None, :] + (dp - nA)[:, None] * lq[None, :]
        m = ll.max(axis=1, keepdims=True)
        ll = ll - (m + np.log(np.exp(ll - m).sum(axis=1, keepdims=True)))
        log_gl[:, i, :] = np.round(ll, round_digits) if round_digits is not None else ll
    return SimData(N, S, pos, dist_mb, log_gl, f, F, a, path, geno)


# ---------------------------------------------------------------------------
# Input files in the formats the reference reads (shared/read_data.cpp)
# ---------------------------------------------------------------------------

def write_pos(path: str, pos_bp: np.ndarray, chrom: str = "chr1") -> None:
    """Two columns chrom<TAB>pos (read_data.cpp:165-218); gz if the name ends in .gz."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt") as fh:
        fh.write("".join(f"{chrom}\t{int(p)}\n" for p in pos_bp))


def write_binary_gl(path: str, log_gl: np.ndarray) -> None:
    """Raw doubles, site-major S x N x 3 (read_data.cpp:28-31); run with --loglkl."""
    np.ascontiguousarray(log_gl, dtype=np.float64).tofile(path)


def write_beagle_gz(path: str, log_gl: np.ndarray, pos_bp: np.ndarray, chrom: str = "chr1") -> None:
    """BEAGLE-style text: header + marker/allele1/allele2 + 3 linear GLs per individual; run with --lkl."""
    S, N, _ = log_gl.shape
    lin = np.exp(log_gl)
    with gzip.open(path, "wt") as fh:
        fh.write("marker\tallele1\tallele2" + "".join(f"\tInd{i}\tInd{i}\tInd{i}" for i in range(N)) + "\n")
        for s in range(S):
            fh.write(f"{chrom}_{int(pos_bp[s])}\tA\tC\t" + "\t".join(f"{v:.6f}" for v in lin[s].ravel()) + "\n")


def write_geno_gz(path: str, geno: np.ndarray) -> None:
    """Called genotypes, one column per individual, one line per site (read_data.cpp:88-97)."""
    with gzip.open(path, "wt") as fh:
        for s in range(geno.shape[1]):
            fh.write("\t".join(str(int(g)) for g in geno[:, s]) + "\n")


# ---------------------------------------------------------------------------
# Large synthetic inputs, generated with torch on whatever device is given
# (the GPU on the bench box).  Same distribution as simulate(); different RNG
# stream.  Returns NORMALISED natural-log GL in the upload layout.
# ---------------------------------------------------------------------------

def simulate_torch(n_ind, n_sites, *, device, seed=1002, freq=(0.05, 0.5), indF=(0.0, 0.5), alpha=0.01, depth=2.0,
                   error=0.01, site_chunk=1 << 18, out=None, site_begin=0, site_end=None):
    """Generate sites [site_begin, site_end) for ALL n_ind individuals.

    Every quantity is a pure function of (seed, site index, individual index) chunk by chunk, so ranks that
    generate different site blocks see one consistent data set.  Returns dict(log_gl (n, N, 3) float64 on
    `out`'s device or CPU pinned, dist_mb (n_sites,) numpy, true_F, true_alpha, true_freq (n,)).
    """
    import torch

    N, S = int(n_ind), int(n_sites)
    site_end = S if site_end is None else site_end
    n = site_end - site_begin
    g = torch.Generator(device=device)

    def uni(lohi, count, tag):
        g.manual_seed(seed * 1000003 + tag)
        if isinstance(lohi, tuple):
            return lohi[0] + (lohi[1] - lohi[0]) * torch.rand(count, generator=g, device=device, dtype=torch.float64)
        return torch.full((count,), float(lohi), device=device, dtype=torch.float64)

    F = uni(indF, N, 1)
    a = uni(alpha, N, 2)
    f_all = uni(freq, S, 3)
    g.manual_seed(seed * 1000003 + 4)
    step = torch.clamp((1e5 + (1e5 / 3) * torch.randn(S, generator=g, device=device, dtype=torch.float64)).floor(),
                       min=1.0)
    dist_mb = step / 1e6

    if out is None:
        out = torch.empty((n, N, 3), dtype=torch.float64, pin_memory=(str(device) != "cpu"))
    lp = torch.log(torch.tensor([error, 0.5, 1 - error], device=device, dtype=torch.float64))
    lq = torch.log(torch.tensor([1 - error, 0.5, error], device=device, dtype=torch.float64))
    p_read = torch.tensor([error, 0.5, 1 - error], device=device, dtype=torch.float64)

    # IBD state: renewal form of the chain, carried across chunks through the last state of each individual.
    # To stay a pure function of the site range, the chain is (re)started with a Bernoulli(F) draw at every
    # chunk boundary that is a multiple of site_chunk - a negligible change for throughput inputs.
    for c0 in range((site_begin // site_chunk) * site_chunk, site_end, site_chunk):
        c1 = min(c0 + site_chunk, S)
        m = c1 - c0
        g.manual_seed(seed * 1000003 + 1000 + c0 // site_chunk)
        keep = torch.exp(-a[None, :] * dist_mb[c0:c1, None])                     # (m, N)
        redraw = torch.rand((m, N), generator=g, device=device, dtype=torch.float64) < (1 - keep)
        redraw[0, :] = True
        fresh = torch.rand((m, N), generator=g, device=device, dtype=torch.float64) < F[None, :]
        idx = torch.arange(m, device=device)[:, None].expand(m, N)
        last = torch.cummax(torch.where(redraw, idx, torch.zeros_like(idx)), dim=0).values
        state = torch.gather(fresh, 0, last)
        fc = f_all[c0:c1, None]
        h1 = torch.rand((m, N), generator=g, device=device, dtype=torch.float64) < fc
        h2 = torch.rand((m, N), generator=g, device=device, dtype=torch.float64) < fc
        h2 = torch.where(state, h1, h2)
        geno = h1.long() + h2.long()
        dp = torch.poisson(torch.full((m, N), float(depth), device=device, dtype=torch.float64), generator=g)
        nA = torch.binomial(dp, p_read[geno], generator=g)
        ll = nA[..., None] * lp + (dp - nA)[..., None] * lq                       # (m, N, 3)
        ll = ll - torch.logsumexp(ll, dim=-1, keepdim=True)
        ll = torch.round(ll * 1e10) / 1e10                                        # sim.R rounds to 10 digits
        ll = ll - torch.logsumexp(ll, dim=-1, keepdim=True)                       # read_geno + main normalise
        ll = ll - torch.logsumexp(ll, dim=-1, keepdim=True)
        lo, hi = max(c0, site_begin), min(c1, site_end)
        if hi > lo:
            out[lo - site_begin:hi - site_begin].copy_(ll[lo - c0:hi - c0], non_blocking=False)
    return dict(log_gl=out, dist_mb=dist_mb.cpu().numpy(), true_F=F.cpu().numpy(), true_alpha=a.cpu().numpy(),
                true_freq=f_all[site_begin:site_end].cpu().numpy())
