"""GPU parity: every hot-path kernel, called through the C ABI, against the CPU oracle.

Tolerances are the north-star's: log-likelihood 1e-9 relative, posterior 1e-8
absolute (sites whose unclamped value sits within 1e-8 of a clamp threshold are
exempt - threshold flips), frequencies / F / alpha 1e-6,
Viterbi identical except at near-ties.
"""
import numpy as np
import pytest

import ngsf_hmm_b200 as nfh
from ngsf_hmm_b200 import sim

pytestmark = pytest.mark.gpu

LKL_RTOL = 1e-9
POST_ATOL = 1e-8
FREQ_ATOL = 1e-6


def _setup(N, S, seed, freq0=0.1, F0=0.1, a0=0.2, **simkw):
    d = sim.simulate(N, S, seed=seed, **simkw)
    ctx = nfh.Context(N, S)
    return d, ctx


def _prepare(oracle, d, ctx, freq0, F0, a0):
    N, S = d.n_ind, d.n_sites
    gl_ind = oracle.normalize_gl(np.transpose(d.log_gl, (1, 0, 2)))          # (N,S,3) as the reference holds it
    gl_site = np.ascontiguousarray(np.transpose(gl_ind, (1, 0, 2)))           # (S,N,3) upload layout
    ctx.upload_gl(gl_site)
    ctx.upload_pos_dist(d.dist_mb)
    freq = np.broadcast_to(np.asarray(freq0, dtype=np.float64), (S,)).copy()
    F = np.broadcast_to(np.asarray(F0, dtype=np.float64), (N,)).copy()
    a = np.broadcast_to(np.asarray(a0, dtype=np.float64), (N,)).copy()
    ctx.set_freq(freq)
    ctx.set_ind_params(F, a)
    ctx.emission_refresh()
    _, e = oracle.freq_emission(gl_ind, None, freq, update_freq=False)
    return gl_ind, freq, F, a, e


def _posterior_check(got, want, unclamped=None):
    diff = np.abs(got - want)
    bad = diff > POST_ATOL
    if bad.any():
        # threshold flips: the oracle's value is a clamp output while ours sits just across EPSILON
        near = (np.abs(got - 1e-5) < 1e-7) | (np.abs(got - (1 - 1e-5)) < 1e-7) | \
               (np.abs(want - 1e-5) < 1e-7) | (np.abs(want - (1 - 1e-5)) < 1e-7)
        flips = bad & ((want == 0) | (want == 1) | (got == 0) | (got == 1)) & (diff < 1.1e-5)
        assert not (bad & ~flips & ~near).any(), f"max posterior diff {diff.max()}"
        assert flips.sum() <= max(3, got.size // 20000), f"{flips.sum()} clamp flips"


@pytest.mark.parametrize("N,S,seed", [(6, 3000, 1), (20, 10000, 12345), (3, 2048, 5), (5, 2049, 6), (2, 17, 7)])
def test_estep_matches_oracle(oracle, N, S, seed):
    d, ctx = _setup(N, S, seed, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        rng = np.random.default_rng(seed)
        F0 = rng.uniform(0.01, 0.6, N); a0 = rng.uniform(0.005, 2.0, N)
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.1, F0, a0)
        lk = ctx.estep()
        post = ctx.get_posterior()
        st, marg1, lk_o = oracle.estep(e, d.dist_mb, F, a)
        assert st == 0
        np.testing.assert_allclose(lk, lk_o, rtol=LKL_RTOL, atol=0)
        _posterior_check(post, marg1)


def test_estep_chromosome_breaks_and_extreme_params(oracle):
    N, S = 4, 5000
    d, ctx = _setup(N, S, 11, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        d.dist_mb[[0, 1000, 2500, 4999]] = np.inf          # chromosome starts (read_data.cpp:207-209)
        F0 = np.array([1e-6, 1 - 1e-6, 0.3, 1e-15]); a0 = np.array([1e-6, 10.0, 1e-15, 0.5])
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.25, F0, a0)
        lk = ctx.estep(); post = ctx.get_posterior()
        st, marg1, lk_o = oracle.estep(e, d.dist_mb, F, a)
        np.testing.assert_allclose(lk, lk_o, rtol=LKL_RTOL)
        _posterior_check(post, marg1)


def test_emission_matches_oracle(oracle):
    N, S = 7, 4000
    d, ctx = _setup(N, S, 3, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        rng = np.random.default_rng(3)
        freq0 = rng.uniform(0.01, 0.99, S)
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, freq0, 0.1, 0.2)
        ctx.emission_refresh(with_e0=True)
        import ctypes as C
        # read the device planes back through the exchange windows with a raw cudaMemcpy via torch
        torch = pytest.importorskip("torch")
        ptr, nbytes, _ = ctx.window(nfh.api.WIN_EMIS_SEND)
        ptr0, _, _ = ctx.window(nfh.api.WIN_E0_SEND)
        ctx.sync()
        ratio = _dev_to_numpy(torch, ptr, nbytes).reshape(ctx.n_ind_local, ctx.site_block)[:N, :S]
        e0 = _dev_to_numpy(torch, ptr0, nbytes).reshape(ctx.n_ind_local, ctx.site_block)[:N, :S]
        np.testing.assert_allclose(np.log(e0), e[:, :, 0], rtol=0, atol=1e-13)
        np.testing.assert_allclose(np.log(ratio), e[:, :, 1] - e[:, :, 0], rtol=0, atol=1e-12)


def _dev_to_numpy(torch, ptr, nbytes):
    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(w, device="cuda").cpu().numpy().copy()


@pytest.mark.parametrize("N,S,seed", [(6, 3000, 21), (20, 10000, 12345), (13, 700, 22), (100, 600, 23)])
def test_freq_update_matches_oracle(oracle, N, S, seed):
    d, ctx = _setup(N, S, seed, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.1, 0.1, 0.2)
        ctx.estep()
        post = ctx.get_posterior()
        f_new = ctx.freq_update(1)
        f_o, e_o = oracle.freq_emission(gl_ind, post, freq, update_freq=True)
        np.testing.assert_allclose(f_new, f_o, rtol=0, atol=1e-11)
        # next E-step on the refreshed emissions agrees too (checks ratio + sum log e0)
        lk2 = ctx.estep()
        st, marg2, lk2_o = oracle.estep(e_o, d.dist_mb, F, a)
        np.testing.assert_allclose(lk2, lk2_o, rtol=LKL_RTOL)
        _posterior_check(ctx.get_posterior(), marg2)


@pytest.mark.parametrize("N,S,no_hybrid", [(150, 300, 0), (125, 300, 0), (450, 64, 0), (600, 48, 0), (820, 40, 0),
                                           (900, 40, 0), (990, 40, 0), (1000, 40, 0), (600, 48, 1), (1100, 40, 0),
                                           (4200, 12, 0)],
                         ids=["warp-G16", "warp-G16-K8", "warp-G32-global-acc", "hybrid-K19", "hybrid-K26",
                              "hybrid-13+16", "hybrid-13+18", "hybrid-14+18", "team-W2", "team-W4", "stream"])
def test_freq_update_large_n_variants(oracle, monkeypatch, N, S, no_hybrid):
    """More individuals than one 8-lane group holds: wider lane groups (tile prefetch with accumulators in
    shared memory or global scratch), the register + shared-memory hybrid, teams of warps, the streaming path."""
    if no_hybrid:
        monkeypatch.setenv("NFH_FREQ_NO_HYBRID", "1")
    d, ctx = _setup(N, S, 31, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.1, 0.1, 0.2)
        ctx.estep(); post = ctx.get_posterior()
        f_new = ctx.freq_update(1)
        f_o, e_o = oracle.freq_emission(gl_ind, post, freq, update_freq=True)
        np.testing.assert_allclose(f_new, f_o, rtol=0, atol=1e-11)
        lk2 = ctx.estep()
        st, marg2, lk2_o = oracle.estep(e_o, d.dist_mb, F, a)
        np.testing.assert_allclose(lk2, lk2_o, rtol=LKL_RTOL)


def test_freq_update_hard_calls_and_monomorphic_sites(oracle):
    """Called genotypes (GL exactly 0/1) with sites fixed for either allele or all heterozygous: the
    frequency reaches exactly 0 or 1 (allele odds 0 / unbounded) and posteriors are clamped to 0 or 1."""
    rng = np.random.default_rng(5)
    N, S = 12, 96
    d, ctx = _setup(N, S, 41, freq=(0.05, 0.5), indF=(0.0, 0.5))
    geno = rng.integers(0, 3, size=(N, S))
    geno[:, 0] = 0; geno[:, 1] = 2; geno[:, 2] = 1; geno[:, 3] = 2; geno[:6, 4] = 0; geno[6:, 4] = 2
    geno[:, 40:48] = 2; geno[:, 60:64] = 0
    gl = np.full((N, S, 3), -np.inf)
    np.put_along_axis(gl, geno[:, :, None], 0.0, axis=2)
    d.log_gl = np.ascontiguousarray(np.transpose(gl, (1, 0, 2)))
    with ctx:
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.2, 0.3, 0.5)
        for _ in range(2):
            ctx.estep()
            post = ctx.get_posterior()
            f_new = ctx.freq_update(1)
            f_o, e_o = oracle.freq_emission(gl_ind, post, freq, update_freq=True)
            assert np.isfinite(f_new).all()
            np.testing.assert_allclose(f_new, f_o, rtol=0, atol=1e-11)
            assert f_new[0] == 0.0 and abs(f_new[1] - 1.0) < 1e-15 and abs(f_new[2] - 0.5) < 1e-15
            freq = f_o


def test_freq_pass_counter_matches_oracle(oracle):
    """nfh_freq_passes (the work figure behind bench.py's roofline) = the oracle's est_maf pass counts."""
    d, ctx = _setup(20, 500, 77, freq=(0.02, 0.5), indF=(0.0, 0.5))
    with ctx:
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.1, 0.1, 0.2)
        ctx.estep(); post = ctx.get_posterior()
        ctx.freq_passes(reset=True)
        f_new = ctx.freq_update(1)
        total = ctx.freq_passes(reset=True)
        want = 0
        for s in range(d.n_sites):
            f_o, n = oracle.est_maf_counted(gl_ind[:, s, :], post[:, s])
            want += n
            assert abs(f_o - f_new[s]) < 1e-11
        assert total == want
        assert ctx.freq_passes() == 0


def test_freq_init_estimate_with_zero_posterior(oracle):
    """--freq e: est_maf with scalar F = 0 (parse_args.cpp:316-318)."""
    N, S = 10, 1500
    d, ctx = _setup(N, S, 41, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.1, 0.1, 0.2)
        f_new = ctx.freq_update(1, posterior_is_zero=True)
        f_o, _ = oracle.freq_emission(gl_ind, np.zeros((N, S)), freq, update_freq=True)
        np.testing.assert_allclose(f_new, f_o, rtol=0, atol=1e-11)


def test_lkl_batch_matches_oracle(oracle):
    N, S = 8, 6000
    d, ctx = _setup(N, S, 51, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.2, 0.1, 0.2)
        rng = np.random.default_rng(51)
        ind, Fs, As = [], [], []
        for i in range(N):
            x = rng.uniform(0.01, 0.9); y = rng.uniform(0.001, 3.0); h = 4.4e-6
            for (ff, aa) in [(x, y), (x - h, y), (x + h, y), (x, y - h), (x, y + h)][: 1 + (i % 5)]:
                ind.append(i); Fs.append(ff); As.append(aa)
        out = ctx.lkl_batch(ind, Fs, As)
        want = np.array([oracle.lkl(e[i], d.dist_mb, f, al) for i, f, al in zip(ind, Fs, As)])
        np.testing.assert_allclose(out, want, rtol=LKL_RTOL)
        # NaN / Inf parameters: the reference returns -1e15 (EM.cpp:454-456)
        out2 = ctx.lkl_batch([0, 1], [np.nan, 0.2], [0.1, np.inf])
        assert (out2 == -1e15).all()


@pytest.mark.parametrize("N,S,seed", [(6, 3000, 61), (20, 10000, 12345), (4, 9, 62)])
def test_viterbi_matches_oracle(oracle, N, S, seed):
    d, ctx = _setup(N, S, seed, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        rng = np.random.default_rng(seed)
        F0 = rng.uniform(0.02, 0.6, N); a0 = rng.uniform(0.005, 1.0, N)
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.15, F0, a0)
        ctx.emission_refresh(with_e0=True)
        path = ctx.viterbi()
        mism = 0
        for i in range(N):
            _, p = oracle.viterbi(e[i], d.dist_mb, F[i], a[i])
            mism += int((p != path[i]).sum())
        assert mism == 0


def test_geno_posterior_matches_oracle(oracle):
    N, S = 5, 800
    d, ctx = _setup(N, S, 71, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.3, 0.1, 0.2)
        rng = np.random.default_rng(71)
        path = (rng.random((N, S)) < 0.3).astype(np.int8)
        got = ctx.geno_posterior(path)
        want = np.empty((S, N, 3))
        for s in range(S):
            for i in range(N):
                prior = oracle.calc_HWE(freq[s], float(path[i, s]), True)
                pp = gl_ind[i, s] + prior
                m = pp.max(); pp = pp - (m + np.log(np.exp(pp - m).sum()))
                want[s, i] = np.exp(pp)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)


def test_long_sequence_adjudicated_by_extended_precision(oracle):
    """100,000 sites: the reference's log-space recursion is itself noisier than the 1e-8 posterior bound
    here (SURVEY.md finding 5), so the device E-step is checked against the long-double restatement;
    the log-likelihood still matches the reference arithmetic to 1e-9 relative."""
    N, S = 3, 100_000
    d, ctx = _setup(N, S, 91, freq=(0.05, 0.5), indF=(0.05, 0.5))
    with ctx:
        d.dist_mb[[0, 40_000]] = np.inf
        F0 = np.array([0.05, 0.3, 0.6]); a0 = np.array([0.02, 0.5, 3.0])
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.2, F0, a0)
        lk = ctx.estep(); post = ctx.get_posterior()
        ctx.emission_refresh(with_e0=True)
        path = ctx.viterbi()
        for i in range(N):
            m_ext, lk_ext = oracle.estep_extended(e[i], d.dist_mb, F[i], a[i])
            assert abs(lk[i] - lk_ext) <= 1e-9 * abs(lk_ext)
            assert abs(lk[i] - oracle.forward(e[i], d.dist_mb, F[i], a[i])) <= 1e-9 * abs(lk_ext)
            clamped = np.where(m_ext < 1e-5, 0.0, np.where(m_ext > 1 - 1e-5, 1.0, m_ext))
            diff = np.abs(post[i] - clamped)
            flips = (diff > 1e-8) & ((np.abs(m_ext - 1e-5) < 1e-9) | (np.abs(m_ext - (1 - 1e-5)) < 1e-9))
            assert ((diff > 1e-8) & ~flips).sum() == 0, diff.max()
            _, p = oracle.viterbi(e[i], d.dist_mb, F[i], a[i])
            mism = np.nonzero(p != path[i])[0]
            assert len(mism) <= 2, (i, len(mism))      # only exact near-ties may differ at this length


@pytest.mark.parametrize("N,S", [(1, 1), (1, 32), (2, 33), (3, 34), (2, 4224), (2, 4225)])
def test_tiny_and_tile_boundary_shapes(oracle, N, S):
    d, ctx = _setup(N, S, 17 + S, freq=(0.05, 0.5), indF=(0.05, 0.5))
    with ctx:
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.2, 0.25, 0.4)
        lk = ctx.estep(); post = ctx.get_posterior()
        st, marg1, lk_o = oracle.estep(e, d.dist_mb, F, a)
        np.testing.assert_allclose(lk, lk_o, rtol=LKL_RTOL)
        _posterior_check(post, marg1)
        f_new = ctx.freq_update(1)
        f_o, e_o = oracle.freq_emission(gl_ind, post, freq, update_freq=True)
        np.testing.assert_allclose(f_new, f_o, rtol=0, atol=1e-11)
        ctx.emission_refresh(with_e0=True)
        path = ctx.viterbi()
        for i in range(N):
            assert (oracle.viterbi(e_o[i], d.dist_mb, F[i], a[i])[1] == path[i]).all()
        out = ctx.lkl_batch(np.arange(N), F, a)
        want = np.array([oracle.lkl(e_o[i], d.dist_mb, F[i], a[i]) for i in range(N)])
        np.testing.assert_allclose(out, want, rtol=LKL_RTOL)


def test_full_size_properties_and_extended_precision(oracle):
    """BASELINE configs[1] width (1,000,000 sites) on 4 individuals.
    (a) Against the long-double restatement of forward/backward/posterior (the adjudicator of SURVEY.md
        finding 5 - the reference's own log-space recursion is noisier than 1e-8 at this length):
        log-likelihood 1e-9 relative, posterior 1e-8 absolute except clamp-threshold flips; the
        reference-arithmetic forward() must also agree to 1e-9 relative.
    (b) Properties that need no oracle: forward == backward log-likelihood (checked in-kernel, EM.cpp:166),
        objective(F, alpha) == -E-step lkl, posterior in [0, 1] with the clamp gaps empty, frequency update
        idempotent given the same posterior, tracts follow the posterior."""
    N, S = 4, 1_000_000
    import torch
    gen = sim.simulate_torch(N, S, device="cuda", seed=5)
    with nfh.Context(N, S) as ctx:
        ctx.upload_gl(gen["log_gl"]); ctx.upload_pos_dist(gen["dist_mb"])
        F = np.array([0.05, 0.2, 0.4, 0.6]); a = np.array([0.01, 0.05, 0.5, 2.0])
        ctx.set_freq(np.full(S, 0.1)); ctx.set_ind_params(F, a)
        ctx.emission_refresh()
        lk = ctx.estep()                                   # raises on NaN or Fw/Bw mismatch
        obj = ctx.lkl_batch(np.arange(N), F, a)
        np.testing.assert_allclose(-obj, lk, rtol=1e-12)
        post = ctx.get_posterior()
        assert post.min() >= 0 and post.max() <= 1
        assert not ((post > 0) & (post < 1e-5)).any() and not ((post < 1) & (post > 1 - 1e-5)).any()
        # (a) the adjudicator, one individual at a time (about a second each)
        gl_ind = np.ascontiguousarray(np.transpose(gen["log_gl"].numpy(), (1, 0, 2)))
        _, e = oracle.freq_emission(gl_ind, None, np.full(S, 0.1), update_freq=False)
        for i in range(N):
            m_ext, lk_ext = oracle.estep_extended(e[i], gen["dist_mb"], F[i], a[i])
            assert abs(lk[i] - lk_ext) <= 1e-9 * abs(lk_ext), (i, lk[i], lk_ext)
            assert abs(lk[i] - oracle.forward(e[i], gen["dist_mb"], F[i], a[i])) <= 1e-9 * abs(lk_ext)
            clamped = np.where(m_ext < 1e-5, 0.0, np.where(m_ext > 1 - 1e-5, 1.0, m_ext))
            diff = np.abs(post[i] - clamped)
            flips = (diff > 1e-8) & ((np.abs(m_ext - 1e-5) < 1e-9) | (np.abs(m_ext - (1 - 1e-5)) < 1e-9))
            assert ((diff > 1e-8) & ~flips).sum() == 0, (i, diff.max())
        del gl_ind, e
        f1 = ctx.freq_update(1)
        lk1 = ctx.estep()
        ctx.set_freq(np.full(S, 0.3))                      # the frequency EM ignores its previous value
        ctx.emission_refresh()
        # posterior window now holds the posterior of the NEW emissions; restore by re-running the sequence
        ctx.set_freq(np.full(S, 0.1)); ctx.emission_refresh(); ctx.estep()
        f2 = ctx.freq_update(1)
        np.testing.assert_array_equal(f1, f2)
        assert (f1 > 0).all() and (f1 < 1).all()
        ctx.emission_refresh(with_e0=True)
        path = ctx.viterbi()
        assert set(np.unique(path)) <= {0, 1}
        # tracts follow the posterior: where the posterior is confidently IBD the path is IBD
        post = ctx.get_posterior()
        sure = post > 0.999
        assert (path[sure] == 1).mean() > 0.99
    del torch


def test_error_statuses_follow_the_reference(oracle):
    """Conditions the reference turns into error()/exit: NaN in a recursion, an estimation
    method it cannot run, arguments outside the context geometry; NaN/Inf objective parameters return
    -1e15 (EM.cpp:454-456)."""
    d, ctx = _setup(4, 600, 9)
    with ctx:
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.1, 0.1, 0.2)
        out = ctx.lkl_batch(np.array([0, 0, 1], dtype=np.int32), np.array([0.1, np.nan, 0.2]),
                            np.array([0.2, 0.2, np.inf]))
        assert np.isfinite(out[0]) and out[1] == -1e15 and out[2] == -1e15
        with pytest.raises(nfh.NfhError) as err:
            ctx.freq_update(2)
        assert err.value.status == 2 and "MAF estimation method" in str(err.value)
        with pytest.raises(nfh.NfhError) as err:
            ctx.upload_gl(np.zeros((10, 4, 3)), first_site=595)
        assert err.value.status == 2
        bad = np.ascontiguousarray(np.transpose(gl_ind, (1, 0, 2))).copy()
        bad[300, 2, :] = np.nan
        ctx.upload_gl(bad)
        ctx.emission_refresh()
        with pytest.raises(nfh.NfhError) as err:
            ctx.estep()
        assert err.value.status == 3 and "invalid Lkl" in str(err.value)


def test_estep_with_batch_equals_the_two_calls(oracle):
    """nfh_estep_with_batch = nfh_estep + nfh_lkl_batch when every individual's first request is its current
    (F, alpha): same objective values, same ind_lkl and posterior (the forward products are shared)."""
    N, S = 7, 9000
    d, ctx = _setup(N, S, 17, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        rng = np.random.default_rng(3)
        F0 = rng.uniform(0.05, 0.6, N); a0 = rng.uniform(0.01, 2.0, N)
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.1, F0, a0)
        lk1 = ctx.estep()
        post1 = ctx.get_posterior()
        ind, Fq, aq = [], [], []
        for i in range(N):
            eh = 3e-6
            for dF, da in ((0, 0), (eh, 0), (-eh, 0), (0, eh), (0, -eh)):
                ind.append(i); Fq.append(F0[i] + dF); aq.append(a0[i] + da)
        want = ctx.lkl_batch(ind, Fq, aq)
        got, lk2 = ctx.estep_with_batch(ind, Fq, aq)
        post2 = ctx.get_posterior()
        np.testing.assert_array_equal(got, want)
        np.testing.assert_allclose(lk2, lk1, rtol=1e-13)
        np.testing.assert_allclose(post2, post1, rtol=0, atol=1e-12)
        np.testing.assert_allclose(-got[0::5], lk1, rtol=1e-12)          # the centre point is the E-step's lkl
        # a first request that is not the current parameter point is refused
        Fq[5] += 1e-3
        with pytest.raises(nfh.NfhError) as err:
            ctx.estep_with_batch(ind, Fq, aq)
        assert err.value.status == 2


def test_iteration_with_parameters_outside_the_optimiser_box(oracle):
    """A caller may hand iter_EM parameters the optimiser's box does not contain (alpha > 10): the
    reference runs forward/backward at the parameters as given (EM.cpp:151-185) and only setulb_ projects them
    for the optimiser.  The fused E-step / first-round path must step aside instead of refusing."""
    N, S = 3, 2000
    d, ctx = _setup(N, S, 23, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        F0 = np.array([0.3, 0.4, 0.2]); a0 = np.array([12.0, 0.2, 0.3])
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.2, F0, a0)
        runner = nfh.EmRank(ctx, freq_est=1)
        Fi, ai = F0.copy(), a0.copy()
        lk, fr = runner.iteration(Fi, ai)
        st, marg1, lk_o = oracle.estep(e, d.dist_mb, F0, a0)
        np.testing.assert_allclose(lk, lk_o, rtol=LKL_RTOL)
        assert (ai <= 10.0).all()                                      # the optimiser projected it


@pytest.mark.parametrize("wave_rows,lookahead", [(1, 0), (2, 3), (3, 1000), (100, 5)])
def test_single_launch_estep_matches_three_launch_path_and_oracle(oracle, monkeypatch, wave_rows, lookahead):
    """nfh_estep's single-launch kernel (csrc/nfh_schedule.h) for several wave shapes, against the three-launch
    path (NFH_ESTEP_FUSED=0) and the oracle: 7 individuals x 30,000 sites = 8 tiles each, parameters per
    individual, a chromosome start in the middle; repeated calls reuse the never-reset device counters."""
    N, S = 7, 30000
    d, ctx = _setup(N, S, 21, freq=(0.05, 0.5), indF=(0.0, 0.5))
    with ctx:
        d.dist_mb[[0, 17000]] = np.inf
        rng = np.random.default_rng(4)
        F0 = rng.uniform(0.01, 0.6, N); a0 = rng.uniform(0.005, 2.0, N)
        gl_ind, freq, F, a, e = _prepare(oracle, d, ctx, 0.2, F0, a0)
        monkeypatch.setenv("NFH_ESTEP_FUSED", "0")
        lk3 = ctx.estep(); post3 = ctx.get_posterior()
        monkeypatch.setenv("NFH_ESTEP_FUSED", "1")
        monkeypatch.setenv("NFH_ESTEP_WAVE_ROWS", str(wave_rows))
        monkeypatch.setenv("NFH_ESTEP_LOOKAHEAD", str(lookahead))
        for rep in range(3):
            lk1 = ctx.estep(); post1 = ctx.get_posterior()
            np.testing.assert_allclose(lk1, lk3, rtol=1e-12, atol=0)
            assert np.abs(post1 - post3).max() < 1e-11
        st, marg1, lk_o = oracle.estep(e, d.dist_mb, F, a)
        np.testing.assert_allclose(lk1, lk_o, rtol=LKL_RTOL, atol=0)
        _posterior_check(post1, marg1)
