"""CPU: host-side logic that needs neither a GPU nor the reference."""
import gzip
import os

import numpy as np

import ngsf_hmm_b200 as nfh
from ngsf_hmm_b200 import sim
from _host import minimize


def test_minimize_finds_box_constrained_optimum():
    f = lambda v: (v[0] - 0.3) ** 2 + 2.0 * (v[1] - 20.0) ** 2      # optimum outside the upper bound of x1
    x, trace, ne = minimize(f, [0.1, 0.2], [1e-15, 1e-15], [1 - 1e-15, 10.0])
    assert abs(x[0] - 0.3) < 1e-5 and x[1] == 10.0
    assert ne == len(trace) and ne < 200


def test_minimize_fixed_coordinate_is_not_evaluated_off_its_value():
    seen = []
    def f(v):
        seen.append(v.copy())
        return (v[0] - 0.7) ** 2 + (v[1] - 0.2) ** 2
    x, _, _ = minimize(f, [0.1, 0.5], [1e-15, 0.5], [1 - 1e-15, 0.5])
    assert x[1] == 0.5 and abs(x[0] - 0.7) < 1e-5
    assert all(p[1] == 0.5 for p in seen)


def test_simulator_shapes_and_normalisation():
    d = sim.simulate(4, 500, seed=3, freq=(0.05, 0.5), indF=(0.0, 0.5))
    assert d.log_gl.shape == (500, 4, 3) and d.dist_mb.shape == (500,)
    np.testing.assert_allclose(np.exp(d.log_gl).sum(-1), 1.0, atol=1e-8)
    assert (d.dist_mb > 0).all() and set(np.unique(d.true_path)) <= {0, 1}
    d2 = sim.simulate(4, 500, seed=3, freq=(0.05, 0.5), indF=(0.0, 0.5))
    np.testing.assert_array_equal(d.log_gl, d2.log_gl)


def test_torch_generator_is_consistent_across_site_ranges():
    full = sim.simulate_torch(5, 3000, device="cpu", site_chunk=1024)
    part = sim.simulate_torch(5, 3000, device="cpu", site_chunk=1024, site_begin=1500, site_end=2600)
    np.testing.assert_array_equal(part["log_gl"].numpy(), full["log_gl"].numpy()[1500:2600])
    np.testing.assert_allclose(np.exp(full["log_gl"].numpy()).sum(-1), 1.0, atol=1e-12)


def test_input_file_writers_roundtrip(tmp_path):
    d = sim.simulate(3, 50, seed=9)
    p = str(tmp_path / "gl.bin")
    sim.write_binary_gl(p, d.log_gl)
    back = np.fromfile(p, dtype=np.float64).reshape(50, 3, 3)
    np.testing.assert_array_equal(back, d.log_gl)
    sim.write_pos(str(tmp_path / "pos.gz"), d.pos_bp)
    lines = gzip.open(str(tmp_path / "pos.gz"), "rt").read().split("\n")
    assert lines[0].split("\t") == ["chr1", str(int(d.pos_bp[0]))] and len(lines) == 51
    sim.write_beagle_gz(str(tmp_path / "b.gz"), d.log_gl, d.pos_bp)
    rows = gzip.open(str(tmp_path / "b.gz"), "rt").read().strip().split("\n")
    assert len(rows) == 51 and len(rows[1].split("\t")) == 3 + 9


def test_blocked_layout_shape():
    assert nfh.em.blocked_owner_layout(4, 13, 4224) == (4, 13, 4224)


def test_posterior_ready_hook_fires_once_for_a_rank_without_individuals():
    """A multi-rank caller starts a collective in the hook of nfh_host_estep_bfgs_update_hook: a rank that owns no
    individual (11 individuals on 8 ranks leave two of them empty) must still reach it, exactly once.  With no
    individual nothing touches the device context, so this runs without a GPU."""
    import ctypes as C
    from _host import HOST_SO
    L = C.CDLL(HOST_SO)
    HOOK = C.CFUNCTYPE(None, C.c_void_p)
    calls = []
    hook = HOOK(lambda user: calls.append(user))
    dp = C.POINTER(C.c_double)
    fn = L.nfh_host_estep_bfgs_update_hook
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint64, dp, dp, C.c_int, C.c_int, dp, C.POINTER(C.c_uint64), HOOK, C.c_void_p]
    empty = (C.c_double * 1)()
    stats = (C.c_uint64 * 3)()
    rc = fn(None, 0, empty, empty, 0, 0, empty, stats, hook, None)
    assert rc == 0 and len(calls) == 1 and stats[0] == 0
