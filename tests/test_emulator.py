"""CPU emulation of the CUDA kernels' per-thread code (tests/emul/lnx_emul.cu) against the oracle.

The emulator executes the exact __host__ __device__ phase functions of leniax_b200/csrc/*.cuh thread by thread, with the
kernels' synchronisation points as loop boundaries: index maps, twiddles, packing tricks and the statistics formulas are
validated here without a GPU.  (The product never loads this library.)
"""
import ctypes
import itertools
import os

import numpy as np
import pytest

from oracle import lenia_oracle as lo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, 'tests', 'emul', 'liblnx_emul.so')
KEYS = ['mass', 'mass_volume', 'mass_density', 'growth', 'growth_volume', 'growth_density', 'mass_speed', 'mass_angle_speed',
        'mass_growth_dist', 'inertia', 'potential_volume']


@pytest.fixture(scope='module')
def emul():
    if not os.path.exists(EMUL):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(EMUL)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def build_tables(emul, Kfull):
    Kt, Kpq = np.zeros((16, 256, 4), np.float32), np.zeros((32, 4, 4), np.float32)
    emul.lnx_emul_build_kt(P(np.ascontiguousarray(Kfull.astype(np.complex64))), P(Kt), P(Kpq))
    return Kt, Kpq


def test_potential_real_symmetric_and_general_kernel(emul, golden_dir):
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-test.yaml'))
    _, K, _ = lo.init(cfg)
    rng = np.random.default_rng(0)
    state = rng.random((128, 128), dtype=np.float32)
    pot = np.zeros((128, 128), np.float32)
    Kt, Kpq = build_tables(emul, K[0, 0, 0])
    emul.lnx_emul_potential(P(state), P(Kt), P(Kpq), P(pot))
    assert np.abs(pot - lo.get_potential_fft(state[None, None], K)[0, 0]).max() < 1e-6
    # non-symmetric kernel => complex spectrum; exercises the packed DC|Nyquist column formulas (Kp, Kq)
    kern = rng.random((128, 128)).astype(np.float32)
    kern /= kern.sum()
    ref = np.real(np.fft.ifft2(np.fft.fft2(state.astype(np.float64)) * np.fft.fft2(kern.astype(np.float64))))
    Kt, Kpq = build_tables(emul, np.fft.fft2(kern))
    emul.lnx_emul_potential(P(state), P(Kt), P(Kpq), P(pot))
    assert np.abs(pot - ref).max() < 5e-7


def _run_fused(emul, cells0, K, gf, m, s, w, mean, T, sf, R, steps):
    Kt, Kpq = build_tables(emul, K)
    stats = np.zeros((11, steps), np.float32)
    cm, N, fin = np.zeros(steps, np.float32), np.zeros(1, np.float32), np.zeros((128, 128), np.float32)
    f = ctypes.c_float
    rc = emul.lnx_emul_run_fused(P(np.ascontiguousarray(cells0)), P(Kt), P(Kpq), gf, f(m), f(s), f(w), mean, f(T), sf, f(R),
                                 f(1. / T), steps, P(stats), P(cm), P(N), P(fin), None)
    assert rc == 0
    return stats, cm, float(N[0]), fin


def test_orbium_trajectory_stats_and_golden(emul, golden_dir):
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-test.yaml'))
    cells, K, _ = lo.init(cfg)
    oc, _, _, ostats = lo.init_and_run(cfg, with_jit=True)
    stats, cm, N, fin = _run_fused(emul, cells[0, 0], K[0, 0, 0], 0, .15, .015, 1., 1, 10., 0, 13., 128)
    tol = dict(zip(KEYS, [2e-6, 3e-7, 2e-6, 5e-6, 3e-7, 1e-5, 5e-5, 0.05, 2e-5, 5e-6, 0.05]))
    for i, k in enumerate(KEYS):
        assert np.abs(stats[i] - ostats[k][:, 0]).max() <= tol[k], k
    assert np.abs(cm - ostats['channel_mass'][:, 0, 0]).max() < 2e-6
    assert N == float(ostats['N'][0]) == 128.
    # state after 127 updates vs the reference's golden fixture (tests/test_pipeline.py:18-54, decimal=4)
    _, _, _, fin127 = _run_fused(emul, cells[0, 0], K[0, 0, 0], 0, .15, .015, 1., 1, 10., 0, 13., 127)
    gold = np.load(os.path.join(golden_dir, 'orbium-test_last_frame.npy'))
    np.testing.assert_array_almost_equal(gold[0], fin127, decimal=4)
    assert np.abs(fin127 - oc[-1, 0, 0]).max() < 5e-5


def test_register_carried_state_flow_matches_smem_flow_and_oracle(emul, golden_dir):
    """The TMEM kernel's cell phase (state chunks in a thread-private store, new state carried in registers into the next
    phase 1, arithmetic packed over the thread's two rows, coordinate table) against the shared-memory kernel's cell phase
    and the oracle."""
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-test.yaml'))
    cells, K, _ = lo.init(cfg)
    n = 64
    stats, cm, N, fin = _run_fused(emul, cells[0, 0], K[0, 0, 0], 0, .15, .015, 1., 1, 10., 0, 13., n)
    Kt, Kpq = build_tables(emul, K[0, 0, 0])
    stats2, cm2, N2, fin2 = np.zeros((11, n), np.float32), np.zeros(n, np.float32), np.zeros(1, np.float32), np.zeros((128, 128), np.float32)
    f = ctypes.c_float
    emul.lnx_emul_run_fused_rs(P(np.ascontiguousarray(cells[0, 0])), P(Kt), P(Kpq), f(.15), f(.015), f(1.), 1, f(10.), f(13.), f(.1), n,
                               P(stats2), P(cm2), P(N2), P(fin2))
    assert np.abs(fin - fin2).max() < 2e-5  # same physics; rounding differs (fused multiply-adds), amplified over 64 steps
    cfg['run_params']['max_run_iter'] = n
    ostats = lo.init_and_run(cfg, with_jit=True)[3]
    tol = dict(zip(KEYS, [2e-6, 3e-7, 2e-6, 5e-6, 3e-7, 1e-5, 5e-5, 0.05, 2e-5, 5e-6, 0.05]))
    for i, k in enumerate(KEYS):
        assert np.abs(stats2[i] - ostats[k][:, 0]).max() <= tol[k], k
    np.testing.assert_array_equal(stats[1], stats2[1])  # counts (mass_volume) are exact in both
    assert N == float(N2[0]) == float(n)


def test_dying_world_stops_like_check_heuristics(emul, golden_dir):
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-test.yaml'))
    cells, K, mapping = lo.init(cfg)
    weak = (cells * 0.2).astype(np.float32)  # fades away: channel mass falls under epsilon
    steps = 40
    upd = lo.build_update_fn(mapping)
    sfn = lo.build_compute_stats_fn(cfg['world_params'], cfg['render_params'])
    ostats, _ = lo.run_scan(weak, K, mapping.get_gf_params(), mapping.get_kernels_weight_per_channel(), np.float32(10.), steps, upd, sfn, False)
    stats, cm, N, _ = _run_fused(emul, weak[0, 0], K[0, 0, 0], 0, .15, .015, 1., 1, 10., 0, 13., steps)
    assert N == float(ostats['N'][0]) and N < steps
    assert np.abs(stats[0] - ostats['mass'][:, 0]).max() < 2e-6


@pytest.mark.parametrize('gf,sf,slug,sslug,params', [(1, 0, 'gaussian', 'v1', (.15, .02)), (2, 1, 'gaussian_target', 'v2', (.2, .05)),
                                                     (5, 0, 'triangle', 'v1', (.15, .05)), (6, 2, 'identity', 'simple', (0., 1.))])
def test_other_growth_and_state_functions(emul, golden_dir, gf, sf, slug, sslug, params):
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-test.yaml'))
    cfg['kernels_params'][0]['gf_slug'] = slug
    cfg['kernels_params'][0]['gf_params'] = list(params)
    cfg['world_params']['get_state_fn_slug'] = sslug
    cfg['run_params']['max_run_iter'] = 6
    cells, K, _ = lo.init(cfg)
    oc, _, _, ostats = lo.init_and_run(cfg, with_jit=True)
    stats, cm, N, fin = _run_fused(emul, cells[0, 0], K[0, 0, 0], gf, params[0], params[1], 1., 1, 10., sf, 13., 5)
    assert np.abs(fin - oc[5, 0, 0]).max() < 2e-5
    assert np.abs(stats[0] - ostats['mass'][:5, 0]).max() < 1e-4 * max(1., np.abs(ostats['mass']).max())


# ---- shared-memory bank conflicts of every access pattern of the exchange buffer (addresses from the real code) ----
def _conflict_degree(byte_addrs, width):
    """Max wavefronts needed by one warp-wide access: lanes are served in groups of 128/width... (8 lanes for 128-bit,
    16 for 64-bit); within a group every 4-byte bank may be touched once."""
    group = {16: 8, 8: 16, 4: 32}[width]
    worst = 1
    for g0 in range(0, 32, group):
        banks = {}
        for a in byte_addrs[g0:g0 + group]:
            for wd in range(width // 4):
                b = ((a // 4) + wd) % 32
                banks.setdefault(b, set()).add((a // 4) + wd)
        worst = max(worst, max(len(v) for v in banks.values()))
    return worst


def test_exchange_layouts_are_bank_conflict_free(emul):
    e1, e2 = emul.lnx_emul_e1_addr, emul.lnx_emul_e2_addr
    k1_of, col_of = emul.lnx_emul_k1_of, emul.lnx_emul_col_of
    for warp in range(8):
        tids = [warp * 32 + ln for ln in range(32)]
        G = [t >> 4 for t in tids]
        sub = [t & 15 for t in tids]
        # P1 stores / P5 loads: 64-bit, lanes (q,l), fixed k1
        for k1 in range(32):
            addrs = [(G[i] * 512 + e1(sub[i] >> 2, k1, sub[i] & 3)) * 8 for i in range(32)]
            assert _conflict_degree(addrs, 8) == 1, ('E1 64-bit', warp, k1)
        # P2 loads / P4 stores: 128-bit, lanes a, fixed (q, s, half)
        for q, s, h in itertools.product(range(4), range(2), range(2)):
            addrs = [(G[i] * 512 + e1(q, k1_of(sub[i], s), 2 * h)) * 8 for i in range(32)]
            assert all(a % 16 == 0 for a in addrs)
            assert _conflict_degree(addrs, 16) == 1, ('E1 128-bit', warp, q, s, h)
        # P2 stores / P4 loads: 128-bit, lanes a, fixed (c, u)
        for c, u in itertools.product(range(4), range(4)):
            addrs = [(G[i] * 512 + e2(col_of(sub[i], c), u)) * 8 for i in range(32)]
            assert _conflict_degree(addrs, 16) == 1, ('E2 group', warp, c, u)
        # P3 loads / stores: 128-bit, lanes (col, bidx), fixed r1
        for r1 in range(16):
            addrs = [(r1 * 512 + e2(t >> 2, t & 3)) * 8 for t in tids]
            assert _conflict_degree(addrs, 16) == 1, ('E2 column', warp, r1)


def test_exchange_layouts_are_bijections(emul):
    e1, e2 = emul.lnx_emul_e1_addr, emul.lnx_emul_e2_addr
    assert sorted(e1(q, k, l) for q in range(4) for k in range(32) for l in range(4)) == list(range(512))
    assert sorted(e2(c, u) + e for c in range(64) for u in range(4) for e in range(2)) == list(range(512))
    cols = sorted(emul.lnx_emul_col_of(a, c) for a in range(16) for c in range(4))
    assert cols == list(range(64))

# ---------------------------------------------------------------------------------------------------------------------
# 64^3 thread-per-line engine (leniax_b200/csrc/lnx_tiled64.cuh), emulated lane by lane (tests/emul/lnx_t64_emul.cu)
# ---------------------------------------------------------------------------------------------------------------------
EMUL64 = os.path.join(ROOT, 'tests', 'emul', 'liblnx_t64_emul.so')


@pytest.fixture(scope='module')
def emul64():
    if not os.path.exists(EMUL64):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(EMUL64)
    f, i, v = ctypes.c_float, ctypes.c_int, ctypes.c_void_p
    lib.lnx_t64_emul_step.argtypes = [v, v, i, f, f, f, i, i, f, v, v, v, v, v]
    lib.lnx_t64h_emul_step.argtypes = [v, v, i, f, f, f, i, i, f, v, v, v, v, v, i]
    return lib


def _shell_kernel_3d(R):
    r = np.arange(-R, R)
    g = np.stack(np.meshgrid(r, r, r, indexing='ij')).astype(np.float32) / R
    dist = np.sqrt((g ** 2).sum(0))
    kern = ((dist < 1) * (4 * dist * (1 - dist)) ** 4).astype(np.float32)
    return (kern / kern.sum())[None]


def test_t64_rfftn_matches_numpy(emul64):
    rng = np.random.default_rng(0)
    w = rng.random((64, 64, 64), dtype=np.float32)
    spec = np.zeros((64, 64, 33), np.complex64)
    emul64.lnx_t64_emul_rfftn(P(w), P(spec))
    ref = np.fft.rfftn(w.astype(np.float64))
    assert np.abs(spec - ref).max() < 2e-7 * np.abs(ref).max()


@pytest.mark.parametrize('engine', ['line', 'half_line', 'half_line_finite'])
@pytest.mark.parametrize('gf_slug,gf_id,sf_slug,sf_id,mean', [('poly_quad4', 0, 'v1', 0, 1), ('gaussian', 1, 'v2', 1, 0), ('gaussian', 1, 'v1', 0, 1),
                                                              ('poly_quad4', 0, 'v1', 0, 0)])
def test_t64_step_matches_oracle(emul64, gf_slug, gf_id, sf_slug, sf_id, mean, engine):
    """One Lenia step of a 64^3 world (plane_fwd -> lead -> plane_inv; 'half_line': the two-threads-per-line kernels of
    lnx_tiled64h.cuh) against the oracle, and the per-plane statistics partial sums against direct sums in the rolled frame
    (statistics.py:64-100)."""
    D, R = 64, 13
    kp = [dict(k_slug='raw', k_params=_shell_kernel_3d(R), kf_slug='poly_quad', kf_params=[4], gf_slug=gf_slug, gf_params=[.15, .015], h=.7,
               c_in=0, c_out=0)]
    oK, om = lo.get_kernels_and_mapping(kp, [D, D, D], 1, R)
    ktab = np.ascontiguousarray((oK[0, 0, 0][:, :, :33] / D ** 3).astype(np.complex64))
    rng = np.random.default_rng(1)
    state = (rng.random((D, D, D), dtype=np.float32) * 0.3).astype(np.float32)
    gf, wt = om.get_gf_params(), om.get_kernels_weight_per_channel()
    ns, of, op = lo.build_update_fn(om, sf_slug, bool(mean))(state[None, None], oK, gf, wt, np.float32(0.1))
    st, pot, fld = state.copy(), np.zeros_like(state), np.zeros_like(state)
    NP = emul64.lnx_t64_emul_np()
    part = np.zeros((64, NP), np.float32)
    shift = np.array([5, 60, 17], np.int32)
    nxt = np.zeros((64, 64, 33), np.complex64)  # fused tail: axes-(1, 2) half spectra of the NEW cells, plane by plane
    args = (P(st), P(ktab), gf_id, float(gf[0, 0]), float(gf[0, 1]), float(wt[0, 0]), mean, sf_id, 0.1, P(shift), P(pot), P(fld), P(part), P(nxt))
    if engine == 'line':
        emul64.lnx_t64_emul_step(*args)
    else:  # '_finite': single-instruction clamps (LNX_RUN_ASSUME_FINITE), otherwise the NaN-propagating forms
        emul64.lnx_t64h_emul_step(*args, int(engine == 'half_line_finite'))
    ref_nxt = np.fft.rfft2(st.astype(np.float64), axes=(1, 2))
    assert np.abs(nxt - ref_nxt).max() < 2e-7 * np.abs(ref_nxt).max()
    assert np.abs(pot - op[0, 0]).max() < 1e-6
    assert np.abs(fld - of[0, 0]).max() < 2e-5
    assert np.abs(st - ns[0, 0]).max() < 2e-6
    a, f = state.astype(np.float64), of[0, 0].astype(np.float64)
    gp = np.maximum(f, 0)
    idx = np.indices((D, D, D))
    xs = [((idx[d] - shift[d]) % D) - D // 2 for d in range(3)]
    exp = ([(a > 1e-7).sum(), gp.sum(), (gp > 1e-7).sum(), (op[0, 0] > 1e-7).sum()] + [(a * x).sum() for x in xs] + [(a * x * x).sum() for x in xs] +
           [(gp * x).sum() for x in xs] + [a.sum()])
    tot = part.astype(np.float64).sum(0)
    scale = [1, 1, 1, 1] + [np.abs(a * x).sum() for x in xs] + [1] * 3 + [np.abs(gp * x).sum() + 1 for x in xs] + [1]
    for i, e in enumerate(exp):
        assert abs(tot[i] - e) <= 3e-5 * max(abs(e), scale[i]), (i, tot[i], e)
    assert np.all(tot[len(exp):] == 0)


# ---------------------------------------------------------------------------------------------------------------------
# 2048^2 four-step engine (leniax_b200/csrc/lnx_tiled2k.cuh), emulated lane by lane (tests/emul/lnx_t2k_emul.cu)
# ---------------------------------------------------------------------------------------------------------------------
EMUL2K = os.path.join(ROOT, 'tests', 'emul', 'liblnx_t2k_emul.so')


@pytest.fixture(scope='module')
def emul2k():
    if not os.path.exists(EMUL2K):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(EMUL2K)
    f, i, v = ctypes.c_float, ctypes.c_int, ctypes.c_void_p
    lib.lnx_t2k_emul_step.argtypes = [v, v, i, f, f, f, i, i, f, v, v, v, v, v, i, i]
    lib.lnx_t2k_emul_rfft2.argtypes = [v, v, i]
    return lib


@pytest.mark.parametrize('real_rows', [0, 1])
def test_t2k_rfft2_matches_numpy(emul2k, real_rows):
    rng = np.random.default_rng(0)
    w = rng.random((2048, 2048), dtype=np.float32)
    spec = np.zeros((2048, 1025), np.complex64)
    emul2k.lnx_t2k_emul_rfft2(P(w), P(spec), real_rows)
    ref = np.fft.rfft2(w.astype(np.float64))
    assert np.abs(spec - ref).max() < 3e-7 * np.abs(ref).max()


@pytest.mark.parametrize('finite,real_rows', [(-1, 0), (0, 0), (1, 0), (-1, 1), (1, 1)])
def test_t2k_step_matches_oracle(emul2k, finite, real_rows):
    """One Lenia step of a 2048^2 world, R = 52 (rows_fwd -> lead -> rows_inv) against the oracle, and the statistics partial
    sums of the row pairs against direct sums in the rolled frame (statistics.py:64-100).  finite = -1: cell phase with per-cell
    selection of the growth function; 0 / 1: the packed poly_quad4 form, NaN-propagating / single-instruction clamps.  real_rows: the
    rows kernels with one real row per warp (round 2) instead of a packed row pair."""
    S, R = 2048, 52
    kp = [dict(k_slug='circle_2d', k_params=[1., [1.]], kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4', gf_params=[.15, .015], h=1.,
               c_in=0, c_out=0)]
    oK, om = lo.get_kernels_and_mapping(kp, [S, S], 1, R)
    Kh = np.ascontiguousarray(oK[0, 0, 0][:, :S // 2 + 1].astype(np.complex64))
    rng = np.random.default_rng(2)
    state = np.zeros((S, S), np.float32)
    for _ in range(40):
        y, x = rng.integers(0, S - 200, 2)
        state[y:y + 200, x:x + 200] = rng.random((200, 200), dtype=np.float32) * 0.35
    gf, wt = om.get_gf_params(), om.get_kernels_weight_per_channel()
    ns, of, op = lo.build_update_fn(om)(state[None, None], oK, gf, wt, np.float32(0.1))
    st, pot, fld = state.copy(), np.zeros_like(state), np.zeros_like(state)
    NP = emul2k.lnx_t2k_emul_np()
    part = np.zeros((S if real_rows else S // 2, NP), np.float32)
    shift = np.array([700, 1999], np.int32)
    nxt = np.zeros((1025, 2048), np.complex64)  # fused tail: transposed row spectra T[k][row] of the NEW cells
    emul2k.lnx_t2k_emul_step(P(st), P(Kh), 0, float(gf[0, 0]), float(gf[0, 1]), float(wt[0, 0]), 1, 0, 0.1, P(shift), P(pot), P(fld), P(part),
                             P(nxt), finite, real_rows)
    ref_nxt = np.fft.rfft(st.astype(np.float64), axis=1).T
    assert np.abs(nxt - ref_nxt).max() < 3e-7 * np.abs(ref_nxt).max()
    assert np.abs(pot - op[0, 0]).max() < 1e-6
    assert np.abs(fld - of[0, 0]).max() < 5e-5
    assert np.abs(st - ns[0, 0]).max() < 5e-6
    a, f = state.astype(np.float64), of[0, 0].astype(np.float64)
    gp = np.maximum(f, 0)
    idx = np.indices((S, S))
    xs = [((idx[d] - shift[d]) % S) - S // 2 for d in range(2)]
    exp = {0: (a > 1e-7).sum(), 1: gp.sum(), 2: (gp > 1e-7).sum(), 4: (a * xs[0]).sum(), 5: (a * xs[1]).sum(), 7: (a * xs[0] ** 2).sum(),
           8: (a * xs[1] ** 2).sum(), 10: (gp * xs[0]).sum(), 11: (gp * xs[1]).sum(), 13: a.sum()}
    tot = part.astype(np.float64).sum(0)
    for i in range(NP):
        if i == 3:  # potential > eps: far from the patterns the potential is rounding noise of that size
            assert abs(tot[3] - (op[0, 0] > 1e-7).sum()) < 1e-3 * S * S
            continue
        e = exp.get(i, 0.)
        assert abs(tot[i] - e) <= 2e-5 * max(abs(e), 1.), (i, tot[i], e)
