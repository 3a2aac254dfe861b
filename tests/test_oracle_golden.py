"""Pin the oracle: the reference's own golden fixtures and known-answer tests.

CPU only.  Every check cites the reference test it mirrors.
"""
import os

import numpy as np
import pytest

from oracle import lenia_oracle as lo


def _run(golden_dir, name, **kw):
    cfg = lo.load_yaml_config(os.path.join(golden_dir, name + '.yaml'))
    return cfg, lo.init_and_run(cfg, **kw)


# tests/test_pipeline.py:18-54 (decimal=4), :56-92 (decimal=4), :94-130 (decimal=3)
@pytest.mark.parametrize('name,steps,decimal', [('orbium-test', 128, 4), ('orbium-scutium-test', 128, 4),
                                                ('aquarium-test', 32, 3)])
def test_golden_last_frame_scan(golden_dir, name, steps, decimal):
    _, (cells, field, potential, stats) = _run(golden_dir, name, with_jit=True)
    gold = np.load(os.path.join(golden_dir, name + '_last_frame.npy'))
    assert cells.shape[0] == steps
    np.testing.assert_array_almost_equal(gold, cells[-1, 0], decimal=decimal)
    assert stats['N'].shape == (1, )


def test_golden_last_frame_python_loop(golden_dir):  # tests/test_pipeline.py:18-35 (with_jit=False)
    _, (cells, _, _, stats) = _run(golden_dir, 'orbium-test', with_jit=False)
    gold = np.load(os.path.join(golden_dir, 'orbium-test_last_frame.npy'))
    assert len(cells) == 128
    np.testing.assert_array_almost_equal(gold, cells[-1, 0], decimal=4)
    assert int(stats['N']) == 127  # loop index when no break (runner.py:114)


def test_fp64_twin_is_close_to_golden(golden_dir):
    _, (cells, _, _, _) = _run(golden_dir, 'orbium-test', with_jit=True, dtype=np.float64)
    gold = np.load(os.path.join(golden_dir, 'orbium-test_last_frame.npy'))
    assert np.abs(cells[-1, 0] - gold).max() < 5e-5


def test_conv_potential_kat():  # tests/test_core.py:55-103 (exact equality there; fp32 sums here)
    C = 2
    cells = np.ones([1, C, 2, 2], np.float32)
    cells[0, 0] = 0.1
    cells[0, 1] = 0.2
    K = np.ones([4, 1, 3, 3], np.float32)
    K[0], K[1], K[2], K[3] = 0.1, 0.2, 0.3, 0
    pot = lo.get_potential_conv(cells, K, tc_indices=(0, 1, 2))
    assert pot.shape == (1, 3, 2, 2)
    np.testing.assert_allclose(pot[0, 0], 0.09, rtol=1e-6)
    np.testing.assert_allclose(pot[0, 1], 0.18, rtol=1e-6)
    np.testing.assert_allclose(pot[0, 2], 0.54, rtol=1e-6)


def test_state_update_kats():  # tests/test_core.py:159-185 (v1) and :187-213 (v2 shapes)
    c1, f = np.full([25, 25], .5, np.float32), np.full([25, 25], 3, np.float32)
    np.testing.assert_array_almost_equal(lo.get_state_v1(c1, f, np.float32(1. / 3.)), np.ones([25, 25]))
    c2 = np.full([25, 25], .2, np.float32)
    np.testing.assert_array_almost_equal(lo.get_state_v1(c2, f, np.float32(1. / 6.)), np.full([25, 25], .7))
    np.testing.assert_array_almost_equal(lo.get_state_v2(c1, f, np.float32(0.5)), np.full([25, 25], 1.75))
    np.testing.assert_array_almost_equal(lo.get_state_simple(c1, f, np.float32(0.5)), np.full([25, 25], 2.))


def test_mass_volume_heuristic_kat():  # tests/test_statistics.py:14-31
    mv = np.array([800, 1600, 2400, 3200], np.float32) / np.float32(13.**2)
    cnt = np.array([10, 70, lo.MASS_VOLUME_STOP_STEP - 1, lo.MASS_VOLUME_STOP_STEP])
    ok, nxt = lo.mass_volume_heuristic(mv, cnt)
    np.testing.assert_array_equal(ok, [True, True, True, False])
    np.testing.assert_array_equal(nxt, [1, 1, lo.MASS_VOLUME_STOP_STEP, lo.MASS_VOLUME_STOP_STEP + 1])


def test_monotonic_heuristic_kat():  # tests/test_statistics.py:33-50
    sign = np.sign(np.array([1.1, 0.9, 0.9, 0.9]) - 1)
    prev = np.array([1, 1, -1, -1])
    cnt = np.array([40, lo.MONOTONIC_STOP_STEP, 30, lo.MONOTONIC_STOP_STEP])
    ok, nxt = lo.monotonic_heuristic(sign, prev, cnt)
    np.testing.assert_array_equal(ok, [True, True, True, False])
    np.testing.assert_array_equal(nxt, [41, 1, 31, lo.MONOTONIC_STOP_STEP + 1])


def test_circle_2d_shape_kat():  # tests/test_kernels.py:12-20
    k = lo.circle_2d(5., [1., [1.]], 'poly_quad', [4])
    assert k.shape == (1, 10, 10)
    np.testing.assert_allclose(k.sum(), 1., rtol=1e-5)


def test_kernel_shapes_like_test_qd(golden_dir):  # tests/test_qd.py:80-88
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-test.yaml'))
    K, m = lo.get_kernels_and_mapping(cfg['kernels_params'], [128, 128], 1, 13, fft=False)
    assert K.shape == (1, 1, 25, 25)  # cropped Orbium kernel is 2R-1, odd and centred
    K, m = lo.get_kernels_and_mapping(cfg['kernels_params'], [128, 128], 1, 13, fft=True)
    assert K.shape == (1, 1, 1, 128, 128) and K.dtype == np.complex64
    # circle kernels are even functions => real spectrum (exploited by the CUDA path, checked there too)
    assert np.abs(K.imag).max() < 1e-6


def test_fft_potential_matches_direct_conv(golden_dir):
    """FFT path (core.py:52-102) vs direct depthwise conv (core.py:105-146): same potential up to fp32."""
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-scutium-test.yaml'))
    cells, Kf, mp = lo.init(cfg, fft=True)
    cfg2 = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-scutium-test.yaml'))
    _, Kc, mp2 = lo.init(cfg2, fft=False)
    pf = lo.get_potential_fft(cells, Kf, lo.tc_indices_of(mp))
    pc = lo.get_potential_conv(cells, Kc, lo.tc_indices_of(mp2))
    assert pf.shape == pc.shape == (1, 2, 128, 128)
    assert np.abs(pf - pc).max() < 2e-6


def test_true_channels_padding():
    """kernels.py:122-143: channels with fewer kernels are padded and masked."""
    kp = [dict(k_slug='circle_2d', k_params=[1., [1.]], kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4',
               gf_params=[.15, .015], h=1., c_in=ci, c_out=co) for ci, co in [(1, 0), (0, 0), (0, 1)]]
    K, m = lo.get_kernels_and_mapping(kp, [32, 32], 2, 5, fft=True)
    assert [p['c_in'] for p in kp] == [0, 0, 1]  # sorted in place, stable
    assert K.shape == (1, 2, 2, 32, 32)
    assert m.true_channels == [True, True, True, False]
    assert lo.tc_indices_of(m) == (0, 1, 2)
    W = m.get_kernels_weight_per_channel()
    np.testing.assert_array_equal(W, [[1, 0, 1], [0, 1, 0]])


def test_cell_decoders(golden_dir):
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium.yaml'))  # gzip/b64/int32 codec, loader.py:105-129
    cells = lo.load_raw_cells(cfg, use_init_cells=False)
    assert cells.shape == (1, 20, 20)
    np.testing.assert_allclose(cells.sum(), 73.66, atol=0.01)
    init_cells = lo.load_raw_cells(cfg, use_init_cells=True)
    assert init_cells.shape == (1, 128, 128)
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-test.yaml'))  # legacy RLE, loader.py:287-350
    rle = lo.load_raw_cells(cfg)
    assert rle.shape == (1, 20, 20)
    assert 0 <= rle.min() and rle.max() <= 1
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-scutium-test.yaml'))
    assert lo.load_raw_cells(cfg).shape[0] == 2


def test_check_heuristics_first_failure_index():
    """runner.py:161-162: N = number of leading steps with should_continue == 1."""
    T, N = 300, 3
    mass = np.ones((T, N), np.float32)
    mass[:, 0] += np.arange(T, dtype=np.float32) * 1e-3  # strictly increasing => monotone stop
    mass[:, 1] = 1 + 0.01 * ((np.arange(T) % 2) * 2 - 1)  # oscillating => survives
    mass[:, 2] = 1 + 0.01 * ((np.arange(T) % 2) * 2 - 1)
    stats = {'mass': mass, 'channel_mass': mass[..., None].copy(), 'mass_volume': np.ones((T, N), np.float32)}
    stats['mass_volume'][:, 2] = 11.  # above threshold for ever => volume stop
    n = lo.check_heuristics(stats).sum(axis=0)
    # t=0: sign 0 == prev 0 -> counter 1; t=1: sign +1 != 0 -> counter 1; counter reaches 129 at t=129
    assert n.tolist() == [129., 300., 128.]


def test_stats_shapes_and_carry(golden_dir):
    cfg, (cells, field, pot, stats) = _run(golden_dir, 'orbium-test', with_jit=True)
    for k in lo.STAT_KEYS:
        assert stats[k].shape == (128, 1), k
    assert stats['channel_mass'].shape == (128, 1, 1)
    assert float(stats['N'][0]) == 128.
    np.testing.assert_allclose(stats['mass'][5, 0], cells[5].sum() / 169., rtol=1e-5)
    assert 0.3 < stats['mass_speed'][64:, 0].mean() < 0.7  # Orbium glides at ~0.5 R/T... sanity only
