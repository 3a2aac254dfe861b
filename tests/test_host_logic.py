"""Host-side logic of the product package and the C-ABI surface (CPU only: no compute calls)."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import leniax_b200
from leniax_b200 import _lib, core, growth_functions, helpers, kernels, loader, runner, statistics, utils
from oracle import lenia_oracle as lo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_loads_and_exports_every_declared_symbol():
    lib = leniax_b200.load_library()
    header = open(os.path.join(ROOT, 'include', 'leniax_b200.h')).read()
    declared = set(re.findall(r'^(?:int|size_t|const char\*)\s+(lnx_[a-z0-9_]+)\(', header, flags=re.M))
    assert declared >= {'lnx_plan_create', 'lnx_plan_destroy', 'lnx_run_scan', 'lnx_kernels_prepare', 'lnx_rfft2', 'lnx_last_error'}
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/leniax_b200.h but not exported'
    assert set(_lib.EXPORTS) == declared
    assert lib.lnx_version() == 100
    if not torch.cuda.is_available():
        assert lib.lnx_device_count() == 0


def test_no_cpu_fallback_without_gpu(golden_dir):
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    cfg = utils.load_config(os.path.join(golden_dir, 'orbium-test.yaml'))
    with pytest.raises(_lib.LeniaxB200Error):
        helpers.init(cfg)  # the kernel spectrum needs lnx_rfft2
    K, mapping = kernels.get_kernels_and_mapping(cfg['kernels_params'], [128, 128], 1, 13, fft=False)
    ufn = helpers.build_update_fn((1, 1, 1, 128, 128), mapping)
    sfn = statistics.build_compute_stats_fn(cfg['world_params'], cfg['render_params'])
    cells = torch.zeros(1, 1, 128, 128)
    with pytest.raises(_lib.LeniaxB200Error):
        runner.run_scan(None, cells, torch.zeros(1, 1, 1, 128, 128, dtype=torch.complex64), mapping.get_gf_params(),
                        mapping.get_kernels_weight_per_channel(), 10., 4, 13, ufn, sfn)
    d = _lib.LnxDesc(nb_dims=2, nb_channels=1, nb_kernels=1, nb_slots=1, R=13., stats_dt=.1)
    d.dims[0] = d.dims[1] = 128
    handle = ctypes.c_void_p()
    rc = leniax_b200.load_library().lnx_plan_create(ctypes.byref(d), ctypes.byref(handle))
    assert rc == _lib.LNX_ERR_NO_DEVICE and b'no CPU fallback' in leniax_b200.load_library().lnx_last_error()


def test_plan_validation_errors():
    lib = leniax_b200.load_library()
    handle = ctypes.c_void_p()
    d = _lib.LnxDesc(nb_dims=2, nb_channels=1, nb_kernels=1, nb_slots=1, R=13., stats_dt=.1)
    d.dims[0], d.dims[1] = 5000, 100  # beyond 4096 (a size that is merely not a power of two gives a statistics-only plan, on a GPU)
    assert lib.lnx_plan_create(ctypes.byref(d), ctypes.byref(handle)) == _lib.LNX_ERR_UNSUPPORTED
    d.dims[0] = d.dims[1] = 128
    d.nb_channels = 99
    assert lib.lnx_plan_create(ctypes.byref(d), ctypes.byref(handle)) == _lib.LNX_ERR_INVALID
    d.nb_channels, d.gf_id[0] = 1, 42
    assert lib.lnx_plan_create(ctypes.byref(d), ctypes.byref(handle)) == _lib.LNX_ERR_INVALID
    assert b'growth function' in lib.lnx_last_error()
    with pytest.raises(ValueError):
        _lib.check(_lib.LNX_ERR_INVALID)
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.LNX_ERR_UNSUPPORTED)


@pytest.mark.parametrize('name', ['orbium-test', 'orbium-scutium-test', 'aquarium-test'])
def test_kernel_rasterisation_and_mapping_match_oracle(golden_dir, name):
    cfg = utils.load_config(os.path.join(golden_dir, name + '.yaml'))
    ocfg = lo.load_yaml_config(os.path.join(golden_dir, name + '.yaml'))
    wp = cfg['world_params']
    K, m = kernels.get_kernels_and_mapping(cfg['kernels_params'], [128, 128], wp['nb_channels'], wp['R'], fft=False, device='cpu')
    Ko, mo = lo.get_kernels_and_mapping(ocfg['kernels_params'], [128, 128], wp['nb_channels'], wp['R'], fft=False)
    assert tuple(K.shape) == Ko.shape
    np.testing.assert_allclose(K.numpy(), Ko, atol=5e-8)
    assert m.true_channels == mo.true_channels and m.cin_kernels == mo.cin_kernels and m.cin_gfs == mo.cin_gfs
    np.testing.assert_array_equal(m.get_gf_params().numpy(), mo.get_gf_params())
    np.testing.assert_array_equal(m.get_kernels_weight_per_channel().numpy(), mo.get_kernels_weight_per_channel())
    assert [p['c_in'] for p in cfg['kernels_params']] == sorted(p['c_in'] for p in cfg['kernels_params'])  # sorted in place


def test_ellipse_kernels_match_oracle():
    for fn, ofn in ((kernels.ellipse_2d, lo.ellipse_2d), (kernels.oriented_ellipse_2d, lo.oriented_ellipse_2d)):
        a = fn(13, [1., [1., .5], 1.2, .8, .25], 'gauss_bump', [4], device='cpu').numpy()
        b = ofn(13, [1., [1., .5], 1.2, .8, .25], 'gauss_bump', [4])
        np.testing.assert_allclose(a, b, atol=2e-7)
    np.testing.assert_array_equal(kernels.circle_2d(5., [1., [1.]], 'poly_quad', [4], device='cpu').shape, [1, 10, 10])  # test_kernels.py:12-20


def test_update_fn_descriptor_layout(golden_dir):
    cfg = utils.load_config(os.path.join(golden_dir, 'aquarium-test.yaml'))
    K, m = kernels.get_kernels_and_mapping(cfg['kernels_params'], [128, 128], 3, 12, fft=False, device='cpu')
    ufn = helpers.build_update_fn((1, 3, 5, 128, 128), m, 'v1', False, True)
    slots, c_in, gf_ids = ufn.kernel_layout(3)
    assert len(slots) == 15 and c_in == (0, ) * 5 + (1, ) * 5 + (2, ) * 5 and set(gf_ids) == {1}
    assert ufn.get_field_fn.average is False and ufn.get_state_fn.slug == 'v1'
    # padded slots (kernels.py:122-143): 2 channels, 2+1 kernels -> tc_indices (0, 1, 2)
    kp = [dict(k_slug='circle_2d', k_params=[1., [1.]], kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4', gf_params=[.15, .015],
               h=1., c_in=ci, c_out=co) for ci, co in [(1, 0), (0, 0), (0, 1)]]
    K, m = kernels.get_kernels_and_mapping(kp, [128, 128], 2, 13, fft=False, device='cpu')
    ufn = helpers.build_update_fn((1, 2, 2, 128, 128), m)
    assert ufn.kernel_layout(2) == ((0, 1, 2), (0, 0, 1), (0, 0, 0))
    with pytest.raises(NotImplementedError):
        helpers.build_update_fn((1, 2, 2, 128, 128), m, fft=False)
    with pytest.raises(NotImplementedError):
        growth_functions.resolve(lambda p, x: x)
    with pytest.raises(NotImplementedError):
        core.update(None, torch.zeros(1, 1, 128, 128), None, None, None, .1, lambda *a: a, lambda *a: a, 'v1')


def test_loaders_and_config_upgrade_match_oracle(golden_dir):
    for name in ('orbium-test', 'orbium-scutium-test', 'aquarium-test', 'orbium'):
        cfg = utils.load_config(os.path.join(golden_dir, name + '.yaml'))
        ocfg = lo.load_yaml_config(os.path.join(golden_dir, name + '.yaml'))
        assert cfg['kernels_params'] == ocfg['kernels_params'] and cfg['render_params']['world_size'] == [128, 128]
        np.testing.assert_array_equal(loader.load_raw_cells(cfg).numpy(), lo.load_raw_cells(ocfg))
    cells = helpers.create_init_cells([128, 128], 1, [loader.load_raw_cells(cfg, False)])
    assert cells.shape == (1, 1, 128, 128) and float(cells[0, 0, 54:74, 54:74].sum()) == pytest.approx(float(cells.sum()))
    assert utils.st2fracs2float('1,2/3,6.7') == pytest.approx([1., 2 / 3, 6.7])  # tests/test_utils.py
    d = {}
    utils.set_param(d, 'kernels_params.1.gf_params.0', .3)
    assert utils.get_param(d, 'kernels_params.1.gf_params.0') == .3
    q = loader.make_array_compressible(torch.tensor([.12345678]))
    assert float(q) == pytest.approx(round(.12345678 * 12543) / 12543)


def test_host_heuristics_kats():  # tests/test_statistics.py:14-50
    mv = torch.tensor([800., 1600., 2400., 3200.]) / 13.**2
    ok, nxt = statistics.mass_volume_heuristic(mv, torch.tensor([10, 70, 127, 128]))
    assert ok.tolist() == [True, True, True, False] and nxt.tolist() == [1, 1, 128, 129]
    sign = torch.sign(torch.tensor([1.1, .9, .9, .9]) - 1)
    ok, nxt = statistics.monotonic_heuristic(sign, torch.tensor([1., 1., -1., -1.]), torch.tensor([40, 128, 30, 128]))
    assert ok.tolist() == [True, True, True, False] and nxt.tolist() == [41, 1, 31, 129]
    T, N = 300, 2
    mass = torch.ones(T, N)
    mass[:, 0] += torch.arange(T) * 1e-3
    mass[:, 1] = 1 + .01 * ((torch.arange(T) % 2) * 2 - 1)
    st = {'mass': mass, 'channel_mass': mass[..., None].clone(), 'mass_volume': torch.ones(T, N)}
    ref = lo.check_heuristics({k: v.numpy() for k, v in st.items()})
    np.testing.assert_array_equal(statistics.check_heuristics(st).numpy(), ref)


def test_grid_archive_index_matches_oracle_restatement():
    """qd.grid_archive_index (torch, any device) against the oracle's restatement of ribs 0.4.0 GridArchive.get_index, including
    values on and beyond the domain borders."""
    import numpy as np
    import torch
    from leniax_b200 import qd
    from oracle import lenia_oracle as lo
    rng = np.random.default_rng(0)
    feats = np.concatenate([rng.uniform(-0.2, 1.2, (500, 2)), [[0., 0.], [1., 1.], [0.05, 0.999999], [0.5, 0.05 * 7]]])
    dom, shape = [[0., 1.], [0., 1.]], [20, 20]
    got = qd.grid_archive_index(torch.from_numpy(feats), shape, dom).numpy()
    want = lo.grid_archive_index(feats, shape, dom)
    assert (got == want).all()
    assert got.min() >= 0 and got.max() <= 19


def test_zoom_nearest_is_scipy_order0_zoom():
    """utils.zoom_nearest restates scipy.ndimage.zoom(x, scale, order=0) (leniax/helpers.py:61)."""
    import numpy as np
    import scipy.ndimage
    import torch
    from leniax_b200 import utils
    rng = np.random.default_rng(0)
    for shape in [(20, 20), (17, 23), (5, 9), (20, 18, 7)]:
        for sc in [2, 3, 4, 1.5, 0.5, 2.5]:
            a = rng.random(shape).astype(np.float32)
            ref = scipy.ndimage.zoom(a, sc, order=0)
            got = utils.zoom_nearest(torch.from_numpy(a), sc).numpy()
            assert ref.shape == got.shape and np.array_equal(ref, got), (shape, sc)


def test_vectorised_heuristic_counter_matches_the_recurrence():
    """runner.run replays the reference loop's counters (statistics.py:287-306, 317-333) over all steps at once."""
    from leniax_b200.runner import _heuristic_counter
    rng = np.random.default_rng(0)
    for p in (0.02, 0.5, 0.98, 1.0, 0.0):
        keep = rng.random(700) < p
        c, ref = 0, []
        for k in keep:
            c = c * int(k) + 1
            ref.append(c)
        assert _heuristic_counter(keep).tolist() == ref


def test_header_compiles_as_plain_c_and_a_c_program_binds_the_library(tmp_path):
    """include/leniax_b200.h is the contract a cgo / JNI / ctypes binding is written against: it must be valid C (no C++, no
    CUDA or torch types) and a plain C program linked against the shared library must reach the entry points.  No compute call
    is made: a plan for an unsupported world shape has to come back as LNX_ERR_UNSUPPORTED with a message."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('gcc not on PATH')
    libdir = os.path.join(ROOT, 'leniax_b200')
    src = tmp_path / 'bind.c'
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "leniax_b200.h"
int main(void) {
    lnx_desc d;
    lnx_plan* plan = NULL;
    memset(&d, 0, sizeof d);
    d.nb_dims = 2; d.dims[0] = 5000; d.dims[1] = 100;  /* (100 x 100 gives a statistics-only plan: any size up to 4096) */
    d.nb_channels = 1; d.nb_kernels = 1; d.nb_slots = 1; d.R = 13.f; d.stats_dt = .1f;
    d.gf_id[0] = LNX_GF_POLY_QUAD4; d.state_fn = LNX_STATE_V1;
    int rc = lnx_plan_create(&d, &plan);
    printf("%d %d %d %zu %s\n", lnx_version(), rc, plan == NULL, sizeof(lnx_desc), lnx_last_error());
    return 0;
}
''')
    exe = tmp_path / 'bind'
    subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe),
                    '-L', libdir, '-lleniax_b200', f'-Wl,-rpath,{libdir}'], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split(' ', 4)
    assert int(out[0]) == 100 and int(out[1]) == _lib.LNX_ERR_UNSUPPORTED and int(out[2]) == 1
    assert int(out[3]) == ctypes.sizeof(_lib.LnxDesc)  # the ctypes mirror and the C struct agree on the layout
    assert 'must be in [1, 4096]' in out[4]


def test_parallel_early_exit_oracle_equals_run_scan(golden_dir):
    """oracle/parallel.py (used by the GPU integer-parity tests): dropping a world from the oracle's batch once its stop criteria
    fired and 128 rows exist gives the same N and the same rows [ns-128, ns) as lo.run_scan; also through the process pool."""
    from oracle import parallel as opar
    cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-test.yaml'))
    rng = np.random.default_rng(0)
    ang = (2 * np.pi * rng.random((7, 3, 4))).astype(np.float32)
    worlds = lo.perlin_from_angles(ang, [128, 128], 13, [.15, .015])
    c, K, m = lo.init(copy.deepcopy(cfg))
    worlds = np.concatenate([worlds, c])
    steps = 150
    keys = ('mass_density', 'mass_speed')
    ref, _ = lo.run_scan(worlds, K, m.get_gf_params(), m.get_kernels_weight_per_channel(), np.float32(10.), steps, lo.build_update_fn(m),
                         lo.build_compute_stats_fn(cfg['world_params'], cfg['render_params']), False)
    for procs in (1, 2):
        r = opar.parallel_scan(cfg['kernels_params'], cfg['world_params'], cfg['render_params'], worlds, steps, keys, chunk=4, procs=procs)
        assert r['N'].tolist() == ref['N'].tolist()
        assert (r['steps'] <= steps).all() and r['steps'].min() < steps  # some worlds really left early
        for k in keys:
            for i in range(len(worlds)):
                ns = max(int(ref['N'][i]), 128)
                assert abs(ref[k][ns - 128:ns, i].mean(dtype=np.float64) - r[k][i]) < 1e-6, (k, i)


def test_kernel_cache_key_distinguishes_numpy_and_tensor_scalars():
    """np.float32 / 0-d tensor parameters used to freeze to None, so different kernels shared one cache entry."""
    import torch
    from leniax_b200 import kernels

    def key(kf_params, bs):
        kp = [dict(k_slug='circle_2d', k_params=[1., bs], kf_slug='poly_quad', kf_params=kf_params, c_in=0)]
        return kernels._kernel_cache_key(kp, [128, 128], 1, 13., True, 'cpu')

    assert key([np.float32(4)], [1.]) != key([np.float32(1)], [1.])
    assert key([4], [np.float32(1), np.float32(.5)]) != key([4], [np.float32(.2), np.float32(1)])
    assert key([torch.tensor(4.)], [1.]) != key([torch.tensor(2.)], [1.])
    assert key([np.float32(4)], [1.]) == key([4.0], [1.0])  # same values, same key
    raw = [dict(k_slug='raw', k_params=np.ones((1, 3, 3), np.float32), kf_slug='poly_quad', kf_params=[4], c_in=0)]
    assert kernels._kernel_cache_key(raw, [128, 128], 1, 13., True, 'cpu') is None  # arrays are not cached
    nested = [dict(k_slug='circle_2d', k_params=[1., [1., np.ones(2)]], kf_slug='poly_quad', kf_params=[4], c_in=0)]
    assert kernels._kernel_cache_key(nested, [128, 128], 1, 13., True, 'cpu') is None  # ... at any nesting level


def test_cells_codec_round_trip_and_clear_error():
    """compress_array / the npz and plain-base64 decoders of leniax/loader.py:33-146 were missing."""
    import base64
    import io

    import torch
    from leniax_b200 import loader
    rng = np.random.default_rng(0)
    cells = loader.make_array_compressible(torch.from_numpy(rng.random((1, 9, 7), dtype=np.float32)))
    s = loader.compress_array(cells)
    assert isinstance(s, str) and torch.equal(loader.decompress_array(s, 3), cells)
    assert np.array_equal(lo.decompress_array_gzip(s), cells.numpy())  # the oracle's decoder reads what we write
    # plain base64 = a pickled array (loader.py:132-146); anything but NumPy reconstruction is refused
    import pickle
    assert torch.equal(loader.decompress_array(base64.b64encode(pickle.dumps(cells.numpy())).decode(), 3), cells)
    evil = base64.b64encode(pickle.dumps(print)).decode()
    with pytest.raises(ValueError, match='no decoder'):
        loader.decompress_array(evil, 3)
    # npz archive as a latin1 string
    buf = io.BytesIO()
    np.savez(buf, x=(rng.random((2, 5, 5)) * 255).astype(np.uint8))
    got = loader.decompress_array(buf.getvalue().decode('latin1'), 3)
    assert got.shape == (2, 5, 5) and float(got.max()) <= 1.
    with pytest.raises(ValueError, match='no decoder'):
        loader.decompress_array('~~~ not a cells string ~~~', 3)


def _schedule(c_in, c_out, C):
    lib = _lib.load_library()
    K = len(c_in)
    d = _lib.LnxDesc(nb_dims=2, nb_channels=C, nb_kernels=K, nb_slots=K, R=13., stats_dt=.1)
    d.dims[0] = d.dims[1] = 128
    for k in range(K):
        d.c_in[k], d.c_out[k], d.slot[k] = c_in[k], c_out[k], k
    arr = lambda n: (ctypes.c_int32 * n)()  # noqa: E731
    slot, first, upd, chan = arr(K), arr(K), arr(K), arr(C)
    rc = lib.lnx_gen2_schedule(ctypes.byref(d), slot, first, upd, chan)
    return rc, list(slot), list(first), list(upd), list(chan)


def test_gen2_schedule_orders_channel_updates_for_two_accumulators():
    """Host side of lnx_world128_gen2 (csrc/lnx_kernel_gen2.cuh: gen2_schedule): a channel is updated after its last kernel has been added
    AND its own spectrum has been taken; never more than two accumulators are live; graphs that need three are refused."""
    # conf/config_qd_cmame_3c6k.yaml: kernels sorted by c_in (kernels.py:90)
    rc, slot, first, upd, chan = _schedule([0, 0, 1, 1, 2, 2], [0, 1, 1, 2, 2, 0], 3)
    assert rc == 1
    assert first == [1, 1, 0, 1, 0, 0]                      # kernels 0, 1, 3 open the accumulators of channels 0, 1, 2
    assert upd == [0, 0, 0b010, 0, 0b100, 0b001]            # channel 1 after kernel 2, channel 2 after kernel 4, channel 0 after kernel 5
    assert slot[0] == slot[5] and slot[1] == slot[2] and slot[3] == slot[4] and slot[0] != slot[1] and slot[3] == slot[1]
    assert chan == [slot[0], slot[1], slot[3]]
    # generic invariants on random graphs: every channel updated exactly once, at or after its last writer and its first reader; a slot is
    # never shared by two live channels; refused graphs really need three
    rng = np.random.default_rng(0)
    accepted = refused = 0
    for _ in range(300):
        C, K = int(rng.integers(1, 5)), int(rng.integers(1, 9))
        c_in = sorted(int(v) for v in rng.integers(0, C, K))
        c_out = [int(v) if rng.random() > .1 else _lib.LNX_COUT_NONE for v in rng.integers(0, C, K)]
        rc, slot, first, upd, chan = _schedule(c_in, c_out, C)
        first_in = {c: min((k for k in range(K) if c_in[k] == c), default=-1) for c in range(C)}
        last_out = {c: max((k for k in range(K) if c_out[k] == c), default=-1) for c in range(C)}
        first_out = {c: min((k for k in range(K) if c_out[k] == c), default=K) for c in range(C)}
        want = {c: max(last_out[c], first_in[c]) if max(last_out[c], first_in[c]) >= 0 else K - 1 for c in range(C)}
        live = lambda k: [c for c in range(C) if first_out[c] <= k <= want[c]]  # noqa: E731  accumulators alive during kernel k
        need = max(len(live(k)) for k in range(K))
        if rc == 0:
            refused += 1
            assert need > 2, (c_in, c_out)
            continue
        accepted += 1
        assert need <= 2
        updated = [c for k in range(K) for c in range(C) if upd[k] >> c & 1]
        assert sorted(updated) == list(range(C))
        for c in range(C):
            assert upd[want[c]] >> c & 1
            assert (chan[c] >= 0) == (last_out[c] >= 0)
        for k in range(K):
            cs = live(k)
            assert len({chan[c] for c in cs}) == len(cs)  # distinct slots while alive together
            if c_out[k] >= 0:
                assert slot[k] == chan[c_out[k]] and first[k] == (1 if k == first_out[c_out[k]] else 0)
    assert accepted > 50 and refused > 20
    # undeclared weight pattern -> not eligible
    assert _schedule([0, 0], [_lib.LNX_COUT_ANY] * 2, 1)[0] == 0


def test_param_summary_one_sync_flags_and_pattern():
    import torch
    from leniax_b200 import runner
    gf = torch.tensor([[[.15, .015], [.2, .03], [.3, .04]]])
    w = torch.tensor([[[.5, 0., .7], [0., 1., 0.]]])
    finite, c_out = runner._param_summary(gf, w, True)
    assert finite and c_out == (0, 1, 0)
    assert runner._param_summary(gf, w, True) == (finite, c_out)                      # cached for the same tensors
    w2 = w.clone()
    w2[0, 1, 0] = .2                                                                   # kernel 0 feeds two channels: pattern undeclared
    assert runner._param_summary(gf, w2, True)[1] == (_lib.LNX_COUT_ANY, ) * 3
    w3 = torch.tensor([[[.5, 0., 0.], [0., 1., 0.]]])                                  # kernel 2 feeds nothing
    assert runner._param_summary(gf, w3, True) == (True, (0, 1, _lib.LNX_COUT_NONE))
    assert runner._param_summary(torch.tensor([[[.15, 0.], [.2, .03], [.3, .04]]]), w, True)[0] is False   # s == 0
    assert runner._param_summary(gf, torch.tensor([[[.5, 0., .7], [0., 0., 0.]]]), True)[0] is False        # zero weight row (mean)
    assert runner._param_summary(gf, torch.tensor([[[.5, 0., .7], [0., 0., 0.]]]), False)[0] is True        # ... fine for weighted_sum
    w.mul_(2.)                                                                         # in-place change invalidates the cache entry
    assert runner._param_summary(gf, w, True) == (True, (0, 1, 0))


def test_taps_recovered_from_a_non_power_of_two_kernel_spectrum():
    """kernels.spatial_from_spectrum: the direct-convolution taps recovered from K = fftn(fftshift(centre-pad(kernel))) of a 100 x 120 world
    (the reference's FFT potential takes any size, core.py:81) reproduce real(ifftn(fftn(state) * K)) as the cross-correlation
    lnx_update_conv computes (lnx_conv.cuh): potential[y][x] = sum_ij state[(y + i - kh/2) mod H][(x + j - kw/2) mod W] taps[i][j]."""
    import torch
    from leniax_b200 import kernels
    from oracle import lenia_oracle as lo
    H, W, R = 100, 120, 13
    kp = [dict(k_slug='circle_2d', k_params=[1., [1., .5]], kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4', gf_params=[.15, .015], h=1.,
               c_in=0, c_out=0)]
    oK, _ = lo.get_kernels_and_mapping(kp, [H, W], 1, R)
    assert not kernels.is_pow2_world([H, W]) and kernels.is_pow2_world([64, 2048])
    taps = kernels.spatial_from_spectrum(torch.from_numpy(oK.astype(np.complex64)), 1, [H, W]).numpy()[0, 0]
    kh, kw = taps.shape
    assert kh % 2 == 1 and kw % 2 == 1 and 2 * R - 3 <= kh <= 2 * R + 3 and abs(taps.sum() - 1.) < 1e-5
    rng = np.random.default_rng(0)
    state = rng.random((H, W))
    ref = np.real(np.fft.ifft2(np.fft.fft2(state) * oK[0, 0, 0].astype(np.complex128)))
    pot = np.zeros_like(state)
    for i in range(kh):
        for j in range(kw):
            pot += np.roll(state, (-(i - kh // 2), -(j - kw // 2)), axis=(0, 1)) * taps[i, j]
    assert np.abs(pot - ref).max() < 2e-6
