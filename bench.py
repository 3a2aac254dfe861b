#!/usr/bin/env python
"""bench.py — cell-updates/s of the Lenia scan on B200 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W                      # B200 arm, BASELINE configs[1] (the quoted metric)
    python bench.py --config C|D|E [--scaling strong|weak] ...         # the other BASELINE configs as the primary line
    python bench.py --impl reference [--config ...] --gpus N ...       # reference arm: the CPU restatement of leniax on host cores

One "step" = one batched scan call over the whole workload through the reference-facing entry point
(`distributed.run_scan_mem_optimized_sharded` -> `runner.run_scan_mem_optimized`): every world simulated for all its steps with
all 12 statistics and the stop criteria, the device-side summary `qd.update_individuals` reads, and one NCCL all-gather of it.

  B  configs[1]: 4096 Orbium worlds (1 channel, 1 kernel, 128x128, R=13, T=10) x 1024 steps.  Worlds are the Orbium of
     conf/species/2d/1c-1k/orbium.yaml at random toroidal shifts, so every world survives and no step is skipped (early stop OFF).
     --scaling strong (default): the 4096 worlds BASELINE names, split over the N GPUs (512 per GPU at N = 8);
     --scaling weak: 4096 worlds per GPU (round-1 form; reported as `weak` inside the default line too).
  C  configs[2]: 3 channels, 6 kernels, 16 solutions x 128 perlin initialisations x 1024 steps; solutions split over the GPUs.
  D  configs[3]: one 2048x2048 world, R = 52, 256 steps (does not shard: N replicas).
  E  configs[4]: 256 worlds 64^3, 64 steps; worlds split over the GPUs.

The default run also times C (on all N GPUs) and, at N = 1, D, E and the search form of B (perlin soups, early stop) and reports
them under `secondary` in the same JSON line.
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WS = 128
FP32_PEAK_ANALYTIC_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.45 at the 1965 MHz max clock
# algorithmic flops (resident kernels: FP32-pipe bound) / bytes (2048^2 multi-pass engine: HBM convention) per cell-update
CONFIGS = {
    'B': dict(index=1, worlds=4096, sim_steps=1024, cells=WS * WS, flop=110.0, bound='fp32',
              metric='cell-updates/sec (batched 128x128 Orbium search, stats on)'),
    'C': dict(index=2, worlds=2048, sim_steps=1024, cells=WS * WS, flop=486.0, bound='fp32',
              metric='cell-updates/sec (3-channel 6-kernel 128x128 QD generation of 2048 evaluations, stats on)'),
    'D': dict(index=3, worlds=1, sim_steps=256, cells=2048 * 2048, flop=150.0, bytes=32.0, bound='hbm',
              metric='cell-updates/sec (single 2048x2048 world, R=52, stats on)'),
    'E': dict(index=4, worlds=256, sim_steps=64, cells=64**3, flop=136.0, bytes=32.0, bound='fp32',
              metric='cell-updates/sec (256 worlds 64^3, 3-D FFT potential, stats on)'),
}


def describe_workload(config, total_worlds, sim_steps):
    """The `config.workload` string: identical in the B200 arm and in the reference arm for the same flags."""
    what = {'B': f'configs[1]: {total_worlds} Orbium worlds in total, 1c1k 128x128 R=13 T=10',
            'Bsearch': f'configs[1] in its search form: {total_worlds} perlin soups, Orbium physics, early stop ON (extension; N unchanged)',
            'C': f'configs[2]: 3 channels 6 kernels 128x128, {total_worlds // 128} solutions x 128 perlin inits in total, per-solution K/gf/W/T',
            'D': f'configs[3]: {total_worlds} world(s) 2048x2048 (one per GPU: the path does not shard a single world), 1c1k R=52, Orbium x4 at 16 positions',
            'E': f'configs[4]: {total_worlds} worlds 64^3 in total, 1c1k R=13 spherical-shell kernel, uniform init'}[config]
    return f'{what}, {sim_steps} sim steps per bench step, 12 statistics + stop criteria every step' + \
        (', early stop OFF (Orbium at random toroidal shifts: all worlds survive)' if config == 'B' else '')


def total_worlds_of(config, scaling, n_gpus, n_override=0):
    base = n_override or (16 if config == 'C' else CONFIGS[config]['worlds'])
    if config == 'C':
        base *= 128
    return base * (n_gpus if (scaling == 'weak' or config == 'D') else 1)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='B', choices=sorted(CONFIGS))
    ap.add_argument('--scaling', default=None, choices=['strong', 'weak'],
                    help='strong: the workload BASELINE names split over the GPUs (default for B, C, E); weak: that workload per GPU (D: replicas)')
    ap.add_argument('--worlds', type=int, default=0, help='override the number of worlds of the workload (B, E) / solutions (C)')
    ap.add_argument('--sim-steps', type=int, default=0)
    ap.add_argument('--cpu-seconds', type=float, default=20.0, help='budget of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the secondary configs of the default run')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# workload descriptions shared by the B200 arm and the CPU arms (pure Python / NumPy, no oracle or engine import)
# ---------------------------------------------------------------------------------------------------------------------
ORBIUM_KP = [dict(k_slug='circle_2d', k_params=[1., [1.]], kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4', gf_params=[.15, .015],
                  h=1., c_in=0, c_out=0)]


def orbium_raw_cells():
    import numpy as np

    from leniax_b200 import loader, utils
    cfg = utils.load_config(os.path.join(ROOT, 'tests', 'golden', 'orbium.yaml'))
    return cfg, loader.load_raw_cells(cfg, use_init_cells=False).numpy().astype(np.float32)  # [1, 20, 20]


def make_worlds_numpy(n, seed):
    """n Orbium worlds at random toroidal shifts, float32 [n, 1, 128, 128] (host)."""
    import numpy as np
    cfg, raw = orbium_raw_cells()
    base = np.zeros((1, WS, WS), np.float32)
    base[:, 54:74, 54:74] = raw
    rng = np.random.default_rng(seed)
    shifts = rng.integers(0, WS, size=(n, 2))
    out = np.empty((n, 1, WS, WS), np.float32)
    for i in range(n):
        out[i] = np.roll(base, (int(shifts[i, 0]), int(shifts[i, 1])), axis=(1, 2))
    return cfg, out


def c3_kernels_params(n_sols, seed=2):
    """conf/config_qd_cmame_3c6k.yaml physics: kernels sorted by c_in, genotype (m, s, h) x 6 drawn U(0, 1) and scaled like lenia.py:131-143."""
    import numpy as np
    pairs = [(0, 0), (0, 1), (1, 1), (1, 2), (2, 2), (2, 0)]
    bs = {(0, 0): [1.], (1, 1): [.5, 1.], (2, 2): [1., .5]}
    base = [dict(k_slug='circle_2d', k_params=[1., bs.get(p, [1.])], kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4',
                 gf_params=[.17, .015], h=1., c_in=p[0], c_out=p[1]) for p in pairs]
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_sols):
        kp = copy.deepcopy(base)
        for k in kp:
            g = rng.random(3)
            k['gf_params'] = [round(.1 + .4 * g[0], 8), round(.005 + .095 * g[1], 8)]  # s kept > 0 for the timing run
            k['h'] = round(.05 + .95 * g[2], 8)
        out.append(kp)
    return out


def d_world_numpy(seed=4):
    import numpy as np
    size, scale = 2048, 4
    big = np.kron(orbium_raw_cells()[1][0], np.ones((scale, scale), np.float32))
    world = np.zeros((size, size), np.float32)
    rng = np.random.default_rng(seed)
    for _ in range(16):
        y, x = rng.integers(0, size - big.shape[0], 2)
        world[y:y + big.shape[0], x:x + big.shape[1]] = np.maximum(world[y:y + big.shape[0], x:x + big.shape[1]], big)
    return world


def sphere_kernel_numpy(R=13):
    """`circle_2d` formula (kernels.py:176-212) with the 3-D distance: 26^3 spherical shell, normalised."""
    import numpy as np
    k = int(np.ceil(R))
    ax = (np.arange(2 * k, dtype=np.float32) - k) / np.float32(R)
    d = np.sqrt(ax[:, None, None]**2 + ax[None, :, None]**2 + ax[None, None, :]**2).astype(np.float32)
    shell = (4 * (d % 1) * (1 - d % 1))**4
    kern = (d < 1) * shell
    return (kern / kern.sum()).astype(np.float32)[None]


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms: the oracle port on host cores, a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_setup(config):
    """-> (make_cells(n, seed) -> [n, C, *dims], K, gf, w, update_fn, stats_fn) of the oracle for one solution of `config`."""
    import numpy as np

    from oracle import lenia_oracle as lo
    if config == 'B':
        K, m = lo.get_kernels_and_mapping(copy.deepcopy(ORBIUM_KP), [WS, WS], 1, 13)
        wp, rp = {'R': 13, 'T': 10, 'nb_channels': 1}, {'world_size': [WS, WS]}
        make = lambda n, seed: make_worlds_numpy(n, seed)[1]  # noqa: E731
    elif config == 'C':
        K, m = lo.get_kernels_and_mapping(c3_kernels_params(1)[0], [WS, WS], 3, 13)
        wp, rp = {'R': 13, 'T': 10, 'nb_channels': 3}, {'world_size': [WS, WS]}

        def make(n, seed):  # smooth soups in [0, 0.45] like the perlin initial states (the sample only needs representative arithmetic)
            rng = np.random.default_rng(seed)
            coarse = rng.random((n, 3, 16, 16), dtype=np.float32)
            return (np.kron(coarse, np.ones((8, 8), np.float32)) * 0.45).astype(np.float32)
    elif config == 'D':
        K, m = lo.get_kernels_and_mapping(copy.deepcopy(ORBIUM_KP), [2048, 2048], 1, 52)
        wp, rp = {'R': 52, 'T': 10, 'nb_channels': 1}, {'world_size': [2048, 2048]}
        make = lambda n, seed: np.stack([d_world_numpy(seed + i)[None] for i in range(n)])  # noqa: E731
    else:
        kp = [dict(ORBIUM_KP[0], k_slug='raw', k_params=sphere_kernel_numpy())]
        K, m = lo.get_kernels_and_mapping(kp, [64, 64, 64], 1, 13)
        wp, rp = {'R': 13, 'T': 10, 'nb_channels': 1}, {'world_size': [64, 64, 64]}
        make = lambda n, seed: (np.random.default_rng(seed).random((n, 1, 64, 64, 64), dtype=np.float32) * 0.3).astype(np.float32)  # noqa: E731
    return make, K, m.get_gf_params(), m.get_kernels_weight_per_channel(), lo.build_update_fn(m), lo.build_compute_stats_fn(wp, rp)


def _cpu_worker(job):
    import numpy as np

    from oracle import lenia_oracle as lo
    config, n, steps, seed = job
    make, K, gf, w, upd, sfn = _cpu_setup(config)
    worlds = make(n, seed)
    t0 = time.perf_counter()
    stats, _ = lo.run_scan(worlds, K, gf, w, np.float32(10.), steps, upd, sfn, False)
    return time.perf_counter() - t0, float(stats['N'].sum())


def cpu_sample(config, seconds, cores):
    """The oracle on `cores` processes (one world batch each), sized for about `seconds` of wall time."""
    import multiprocessing as mp
    c = CONFIGS[config]
    probe_steps = 4 if config in 'DE' else 16
    dt, _ = _cpu_worker((config, 1, probe_steps, 7))  # calibrate on one core
    per_world_step = dt / probe_steps
    steps = int(max(4, min(c['sim_steps'], seconds / (2 * per_world_step))))
    wpp = int(max(1, min(64, seconds / (steps * per_world_step))))  # worlds per process
    jobs = [(config, wpp, steps, 100 + i) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context('spawn').Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)  # exclude interpreter start-up of the pool: slowest worker's own timer
    return cores * wpp * steps * c['cells'] / busy, f'{cores * wpp} worlds x {steps} steps of config {config}, {cores} process(es), wall {wall:.1f}s'


def run_reference(args):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    cores = os.cpu_count() or 1
    c = CONFIGS[args.config]
    scaling = args.scaling or ('weak' if args.config == 'D' else 'strong')
    vals, sample = [], ''
    per_step_budget = max(5.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, sample = cpu_sample(args.config, per_step_budget, cores)
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    line = {
        'impl': 'reference', 'metric': c['metric'], 'value': value, 'unit': 'cell-updates/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * (time.perf_counter() - t_all) / (args.warmup + args.steps), 'higher_is_better': True,
        'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': describe_workload(args.config, total_worlds_of(args.config, scaling, args.gpus, args.worlds),
                                                 args.sim_steps or c['sim_steps'])},
        'arm': 'the reference is Python on JAX, which is not installable in this image; this arm times the NumPy/scipy.fft restatement of it '
               '(oracle/, validated on the reference golden fixtures) on all host cores; each step = a bounded CPU sample of the workload '
               '(worlds are independent and the cost per world-step is constant, so the rate carries over)',
        'cpu_baseline': {'value': value, 'unit': 'cell-updates/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _loop(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
class Workload:
    """Everything one rank needs to run its share of a config: pinned host cells, device cells, per-solution parameters."""

    def __init__(self, config, scaling, rank, world, dev, n_override=0, sim_steps=0):
        import numpy as np
        import torch

        from leniax_b200 import distributed, helpers, initializations, kernels, statistics
        c = CONFIGS['B' if config == 'Bsearch' else config]
        self.config, self.scaling, self.c = config, scaling, c
        self.sim_steps = sim_steps or c['sim_steps']
        self.early_stop = False
        f32 = torch.float32
        if config in ('B', 'Bsearch'):
            total = n_override or c['worlds']
            a, b = distributed.shard_range(total, rank, world) if scaling == 'strong' else (0, total)
            self.total_worlds = total if scaling == 'strong' else total * world
            self.R, dims, C = 13, [WS, WS], 1
            K, mapping = kernels.get_kernels_and_mapping(copy.deepcopy(ORBIUM_KP), dims, C, self.R, device=dev)
            if config == 'B':
                cells = torch.from_numpy(make_worlds_numpy(total if scaling == 'strong' else b - a, seed=1 + (0 if scaling == 'strong' else rank))[1][a:b])
            else:  # the search form: perlin soups (most die early), early stop on
                cells = initializations.perlin(initializations.RngKey(1), total, dims, self.R, [.15, .015], device=dev)[1][a:b].cpu()
                self.early_stop = True
            self.host_cells = cells.reshape(1, b - a, C, *dims).contiguous().pin_memory()
            self.K, self.gf, self.w = K[None], mapping.get_gf_params(dev)[None], mapping.get_kernels_weight_per_channel(dev)[None]
            self.T = torch.tensor([10.], device=dev)
            self.shard_axis = 'inits'
            self.kernel = 'lnx_world128_tm'
            self.launches_per_step = 3  # lnx_prepare_kernel + the fused world kernel + lnx_summarize_kernel
        elif config == 'C':
            n_sols, n_init, C = n_override or 16, 128, 3
            a, b = distributed.shard_range(n_sols, rank, world) if scaling == 'strong' else (0, n_sols)
            self.total_worlds = n_sols * n_init * (1 if scaling == 'strong' else world)
            self.R, dims = 13, [WS, WS]
            kps = c3_kernels_params(n_sols, seed=2 if scaling == 'strong' else 2 + rank)
            Ks, gfs, ws, cells = [], [], [], []
            key = initializations.RngKey(2)
            mapping = None
            for s in range(n_sols):  # the key chain runs over all solutions so that a rank's worlds do not depend on the sharding
                key, noise = initializations.perlin(key, C * n_init, dims, self.R, kps[s][0]['gf_params'], device=dev)
                if a <= s < b:
                    K, mapping = kernels.get_kernels_and_mapping(kps[s], dims, C, self.R, device=dev)
                    Ks.append(K)
                    gfs.append(mapping.get_gf_params(dev))
                    ws.append(mapping.get_kernels_weight_per_channel(dev))
                    cells.append(noise.reshape(n_init, C, *dims))
            self.host_cells = torch.stack(cells).cpu().contiguous().pin_memory()
            self.K, self.gf, self.w = torch.stack(Ks), torch.stack(gfs), torch.stack(ws)
            self.T = torch.full((b - a, ), 10., device=dev)
            K = Ks[0]
            self.shard_axis = 'sols'
            self.kernel = 'lnx_world128_gen_tm'
            self.launches_per_step = 3
        elif config == 'D':
            self.total_worlds = world  # replicas only: the path does not shard a single world
            self.R, dims, C = 52, [2048, 2048], 1
            K, mapping = kernels.get_kernels_and_mapping(copy.deepcopy(ORBIUM_KP), dims, C, self.R, device=dev)
            self.host_cells = torch.from_numpy(d_world_numpy(4 + rank))[None, None, None].contiguous().pin_memory()
            self.K, self.gf, self.w = K[None], mapping.get_gf_params(dev)[None], mapping.get_kernels_weight_per_channel(dev)[None]
            self.T = torch.tensor([10.], device=dev)
            self.shard_axis = 'sols'
            self.kernel = 't2k::lead_kernel + t2k::rows_inv_kernel (graph-replayed step)'
            self.launches_per_step = 2 * self.sim_steps + 6  # per sim step lead + fused rows; + table gathers, first rows, pass D, summary
        else:
            total = n_override or c['worlds']
            a, b = distributed.shard_range(total, rank, world) if scaling == 'strong' else (0, total)
            self.total_worlds = total if scaling == 'strong' else total * world
            self.R, dims, C = 13, [64, 64, 64], 1
            kern = torch.from_numpy(sphere_kernel_numpy(self.R)).to(dev)
            kp = [dict(ORBIUM_KP[0], k_slug='raw', k_params=kern)]
            K, mapping = kernels.get_kernels_and_mapping(kp, dims, C, self.R, device=dev)
            cells = initializations.random_uniform(initializations.RngKey(5), total, dims, self.R, [.15, .015], device=dev)[1][a:b]
            self.host_cells = cells.reshape(1, b - a, C, *dims).cpu().contiguous().pin_memory()
            self.K, self.gf, self.w = K[None], mapping.get_gf_params(dev)[None], mapping.get_kernels_weight_per_channel(dev)[None]
            self.T = torch.tensor([10.], device=dev)
            self.shard_axis = 'inits'
            self.kernel = 't64h::lead_h_kernel + t64h::plane_step_kernel (+ pass D) per step'
            self.launches_per_step = 3 * self.sim_steps + 4
        self.dims, self.C = dims, C
        self.dev_cells = self.host_cells.to(dev)
        self.ufn = helpers.build_update_fn(K.shape, mapping)
        self.sfn = statistics.build_compute_stats_fn({'R': self.R, 'T': 10}, {'world_size': dims})
        self.local_worlds = self.host_cells.shape[0] * self.host_cells.shape[1]
        self.cell_updates_local = self.local_worlds * int(np.prod(dims)) * self.sim_steps
        self.cell_updates_total = self.total_worlds * int(np.prod(dims)) * self.sim_steps

    def step(self, cells):
        """One batched scan through the reference-facing sharded entry point -> gathered [N_sols, N_init, 12] summary."""
        from leniax_b200 import distributed
        summary, _, _ = distributed.run_scan_mem_optimized_sharded(None, cells, self.K, self.gf, self.w, self.T, self.sim_steps, self.R, self.ufn,
                                                                   self.sfn, early_stop=self.early_stop, sharded_inputs=self.shard_axis)
        return summary

    def kernel_only_ms(self, reps):
        """The scan launch(es) alone: no table prepare, no torch glue; CUDA events on the launching (current) stream."""
        import torch

        import leniax_b200
        from leniax_b200 import _lib
        lib = leniax_b200.load_library()
        plan = [p for p in leniax_b200.engine.Plan._cache.values()  # the plan the runner built for this workload (the latest that matches)
                if tuple(p.key[0]) == tuple(self.dims) and p.desc.nb_channels == self.C and p.device == self.dev_cells.device][-1]
        run_flags = _lib.LNX_RUN_ASSUME_FINITE | (_lib.LNX_RUN_WEIGHTS_MATCH_COUT if plan.desc.c_out[0] != _lib.LNX_COUT_ANY else 0)
        if self.config == 'C':
            self.kernel = {'generic2': 'lnx_world128_gen2', 'generic': 'lnx_world128_gen_tm'}[plan.variant(False)]
        n_sols, n_init = self.dev_cells.shape[:2]
        dev, f32 = self.dev_cells.device, torch.float32
        table = plan.prepare_kernels(self.K.reshape((n_sols, -1) + tuple(self.dims)), n_sols)
        stats = torch.empty((_lib.LNX_NB_STATS, n_sols, self.sim_steps, n_init), dtype=f32, device=dev)
        cm = torch.empty((n_sols, self.sim_steps, n_init, self.C), dtype=f32, device=dev)
        na = torch.empty((n_sols, n_init), dtype=f32, device=dev)
        ws = torch.empty(int(lib.lnx_workspace_bytes_for(plan.handle, n_sols, n_init)), dtype=torch.uint8, device=dev)
        dt = (1. / self.T).contiguous()
        gf, w = self.gf.reshape(n_sols, -1, 2).contiguous(), self.w.reshape(n_sols, self.C, -1).contiguous()
        stream = torch.cuda.current_stream().cuda_stream

        def launch():
            _lib.check(lib.lnx_run_scan(plan.handle, n_sols, n_init, self.sim_steps, run_flags, self.dev_cells.data_ptr(),
                                        table.data_ptr(), gf.data_ptr(), w.data_ptr(), dt.data_ptr(), stats.data_ptr(), cm.data_ptr(), na.data_ptr(),
                                        None, None, None, None, ws.data_ptr(), ws.numel(), stream))

        launch()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(reps):
            launch()
        k1.record()
        torch.cuda.synchronize()
        return k0.elapsed_time(k1) / reps


def measure(wl, steps, warmup, world, dev, dist):
    """-> (ms of `steps` resident steps, ms of `steps` end-to-end steps, last gathered block on the host), max over ranks."""
    import torch

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, w):
        out = None
        for _ in range(w):
            # (the previous result stays alive while the next call allocates, exactly as in the timed loop: the caching allocator then
            # owns both generations of output blocks and no cudaMalloc - an implicit device synchronisation - falls into the timed region)
            out = fn()  # noqa: F841
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    ms, _ = timed(lambda: wl.step(wl.dev_cells), steps, warmup)
    # end to end: the public entry point takes the pinned HOST tensor (H2D inside the call and inside the timed region: for one
    # solution with many initialisations the first wave of worlds at once and the rest on a copy stream under that wave's compute),
    # and the gathered block is read back to the host every step
    ms_e2e, block_h = timed(lambda: wl.step(wl.host_cells).cpu(), steps, max(1, warmup - 2))
    return ms, ms_e2e, block_h


def roofline_of(wl, kernel_ms, value_local, fp32_peak, peaks):
    c = wl.c
    if c['bound'] == 'fp32':
        achieved = value_local * c['flop'] / 1e12
        r = {'bound': 'fp32', 'achieved': achieved, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': achieved / fp32_peak,
             'flop_per_cell_update': c['flop'], 'peak_source': 'measured live: lnx_measure_fp32_peak FMA loop (MEASURED_PEAKS.json has no FP32 entry)',
             'peak_analytic_tflops': FP32_PEAK_ANALYTIC_TFLOPS, 'frac_of_analytic': achieved / FP32_PEAK_ANALYTIC_TFLOPS}
    else:
        hbm = peaks.get('hbm_gbs', 6650.0)
        achieved = value_local * c['bytes'] / 1e9
        r = {'bound': 'hbm', 'achieved': achieved, 'peak': hbm, 'unit': 'GB/s', 'frac': achieved / hbm, 'bytes_per_cell_update': c['bytes'],
             'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback of B200_PROFILING.md',
             'note': 'algorithmic bytes (SURVEY 8d convention: three streaming passes); the 48 MiB working set of one 2048^2 world sits in L2'}
    if 'bytes' in c and c['bound'] == 'fp32':  # config E: multi-pass engine, also report the HBM convention
        # engine_*: what the two passes of this engine really stream per cell-update (lead: half spectrum in + out = 2 x 4.125 B; plane_step:
        # spectrum in + out, state in + out = 16.25 B; DRAM traffic measured by ncu equals it, profiles/r2_t64h_ncu_v2.txt)
        r['hbm_convention'] = {'bytes_per_cell_update': c['bytes'], 'achieved_gbs': value_local * c['bytes'] / 1e9, 'peak_gbs': peaks.get('hbm_gbs'),
                               'engine_bytes_per_cell_update': 24.5, 'engine_achieved_gbs': value_local * 24.5 / 1e9,
                               'engine_frac': value_local * 24.5 / 1e9 / peaks['hbm_gbs'] if peaks.get('hbm_gbs') else None}
    r['kernel'], r['kernel_ms'] = wl.kernel, kernel_ms
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at this exact shape, from a committed ncu capture (or null)
    try:
        table = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        hit = table.get(f'{wl.config}|{wl.local_worlds}x{wl.sim_steps}')
        r['traffic'], r['traffic_source'] = (hit['bytes'], hit['source']) if hit else (None, None)
    except Exception:
        r['traffic'] = None
    return r


def run_b200(args):
    import torch
    import torch.distributed as dist

    import leniax_b200
    from leniax_b200 import _lib

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py (b200 arm) needs a GPU: leniax_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = leniax_b200.load_library()
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        peaks = {}
    scaling = args.scaling or ('weak' if args.config == 'D' else 'strong')

    def fp32_peak():
        tf, pk_ms = _lib.ctypes.c_double(), _lib.ctypes.c_double()
        _lib.check(lib.lnx_measure_fp32_peak(4096, _lib.ctypes.byref(tf), _lib.ctypes.byref(pk_ms), torch.cuda.current_stream().cuda_stream))
        return float(tf.value)

    def run_config(config, scal, steps, warmup, with_clocks=False, n_override=0, sim_steps=0):
        wl = Workload(config, scal, rank, world, dev, n_override, sim_steps)
        sampler = ClockSampler(local) if (rank == 0 and with_clocks) else None
        if sampler:
            sampler.start()
        ms, ms_e2e, block_h = measure(wl, steps, warmup, world, dev, dist)
        clocks = sampler.stop() if sampler else None
        resident = config in ('B', 'Bsearch', 'C')
        kernel_ms = wl.kernel_only_ms(steps) if resident and not wl.early_stop else ms / steps
        value = wl.cell_updates_total * steps / (ms * 1e-3)
        res = {'value': value, 'ms_per_step': ms / steps, 'e2e_value': wl.cell_updates_total * steps / (ms_e2e * 1e-3), 'e2e_ms_per_step': ms_e2e / steps,
               'h2d': int(wl.host_cells.numel() * 4) * world, 'd2h': int(block_h.numel() * 4), 'clocks': clocks, 'block': block_h, 'wl': wl,
               'kernel_ms': kernel_ms, 'value_local_kernel': wl.cell_updates_local / (kernel_ms * 1e-3)}
        return res

    def describe(wl, scal):
        return describe_workload(wl.config, wl.total_worlds, wl.sim_steps)

    primary = run_config(args.config, scaling, args.steps, args.warmup, with_clocks=True, n_override=args.worlds, sim_steps=args.sim_steps)
    wl = primary['wl']
    peak = fp32_peak()
    line = None
    if rank == 0:
        block_h = primary['block']
        line = {
            'metric': wl.c['metric'], 'value': primary['value'], 'unit': 'cell-updates/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': primary['ms_per_step'], 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': describe(wl, scaling)},
            'arm': {
                'per_gpu': f'{wl.local_worlds} world(s) on rank 0',
                'parallelism': f'worlds sharded over {world} GPU(s) ({wl.shard_axis} axis, no data-path collective), one NCCL all-gather of the '
                               '[worlds, 12] fitness / behaviour block per step',
                'cache': 'resident kernels keep state and spectra on-chip (tensor memory + shared memory); inputs are read from HBM once per '
                         'step (B: 268 MB > 126 MB L2 at 4096 worlds)' if wl.c['bound'] == 'fp32' and wl.config != 'E' else
                         'multi-pass engine: the working set is streamed every sim step (E: 256 MB of state > L2; D: 48 MiB, L2-resident by design)',
            },
            'e2e': {'value': primary['e2e_value'], 'unit': 'cell-updates/s', 'h2d_bytes_per_step': primary['h2d'], 'd2h_bytes_per_step': primary['d2h'],
                    'ms_per_step': primary['e2e_ms_per_step']},
            'gpu_launches': wl.launches_per_step * args.steps,
            'clocks': primary['clocks'],
            'roofline': roofline_of(wl, primary['kernel_ms'], primary['value_local_kernel'], peak, peaks),
            'checks': {'mean_N': float(block_h[..., 0].mean().item()), 'all_alive': bool((block_h[..., 0] == wl.sim_steps).all().item()),
                       'mean_mass': float(block_h[..., 1].mean().item())},
        }
        if wl.c['bound'] == 'fp32' and wl.config != 'E':
            line['roofline']['hbm_achieved_gbs'] = 48.0 * wl.local_worlds * wl.sim_steps / (primary['kernel_ms'] * 1e-3) / 1e9
            line['roofline']['hbm_peak_gbs'] = peaks.get('hbm_gbs')

    # ---- secondary measurements of the default run (same JSON line, key `secondary`) ----
    if args.config == 'B' and not args.no_secondary and not args.worlds and not args.sim_steps:
        secondary = {}

        def add(name, config, scal, steps, warmup):
            try:
                r = run_config(config, scal, steps, warmup)
                w2 = r['wl']
                if rank == 0:
                    secondary[name] = {'workload': describe(w2, scal), 'scaling': scal, 'value': r['value'], 'unit': 'cell-updates/s',
                                       'ms_per_step': r['ms_per_step'], 'e2e_value': r['e2e_value'],
                                       'roofline': roofline_of(w2, r['kernel_ms'], r['value_local_kernel'], peak, peaks),
                                       'mean_N': float(r['block'][..., 0].mean().item())}
                del r, w2
            except Exception as e:  # a secondary config never takes the headline line down
                if rank == 0:
                    secondary[name] = {'error': f'{type(e).__name__}: {e}'}
            torch.cuda.empty_cache()

        if world > 1:
            add('B_weak', 'B', 'weak', 3, 2)
        add('C', 'C', 'strong', 2, 2)
        if world == 1:
            add('B_search_early_stop', 'Bsearch', 'strong', 3, 2)
            add('D', 'D', 'weak', 3, 3)
            add('E', 'E', 'strong', 3, 3)
        if rank == 0:
            line['secondary'] = secondary

    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            v, sample = cpu_sample(args.config, args.cpu_seconds, 1)
            line['cpu_baseline'] = {'value': v, 'unit': 'cell-updates/s', 'cores': 1, 'kind': 'port', 'sample': sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
