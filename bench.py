#!/usr/bin/env python
"""bench.py — cell-updates/s of the batched 128x128 Orbium search (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU restatement of leniax on host cores

One "step" = one `runner.run_scan_mem_optimized` call over the whole batch: 4096 Orbium worlds (1 channel, 1 kernel,
128x128, R=13, T=10) x 1024 simulation steps with all 12 statistics and the stop criteria.  Worlds are the Orbium of
conf/species/2d/1c-1k/orbium.yaml at random toroidal shifts, so every world survives all 1024 steps and no work is
skipped (early stop is OFF).  Weak scaling: every rank runs its own 4096 worlds; after each step the ranks all-gather
the [worlds, 13] fitness/behaviour block the QD archive consumes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORLDS_PER_GPU = 4096
SIM_STEPS = 1024
WS = 128
FLOP_PER_CELL_UPDATE = 110.0  # rfft2 + irfft2 + spectrum product + growth + update + statistics, 1c1k 128^2
FP32_PEAK_ANALYTIC_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.45 at the 1965 MHz max clock


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--worlds', type=int, default=WORLDS_PER_GPU, help='worlds per GPU (default: the BASELINE config)')
    ap.add_argument('--sim-steps', type=int, default=SIM_STEPS)
    ap.add_argument('--cpu-seconds', type=float, default=20.0, help='budget of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------------
def orbium_config():
    from oracle import lenia_oracle as lo  # config decoding only; used by the CPU arms
    return lo.load_yaml_config(os.path.join(ROOT, 'tests', 'golden', 'orbium.yaml'))


def make_worlds_numpy(n, seed):
    """n Orbium worlds at random toroidal shifts, float32 [n, 1, 128, 128] (host)."""
    import numpy as np

    from leniax_b200 import loader, utils
    cfg = utils.load_config(os.path.join(ROOT, 'tests', 'golden', 'orbium.yaml'))
    raw = loader.load_raw_cells(cfg, use_init_cells=False).numpy()  # [1, 20, 20]
    base = np.zeros((1, WS, WS), np.float32)
    base[:, 54:74, 54:74] = raw
    rng = np.random.default_rng(seed)
    shifts = rng.integers(0, WS, size=(n, 2))
    out = np.empty((n, 1, WS, WS), np.float32)
    for i in range(n):
        out[i] = np.roll(base, (int(shifts[i, 0]), int(shifts[i, 1])), axis=(1, 2))
    return cfg, out


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) — bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import numpy as np

    from oracle import lenia_oracle as lo
    worlds, steps = args
    cfg = orbium_config()
    K, mapping = lo.get_kernels_and_mapping(cfg['kernels_params'], [WS, WS], 1, cfg['world_params']['R'])
    upd = lo.build_update_fn(mapping)
    sfn = lo.build_compute_stats_fn(cfg['world_params'], cfg['render_params'])
    t0 = time.perf_counter()
    stats, final = lo.run_scan(worlds, K, mapping.get_gf_params(), mapping.get_kernels_weight_per_channel(), np.float32(10.), steps,
                               upd, sfn, False)
    return time.perf_counter() - t0, float(stats['N'].sum())


def cpu_sample(seconds, cores):
    """Run the oracle on `cores` processes (one world batch each) sized for about `seconds` of wall time."""
    import multiprocessing as mp
    _, probe = make_worlds_numpy(1, seed=7)
    dt, _ = _cpu_worker((probe, 32))  # calibrate on one core: 1 world x 32 steps
    per_world_step = dt / 32
    steps = int(max(64, min(SIM_STEPS, seconds / (2 * per_world_step))))
    wpp = int(max(2, min(64, seconds / (steps * per_world_step))))  # worlds per process
    _, worlds = make_worlds_numpy(cores * wpp, seed=7)
    jobs = [(worlds[wpp * i:wpp * (i + 1)], steps) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context('spawn').Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)  # exclude interpreter start-up of the pool: slowest worker's own timer
    cell_updates = cores * wpp * steps * WS * WS
    return cell_updates / busy, f'{cores * wpp} worlds x {steps} steps of the same Orbium batch, {cores} process(es), wall {wall:.1f}s'


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
# reference arm
# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    sample = ''
    per_step_budget = max(5.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, sample = cpu_sample(per_step_budget, cores)
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    line = {
        'impl': 'reference', 'metric': 'cell-updates/sec (batched 128x128 Orbium search, stats on)', 'value': value, 'unit': 'cell-updates/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * (time.perf_counter() - t_all) / (args.warmup + args.steps), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'configs[1]: Orbium 1c1k 128x128 R=13 T=10, stats + stop criteria; each step = bounded CPU sample of it',
                   'note': 'the reference is Python on JAX, which is not installable in this image; this arm times the NumPy/scipy.fft '
                           'restatement of it (oracle/, validated on the reference golden fixtures) on all host cores'},
        'cpu_baseline': {'value': value, 'unit': 'cell-updates/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import leniax_b200
    from leniax_b200 import _lib, helpers, kernels, qd, runner, statistics

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
    variant_flag = 0
    kernel_name = 'lnx_world128_tm'

    n_worlds, sim_steps = args.worlds, args.sim_steps
    cfg, worlds_np = make_worlds_numpy(n_worlds, seed=1 + rank)
    wp = cfg['world_params']
    K, mapping = kernels.get_kernels_and_mapping(cfg['kernels_params'], [WS, WS], 1, wp['R'], device=dev)
    ufn = helpers.build_update_fn(K.shape, mapping, wp.get('get_state_fn_slug', 'v1'), wp.get('weighted_average', True), True)
    sfn = statistics.build_compute_stats_fn(wp, cfg['render_params'])
    gf, w = mapping.get_gf_params(dev)[None], mapping.get_kernels_weight_per_channel(dev)[None]
    T = torch.tensor([float(wp['T'])], device=dev)
    Kb = K[None]
    host_cells = torch.from_numpy(worlds_np)[None].pin_memory()  # [1, n, 1, 128, 128] pinned host memory
    dev_cells = host_cells.to(dev)
    cell_updates_per_rank = n_worlds * WS * WS * sim_steps

    def gather_block(stats):
        """What the QD archive consumes per world (qd.py:168-186): N + mean of the last 128 rows of each statistic."""
        block = qd.summarize_stats(stats)[0][0].contiguous()  # [worlds, 12], lnx_summarize_stats on the device
        if world > 1:
            out = [torch.empty_like(block) for _ in range(world)]
            dist.all_gather(out, block)
            block = torch.cat(out)
        return block

    def step_resident():
        stats, final = runner.run_scan_mem_optimized(None, dev_cells, Kb, gf, w, T, sim_steps, wp['R'], ufn, sfn)
        return gather_block(stats)

    def step_e2e():
        # the public API takes the pinned HOST tensor: the H2D copy happens inside the call (first wave of worlds at once,
        # the rest on a copy stream under that wave's compute) and inside the timed region
        stats, final = runner.run_scan_mem_optimized(None, host_cells, Kb, gf, w, T, sim_steps, wp['R'], ufn, sfn)
        return gather_block(stats).cpu()  # D2H of the step's result

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, block = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e, block_h = timed(step_e2e, args.steps, max(1, args.warmup - 2))

    # dominant kernel alone (no table prepare, no torch glue), CUDA events on the launching (current) stream
    plan = next(iter(leniax_b200.engine.Plan._cache.values()))
    table = plan.prepare_kernels(Kb.reshape(1, 1, WS, WS), 1)
    f32 = torch.float32
    stats_buf = torch.empty((_lib.LNX_NB_STATS, 1, sim_steps, n_worlds), dtype=f32, device=dev)
    cm_buf = torch.empty((1, sim_steps, n_worlds, 1), dtype=f32, device=dev)
    n_buf = torch.empty((1, n_worlds), dtype=f32, device=dev)
    ws_buf = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
    dt = (1. / T).contiguous()
    stream = torch.cuda.current_stream().cuda_stream

    def launch_kernel():
        _lib.check(lib.lnx_run_scan(plan.handle, 1, n_worlds, sim_steps, _lib.LNX_RUN_ASSUME_FINITE | variant_flag, dev_cells.data_ptr(), table.data_ptr(),
                                    gf.data_ptr(), w.data_ptr(), dt.data_ptr(), stats_buf.data_ptr(), cm_buf.data_ptr(), n_buf.data_ptr(),
                                    None, None, None, None, ws_buf.data_ptr(), ws_buf.numel(), stream))

    launch_kernel()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        launch_kernel()
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps

    tf, pk_ms = _lib.ctypes.c_double(), _lib.ctypes.c_double()
    _lib.check(lib.lnx_measure_fp32_peak(4096, _lib.ctypes.byref(tf), _lib.ctypes.byref(pk_ms), stream))
    fp32_peak = float(tf.value)

    if rank == 0:
        total_cu = cell_updates_per_rank * world
        value = total_cu * args.steps / (ms * 1e-3)
        e2e_value = total_cu * args.steps / (ms_e2e * 1e-3)
        achieved_tflops = cell_updates_per_rank * FLOP_PER_CELL_UPDATE / (kernel_ms * 1e-3) / 1e12
        stats_bytes = (12 * 4) * n_worlds * sim_steps  # HBM traffic of the resident kernel: statistics rows only
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            peaks = {}
        line = {
            'metric': 'cell-updates/sec (batched 128x128 Orbium search, stats on)', 'value': value, 'unit': 'cell-updates/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {
                'workload': f'configs[1]: {n_worlds} Orbium worlds per GPU, 1c1k 128x128 R=13 T=10, {sim_steps} sim steps per bench step, '
                            '12 statistics + stop criteria every step, early stop OFF (Orbium at random toroidal shifts: all worlds survive)',
                'parallelism': f'worlds sharded over {world} GPU(s), one NCCL all-gather of [worlds,12] per step',
                'cache': 'state/spectra are on-chip resident (tensor memory + shared memory) by design; inputs 268 MB per GPU (> 126 MB L2), read once per step',
            },
            'e2e': {'value': e2e_value, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': int(host_cells.numel() * 4) * world,
                    'd2h_bytes_per_step': int(block_h.numel() * 4), 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': 3 * args.steps,  # per step: lnx_prepare_kernel + the fused world kernel + lnx_summarize_kernel
            'clocks': clocks,
            'roofline': {
                'bound': 'fp32', 'achieved': achieved_tflops, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': achieved_tflops / fp32_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at this exact workload, from an ncu
                # capture (profiles/r1_tm_dram_traffic_bench_launch.csv); algorithmic bytes = 268 MB state in + 218 MB rows out
                'traffic': 762075392 if (kernel_name == 'lnx_world128_tm' and n_worlds == 4096 and sim_steps == 1024) else None,
                'kernel': kernel_name, 'kernel_ms': kernel_ms,
                'flop_per_cell_update': FLOP_PER_CELL_UPDATE, 'peak_source': 'measured live: lnx_measure_fp32_peak FMA loop (MEASURED_PEAKS.json has no FP32 entry)',
                'peak_analytic_tflops': FP32_PEAK_ANALYTIC_TFLOPS, 'frac_of_analytic': achieved_tflops / FP32_PEAK_ANALYTIC_TFLOPS,
                'hbm_achieved_gbs': stats_bytes / (kernel_ms * 1e-3) / 1e9, 'hbm_peak_gbs': peaks.get('hbm_gbs'),
            },
            'checks': {'all_alive': bool((block_h[:, 0] == sim_steps).all().item()), 'mean_mass': float(block_h[:, 1].mean().item())},
        }
        if not args.no_cpu_baseline and world == 1:
            v, sample = cpu_sample(args.cpu_seconds, 1)
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
