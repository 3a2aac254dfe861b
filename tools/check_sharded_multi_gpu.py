#!/usr/bin/env python
"""Real-NCCL check of the sharded entry point (leniax_b200.distributed.run_scan_mem_optimized_sharded) on N GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/check_sharded_multi_gpu.py
Every rank simulates its slice of a (2 solutions x 7 initialisations) batch; the all-gathered summary must equal the summary of
an unsharded run on every rank (worlds are independent, so the rows are bit-identical)."""
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leniax_b200 import distributed as lnx_dist, helpers, qd, runner, statistics, utils  # noqa: E402

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
cfg = utils.load_config(os.path.join(ROOT, 'tests', 'golden', 'orbium-test.yaml'))
cells, K, mapping = helpers.init(copy.deepcopy(cfg), device=dev)
wp = cfg['world_params']
ufn = helpers.build_update_fn(K.shape, mapping)
sfn = statistics.build_compute_stats_fn(wp, cfg['render_params'])
gf, w = mapping.get_gf_params(dev), mapping.get_kernels_weight_per_channel(dev)
worlds = torch.stack([torch.roll(cells[0], (9 * i, 4 * i), dims=(1, 2)) * (1. if i % 3 else 0.3) for i in range(14)]).reshape(2, 7, 1, 128, 128)
args = (worlds, torch.stack([K, K]), torch.stack([gf, gf]), torch.stack([w, w]), torch.tensor([10., 10.], device=dev))
steps = 150
summary, keys, _ = lnx_dist.run_scan_mem_optimized_sharded(None, *args, steps, 13, ufn, sfn)
stats, _ = runner.run_scan_mem_optimized(None, *args, steps, 13, ufn, sfn)
ref, rkeys = qd.summarize_stats(stats)
ok = keys == rkeys and torch.equal(summary, ref)
# the same batch with every rank holding ONLY its part (ragged: 7 initialisations over N ranks / 2 solutions over N ranks), plus a
# 3-channel 6-kernel batch (lnx_world128_gen2) split by solutions
rank, world = dist.get_rank(), dist.get_world_size()
i0, i1 = lnx_dist.shard_range(7, rank, world)
by_inits, _, _ = lnx_dist.run_scan_mem_optimized_sharded(None, worlds[:, i0:i1], *args[1:], steps, 13, ufn, sfn, sharded_inputs='inits')
s0, s1 = lnx_dist.shard_range(2, rank, world)
by_sols, _, _ = lnx_dist.run_scan_mem_optimized_sharded(None, *[a[s0:s1] for a in args], steps, 13, ufn, sfn, sharded_inputs='sols')
ok = ok and torch.equal(by_inits, ref) and torch.equal(by_sols, ref)
import bench  # noqa: E402
from leniax_b200 import initializations, kernels  # noqa: E402
kps = bench.c3_kernels_params(world)
Ks, maps = kernels.get_kernels_and_mapping_batch(copy.deepcopy(kps), [128, 128], 3, 13, device=dev)
_, c3 = initializations.perlin_batch([initializations.RngKey(50 + i) for i in range(world)], 12, [128, 128], 13, [kp[0]['gf_params'] for kp in kps], device=dev)
c3 = c3.reshape(world, 4, 3, 128, 128)
ufn3 = helpers.build_update_fn(Ks[0].shape, maps[0])
sfn3 = statistics.build_compute_stats_fn({'R': 13, 'T': 10}, {'world_size': [128, 128]})
a3 = (c3, Ks, torch.stack([m.get_gf_params(dev) for m in maps]), torch.stack([m.get_kernels_weight_per_channel(dev) for m in maps]),
      torch.full((world, ), 10., device=dev))
ref3 = qd.summarize_stats(runner.run_scan_mem_optimized(None, *a3, 60, 13, ufn3, sfn3)[0])[0]
got3, _, _ = lnx_dist.run_scan_mem_optimized_sharded(None, *[a[rank:rank + 1] for a in a3], 60, 13, ufn3, sfn3, sharded_inputs='sols')
ok = ok and torch.equal(got3, ref3)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if dist.get_rank() == 0:
    print('sharded == unsharded on all %d ranks: %s; N = %s' % (dist.get_world_size(), bool(flag.item()), summary[..., 0].tolist()))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
