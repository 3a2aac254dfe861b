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
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if dist.get_rank() == 0:
    print('sharded == unsharded on all %d ranks: %s; N = %s' % (dist.get_world_size(), bool(flag.item()), summary[..., 0].tolist()))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
