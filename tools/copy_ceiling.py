#!/usr/bin/env python
"""What the BOX can move: concurrent pinned host<->device copies on N GPUs.

The end-to-end leg of bench.py moves, per exposure and per GPU, the NSAMP reads
to the host (126 MB at configs[3]) and -- when the caller hands the planet signal
over as a dense host array -- 135 MB to the device.  Whether 8 GPUs can do that
at once is a property of the host (PCIe root complexes, memory controllers, the
VM), not of the kernels.  This tool measures it with nothing else running:

  * one process per GPU (torchrun, same NUMA binding as bench.py),
  * per rank a D2H stream copying a 126 MB device buffer into pinned memory and
    an H2D stream copying a 135 MB pinned buffer to the device, back to back,
    `--seconds` long, one cudaMemcpyAsync per buffer,
  * three modes: d2h only, h2d only, both at once.

Prints one JSON line (rank 0): per-mode aggregate GB/s and the exposures/s
ceiling they imply for the two e2e forms of bench.py.

    python tools/copy_ceiling.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/copy_ceiling.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def measure(dev, world=1, seconds=2.0, d2h_bytes=125829120, h2d_bytes=134873088):
    """Aggregate GB/s of concurrent pinned copies on `world` GPUs (one process per GPU; every rank
    calls this; torch.distributed must be initialised when world > 1).  Returns, on every rank,
    {'aggregate': {mode: {'d2h_gbs', 'h2d_gbs'}}, 'ceiling_exposures_per_s': {...}}."""
    n_d2h, n_h2d = int(d2h_bytes), int(h2d_bytes)
    d_out = torch.empty(n_d2h, dtype=torch.uint8, device=dev)
    h_out = [torch.empty(n_d2h, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    h_in = torch.empty(n_h2d, dtype=torch.uint8, pin_memory=True)
    d_in = [torch.empty(n_h2d, dtype=torch.uint8, device=dev) for _ in range(2)]
    s_down, s_up = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def run(d2h, h2d, secs):
        """Copies issued for `secs`, at most 3 in flight per stream and never an empty queue;
        returns (bytes down, bytes up, elapsed s) of this rank."""
        barrier()
        t0 = time.perf_counter()
        k = down = up = 0
        ev_down, ev_up = [], []
        while time.perf_counter() - t0 < secs:
            if d2h:
                if len(ev_down) >= 3:
                    ev_down.pop(0).synchronize()
                with torch.cuda.stream(s_down):
                    h_out[k & 1].copy_(d_out, non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(s_down)
                ev_down.append(e)
                down += n_d2h
            if h2d:
                if len(ev_up) >= 3:
                    ev_up.pop(0).synchronize()
                with torch.cuda.stream(s_up):
                    d_in[k & 1].copy_(h_in, non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(s_up)
                ev_up.append(e)
                up += n_h2d
            k += 1
        torch.cuda.synchronize(dev)
        el = time.perf_counter() - t0
        return down, up, el

    out = {}
    for name, d2h, h2d in (('d2h_only', True, False), ('h2d_only', False, True), ('both', True, True)):
        run(d2h, h2d, min(seconds, 0.3))      # warm-up
        down, up, el = run(d2h, h2d, seconds)
        t = torch.tensor([down / el, up / el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        out[name] = {'d2h_gbs': float(t[0]) / 1e9, 'h2d_gbs': float(t[1]) / 1e9}
    both = out['both']
    return {
        'aggregate': out,
        # exposures/s the box could move if the GPUs were infinitely fast
        'ceiling_exposures_per_s': {
            'dense_host_signal (reads down + planet signal up, concurrent)':
                min(both['d2h_gbs'] * 1e9 / n_d2h, both['h2d_gbs'] * 1e9 / n_h2d),
            'factored_signal (reads down only)': out['d2h_only']['d2h_gbs'] * 1e9 / n_d2h,
        },
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seconds', type=float, default=2.0)
    ap.add_argument('--d2h-mb', type=float, default=125.829120)      # 15 x 1024 x 1024 float64
    ap.add_argument('--h2d-mb', type=float, default=134.873088)      # 4116 x 4096 float64
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    from wayne_b200.engine import bind_to_gpu_numa_node
    bound = bind_to_gpu_numa_node(local) if world > 1 else False
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # stdout carries the JSON line only
        dist.init_process_group('nccl', device_id=dev)
    res = measure(dev, world, args.seconds, args.d2h_mb * 1e6, args.h2d_mb * 1e6)
    if rank == 0:
        line = {'tool': 'copy_ceiling', 'n_gpus': world, 'numa_bound': bool(bound), 'seconds': args.seconds,
                'd2h_mb': args.d2h_mb, 'h2d_mb': args.h2d_mb}
        line.update(res)
        line['host'] = {'cpus': os.cpu_count(), 'gpu': torch.cuda.get_device_name(local)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
