"""BASELINE.json configs[4], second half: the Ulysses all-to-all at the sizes of the path (payload = one rank's
[L/P, 40, 128] bf16 q / k / v / out tensor; L in {32 760, 75 600}; P = world size), timed three ways on the device:
  nccl_raw    : torch.distributed.all_to_all_single on an already packed buffer (the collective alone)
  nccl_path   : what the training path runs: prfl_a2a_pack (one staging kernel) + all_to_all_single
  p2p_scatter : what the no-grad forward runs: prfl_a2a_scatter_p2p (peer stores over NVLink straight into the owners'
                receive buffers, no staging, no NCCL) + the cross-rank signal-pad barrier
Algorithmic bus bandwidth per rank = payload * (P-1)/P / time, against 900 GB/s per direction (NVLink 5).
Launch: python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 tools/a2a_sweep.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prfl_b200 import ops, parallel  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # max over ranks
    return float(t)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P = dist.get_world_size()
    parallel.initialize_sequence_parallel_state(P)
    H, rows = 40, []
    for L in (32760, 75600):
        Ll = L // P
        x = torch.randn(Ll, H, 128, device=dev).bfloat16()
        payload = x.numel() * 2
        packed = torch.empty(P, Ll, H // P, 128, dtype=torch.bfloat16, device=dev)
        out = torch.empty_like(packed)
        t_raw = timeit(lambda: dist.all_to_all_single(out, packed, group=parallel.nccl_info.group))
        t_path = timeit(lambda: parallel.ulysses_scatter_tokens(x, P))
        row = {"L": L, "P": P, "payload_MiB": payload / 2**20, "nccl_raw_ms": t_raw, "nccl_path_ms": t_path,
               "nccl_raw_busGBps": payload * (P - 1) / P / t_raw / 1e6, "nccl_path_busGBps": payload * (P - 1) / P / t_path / 1e6}
        p2p = parallel.get_p2p_ulysses(L, H, dev)
        if p2p is not None:
            def scatter():
                ops.a2a_scatter_p2p(x, p2p.qkv_ptrs, P, p2p.rank)
                p2p.h_qkv.barrier(channel=0)
            t_p2p = timeit(scatter)
            row.update({"p2p_scatter_ms": t_p2p, "p2p_scatter_busGBps": payload * (P - 1) / P / t_p2p / 1e6})
        rows.append(row)
    if dist.get_rank() == 0:
        print(json.dumps({"what": "Ulysses all-to-all sweep, bf16, 40 heads x 128", "peak_GBps_per_dir": 900, "rows": rows}, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
