"""Host <-> device copy ceiling of the box for bench.py's end-to-end numbers.

  python tools/pcie_ceiling.py                                   (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_ceiling.py

Every rank copies bench.py's buffer sizes between pinned host memory (allocated after binding to the GPU's NUMA
node, like bench.py) and its GPU with plain cudaMemcpyAsync, all ranks at once: H2D alone, D2H alone, and the two
mixes of the compress (9.2 GB up, 1.9 GB down) and decompress (1.9 GB up, 9.2 GB down) end-to-end steps on two
streams.  Prints ONE JSON line; the "*_fastq_gbs" figures are in bench.py's unit (FASTQ bytes of all ranks per
second) and are the ceiling its e2e values can reach on this box at this N."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from bench import _bind_to_gpu_numa_node

F = 9_200_652_834  # FASTQ bytes per rank (config 2)
Z = 1_877_367_034  # .fqz bytes per rank


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h_big = torch.empty(F, dtype=torch.uint8, pin_memory=True)
    h_small = torch.empty(Z, dtype=torch.uint8, pin_memory=True)
    h_big.fill_(1)
    h_small.fill_(2)
    d_big = torch.empty(F, dtype=torch.uint8, device="cuda")
    d_small = torch.empty(Z, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(up, dn, reps=3):
        """up / dn: (host, device) pairs copied on their own streams at the same time; returns max-over-ranks seconds per rep"""
        best = None
        for _ in range(reps + 1):  # first pass warms up
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s_up.wait_event(e0)
            s_dn.wait_event(e0)
            if up:
                with torch.cuda.stream(s_up):
                    up[1].copy_(up[0], non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn):
                    dn[0].copy_(dn[1], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_up)
            torch.cuda.current_stream().wait_stream(s_dn)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            v = float(t.item())
            best = v if best is None or _ == 1 else min(best, v)
        return best

    t_h2d = run((h_big, d_big), None)
    t_d2h = run(None, (h_big, d_big))
    t_comp = run((h_big, d_big), (h_small, d_small))
    t_dec = run((h_small, d_small), (h_big, d_big))
    if rank == 0:
        print(json.dumps({
            "n_gpus": world, "fastq_bytes_per_rank": F, "fqz_bytes_per_rank": Z, "cpus_bound_to_gpu_numa_node": numa,
            "h2d_gbs": world * F / t_h2d / 1e9, "d2h_gbs": world * F / t_d2h / 1e9,
            "compress_mix_fastq_gbs": world * F / t_comp / 1e9, "decompress_mix_fastq_gbs": world * F / t_dec / 1e9,
            "note": "aggregate over all ranks, best of 3, plain pinned cudaMemcpyAsync, both directions on separate streams in the mixes",
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
