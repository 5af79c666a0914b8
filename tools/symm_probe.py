"""Probe (torchrun, NCCL): symmetric-memory allocation of a table-sized buffer, peer pointers, remote row-gather bandwidth."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    N, w = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, 514
    t0 = time.perf_counter()
    t = symm_mem.empty((N, w), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"symm alloc + rendezvous of {N * w * 4 / 1e9:.2f} GB: {time.perf_counter() - t0:.2f} s; ptrs {[hex(p) for p in hdl.buffer_ptrs]} "
              f"multicast {hdl.has_multicast_support}", flush=True)
    t.fill_(float(rank + 1))
    dist.barrier()
    torch.cuda.synchronize()
    peer = (rank + 1) % world
    pt = hdl.get_buffer(peer, (N, w), torch.float32)
    ok = float(pt[12345, 7].item()) == float(peer + 1)
    g = torch.Generator(device=dev).manual_seed(rank)
    idx = torch.randint(0, N, (50_500,), generator=g, device=dev)
    for name, src in (("local", t), ("peer", pt)):
        out = torch.empty(idx.numel(), w, device=dev)
        for _ in range(3):
            torch.index_select(src, 0, idx, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.index_select(src, 0, idx, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if rank == 0:
            print(f"row gather 50500 x {w * 4} B from {name}: {ms * 1e3:.1f} us  {idx.numel() * w * 4 / ms / 1e6:.0f} GB/s  (peer value ok {ok})", flush=True)
    # mixed: rows spread over all ranks' buffers (what a sharded K3 does)
    bufs = [hdl.get_buffer(k, (N, w), torch.float32) for k in range(world)]
    out = torch.empty(idx.numel(), w, device=dev)
    parts = [idx[k::world] for k in range(world)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = 0
        for k in range(world):
            torch.index_select(bufs[k], 0, parts[k], out=out[o:o + parts[k].numel()])
            o += parts[k].numel()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"row gather spread over {world} ranks: {ms * 1e3:.1f} us  {idx.numel() * w * 4 / ms / 1e6:.0f} GB/s", flush=True)
    # in-place all_gather of the owned row blocks (replica sync)
    rows = (N + world - 1) // world
    if N % world == 0:
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        dist.all_gather_into_tensor(t.view(-1), t[rank * rows:(rank + 1) * rows].reshape(-1))
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"in-place all_gather of the table: {e0.elapsed_time(e1):.2f} ms; row 0 of last block = {t[(world - 1) * rows, 0].item()}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
