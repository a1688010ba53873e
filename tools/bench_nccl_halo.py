"""Development micro-benchmark: what the halo exchange `north_star` names as the baseline for a decomposed mesh would cost —
after every colour pass an ncclSend / ncclRecv of the interface vertices that pass modified, to each neighbour (SURVEY 8e:
<= 122 KB per neighbour and sweep, typically 1/8 of it per colour pass) — against the library's exchange, which is a store into
peer memory issued by the substep kernel itself (no host-side step at all).
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/bench_nccl_halo.py"""
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
peer = rank ^ 1
passes = 920                      # colour + collision steps of one frame (10 substeps x 10 iterations x 9 + predict/commit)
for floats in (4 * 1275, 4 * 10201):      # 1/8 of a 101 x 101 cut plane per pass (x, y, z, tag), and the whole plane
    send = torch.ones(floats, device="cuda")
    recv = torch.empty(floats, device="cuda")
    for rep in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(passes):
            ops = [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            send.add_(recv, alpha=1e-9)      # the next pass depends on what arrived (as the next colour does)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0 and rep == 2:
            print("%d dependent exchanges of %6d bytes with one neighbour over NCCL (ncclSend + ncclRecv in a group): %.2f ms "
                  "per frame, %.1f us per exchange" % (passes, 4 * floats, ms.item(), 1e3 * ms.item() / passes), flush=True)
dist.destroy_process_group()
