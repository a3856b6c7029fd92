"""Concurrent host<->device copy ceiling of this box: every rank copies the bench's per-step input
(H2D) and output (D2H) volumes between page-locked host memory and its GPU at the same time, on two
streams, as speedyBatchProcess does.  Run under torchrun with N = 1, 2, 4, 8:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      profiles/tools/pcie_ceiling.py [out.jsonl]
Prints one JSON line (rank 0): per-rank and aggregate GB/s for H2D alone, D2H alone and both together."""
import json, os, sys, time
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
IN_B, OUT_B = 1966080000, 1085648980  # config 2: 1024 x 960000 x 2 bytes in, what it produces out
h_in = torch.empty(IN_B // 2, dtype=torch.int16, pin_memory=True); h_in.zero_()
h_out = torch.empty(OUT_B // 2, dtype=torch.int16, pin_memory=True); h_out.zero_()
d_in = torch.empty(IN_B // 2, dtype=torch.int16, device='cuda')
d_out = torch.zeros(OUT_B // 2, dtype=torch.int16, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

def run(do_in, do_out, reps=4):
    best = 1e9
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        if do_in:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device='cuda')
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best

t_in, t_out, t_both = run(True, False), run(False, True), run(True, True)
if rank == 0:
    line = {"n_gpus": world, "h2d_bytes": IN_B, "d2h_bytes": OUT_B,
            "h2d_alone_ms": t_in * 1e3, "d2h_alone_ms": t_out * 1e3, "both_ms": t_both * 1e3,
            "h2d_alone_gbs_per_gpu": IN_B / t_in / 1e9, "d2h_alone_gbs_per_gpu": OUT_B / t_out / 1e9,
            "both_gbs_per_gpu": (IN_B + OUT_B) / t_both / 1e9, "both_gbs_aggregate": world * (IN_B + OUT_B) / t_both / 1e9,
            "e2e_step_floor_ms": t_both * 1e3,
            "cpu_affinity": sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]}
    print(json.dumps(line))
    if len(sys.argv) > 1:
        open(sys.argv[1], "a").write(json.dumps(line) + "\n")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
