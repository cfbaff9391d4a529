"""Probe: does torch symmetric memory (peer-mapped buffers + device barrier) work on this box?  torchrun, 2+ ranks."""
import os, sys, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as sm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    t = sm.empty((1024, 64), dtype=torch.bfloat16, device=dev)
    hdl = sm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; ptrs", [hex(p) for p in hdl.buffer_ptrs], "attrs", [a for a in dir(hdl) if not a.startswith("_")][:30], flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    peer = (rank + 1) % world
    remote = hdl.get_buffer(peer, (1024, 64), torch.bfloat16)
    remote[rank * 8:(rank + 1) * 8].fill_(100.0 + rank)     # write into the peer's buffer
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    src = (rank - 1) % world
    ok = bool((t[src * 8:(src + 1) * 8] == 100.0 + src).all()) and float(t[500, 0]) == rank + 1
    # barrier latency
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        hdl.barrier(channel=0)
    e1.record(); torch.cuda.synchronize()
    # graph capture of barriers
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            hdl.barrier(channel=0)
            t.add_(1.0)
            hdl.barrier(channel=0)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    print(rank, "peer write visible:", ok, "barrier us:", e0.elapsed_time(e1) * 1000 / 200, "graph ok", float(t[500, 0]), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "SYMM FAILED", repr(e), flush=True)
sys.stdout.flush()
os._exit(0)
