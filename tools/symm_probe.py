"""2+ rank probe of the peer-memory communicator (torchrun): symmetric allocation, barrier, all-reduce vs NCCL, small
all-reduce, and the same sequence replayed from a CUDA graph. usage: torchrun --nproc-per-node N tools/symm_probe.py"""
import datetime, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from iswm_b200.peer import PeerComm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
comm = PeerComm(dev)
n = 40_347_300
buf, ptrs = comm.alloc(n, torch.float32)
pa = comm._ptr_array(ptrs)
g = torch.Generator(device=dev).manual_seed(rank)
src = torch.randn(n, device=dev, generator=g)
buf.copy_(src)
ref = src.clone()
dist.all_reduce(ref)
st = torch.cuda.current_stream().cuda_stream
comm.allreduce_f32(pa, 0, n, st)
torch.cuda.synchronize()
err = float((buf - ref).abs().max() / ref.abs().max())
# identical bits on every rank
chk = buf.double().sum().reshape(1)
lst = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(lst, chk)
same = all(float(a) == float(lst[0]) for a in lst)
h = torch.tensor([rank + 1, 10 * (rank + 1)], dtype=torch.int64, device=dev)
comm.small_allreduce_(h, 0, st)
f = torch.tensor([0.5 * (rank + 1)], dtype=torch.float64, device=dev)
comm.small_allreduce_(f, 1, st)
torch.cuda.synchronize()
tot = world * (world + 1) // 2
ok_small = h.tolist() == [tot, 10 * tot] and abs(float(f) - 0.5 * tot) < 1e-12
# timing + graph replay
buf.copy_(src)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    comm.allreduce_f32(pa, 0, n, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
gr = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    comm.allreduce_f32(pa, 0, n, s.cuda_stream)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize(); dist.barrier()
with torch.cuda.graph(gr):
    cs = torch.cuda.current_stream().cuda_stream
    comm.allreduce_f32(pa, 0, n, cs)
    comm.small_allreduce_(h, 0, cs)
buf.copy_(src)
h.copy_(torch.tensor([rank + 1, 10 * (rank + 1)], device=dev))
torch.cuda.synchronize(); dist.barrier()
gr.replay()
torch.cuda.synchronize()
err_g = float((buf - ref).abs().max() / ref.abs().max())
ok_g = h.tolist() == [tot, 10 * tot]
e0.record()
for _ in range(5):
    gr.replay()
e1.record(); torch.cuda.synchronize()
ms_g = e0.elapsed_time(e1) / 5
# NCCL for comparison
e0.record()
for _ in range(5):
    dist.all_reduce(ref)
e1.record(); torch.cuda.synchronize()
ms_n = e0.elapsed_time(e1) / 5
if rank == 0:
    print(f"PROBE world={world} n={n} rel_err_vs_nccl={err:.3e} same_bits={same} small_ok={ok_small} graph_err={err_g:.3e} graph_small_ok={ok_g} "
          f"peer_ms={ms:.3f} ({n * 4 * 2 * (world - 1) / world / ms / 1e6:.1f} GB/s bus) graph_ms={ms_g:.3f} nccl_ms={ms_n:.3f}")
dist.barrier()
dist.destroy_process_group()
