"""Does the kind of pinned allocation change the box's aggregate device-to-host bandwidth?  
  torchrun --nproc-per-node N profiles/pcie_wc.py
Times cudaMemcpyAsync D2H of 174 MB (one step's result rows of 1M Heavy-v0 envs) from every GPU at once into
  (a) torch's pinned allocator (cudaHostAlloc default),  (b) cudaHostAllocWriteCombined,  (c) cudaHostAllocPortable | Mapped,
  (d) malloc + cudaHostRegister, and prints the host topology the ranks see."""
import ctypes as C
import os
import subprocess

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl")
rt = C.CDLL([p for p in (os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart.so.12"), "libcudart.so.12", "libcudart.so") if os.path.exists(p) or "/" not in p][0])
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
n = 174 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda")
libc = C.CDLL("libc.so.6")
libc.malloc.restype = C.c_void_p
libc.malloc.argtypes = [C.c_size_t]


def alloc(kind):
    p = C.c_void_p()
    if kind == "torch":
        t = torch.empty(n, dtype=torch.uint8).pin_memory()
        return t.data_ptr(), t
    if kind == "registered":
        q = libc.malloc(n + 4096)
        q = (q + 4095) & ~4095
        C.memset(q, 0, n)
        assert rt.cudaHostRegister(C.c_void_p(q), n, 0) == 0
        return q, None
    flags = {"default": 0, "write_combined": 4, "portable_mapped": 1 | 2}[kind]
    assert rt.cudaHostAlloc(C.byref(p), n, flags) == 0
    return p.value, None


def timed(ptr, reps=10):
    st = torch.cuda.current_stream().cuda_stream
    def fn():
        rt.cudaMemcpyAsync(C.c_void_p(ptr), C.c_void_p(d.data_ptr()), n, 2, C.c_void_p(st))
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            keep = [l for l in out.splitlines() if cmd[0] == "nvidia-smi" or any(k in l for k in ("Model name", "Socket", "NUMA", "CPU(s):", "Thread"))]
            print("\n".join(keep[:24]), flush=True)
        except Exception as e:  # noqa: BLE001
            print(cmd, "failed:", e)
    print("affinity of rank 0:", sorted(os.sched_getaffinity(0)), flush=True)
for kind in ("torch", "default", "write_combined", "portable_mapped", "registered"):
    ptr, keep = alloc(kind)
    ms = timed(ptr)
    if rank == 0:
        print(f"{world} GPUs at once, {kind:16s}: D2H {n / 1e6:.0f} MB per GPU in {ms:.2f} ms = {n / ms / 1e6:.1f} GB/s per GPU, {world * n / ms / 1e6:.1f} GB/s aggregate", flush=True)
if world > 1:
    dist.destroy_process_group()
