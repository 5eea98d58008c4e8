"""Host->device copy bandwidth with all ranks copying at once: default pinned memory vs write-combined pinned memory
(cudaHostAllocWriteCombined). Development aid for the 8-GPU e2e floor. torchrun --nproc-per-node N bench_micro/h2d_wc_probe.py"""
import ctypes, json, os, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rt = ctypes.CDLL("libcudart.so.12")
NB = 32 * 4096 * 12  # one step's inputs: 1.57 MB
NBUF = 8
dst = torch.empty(NB, dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(device=dev)
def barrier():
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
res = {"rank": rank}
for name, flags in (("default", 0), ("write_combined", 4), ("default_again", 0)):
    ptrs = []
    for _ in range(NBUF):
        p = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NB), ctypes.c_uint(flags)) == 0
        ctypes.memset(p, 1, NB)
        ptrs.append(p)
    def run(n):
        for i in range(n):
            rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), ptrs[i % NBUF], ctypes.c_size_t(NB), ctypes.c_int(1), ctypes.c_void_p(stream.cuda_stream))
        stream.synchronize()
    run(50)
    barrier()
    t0 = time.perf_counter()
    run(400)
    barrier()
    dt = time.perf_counter() - t0
    res[name + "_gb_per_s"] = round(NB * 400 / dt / 1e9, 2)
    for p in ptrs:
        rt.cudaFreeHost(p)
print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
