"""Device -> host copy rate of the e2e leg's result buffer (545 MB per step per rank) into (a) torch's pinned allocation (cudaHostAlloc, 4 KB pages) and
(b) a 2 MB-aligned anonymous mapping with MADV_HUGEPAGE, touched and cudaHostRegister'ed.  Question: is the ~92 GB/s aggregate ceiling of the multi-GPU e2e
leg a property of the host's translation path (IOMMU / page size) that bigger pages lift?
    torchrun --nproc-per-node N tools/d2h_host_path_bench.py"""
import ctypes, json, mmap, os, sys
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NB = (1 << 20) * 2 * 65 * 4   # obs of 1 Mi mazes
HP = 2 << 20


def huge_registered(nbytes):
    sz = (nbytes + HP - 1) & ~(HP - 1)
    m = mmap.mmap(-1, sz + HP, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(m)); al = (addr + HP - 1) & ~(HP - 1)
    rc_adv = ctypes.CDLL(None, use_errno=True).madvise(ctypes.c_void_p(al), ctypes.c_size_t(sz), 14)   # MADV_HUGEPAGE
    ctypes.memset(al, 0, sz)
    rc = torch.cuda.cudart().cudaHostRegister(al, sz, 0)
    t = torch.frombuffer((ctypes.c_char * sz).from_address(al), dtype=torch.uint8)[:nbytes]
    return t, m, int(rc_adv), int(rc)


def thp_state():
    try:
        return open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
    except Exception as e:
        return str(e)


def anon_huge_kb():
    try:
        for ln in open("/proc/self/smaps_rollup"):
            if ln.startswith("AnonHugePages"):
                return int(ln.split()[1])
    except Exception:
        return -1


d = torch.empty(NB, dtype=torch.uint8, device="cuda").random_(0, 255)
res = {}
for name in ("cudaHostAlloc", "hugepage_registered"):
    if name == "cudaHostAlloc":
        h = torch.empty(NB, dtype=torch.uint8, pin_memory=True); extra = {}
    else:
        h, keep, rc_adv, rc_reg = huge_registered(NB); extra = {"madvise_rc": rc_adv, "register_rc": rc_reg, "AnonHugePages_kB": anon_huge_kb(), "is_pinned": bool(h.is_pinned())}
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res[name] = {"ms_per_copy_max_over_ranks": float(ms), "GBs_per_rank": NB / float(ms) / 1e6, "GBs_aggregate": world * NB / float(ms) / 1e6, **extra}
    ok = bool((h[:4096] == d[:4096].cpu()).all())
    res[name]["verified"] = ok
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes": NB, "thp": thp_state(), **res}))
if world > 1:
    dist.destroy_process_group()
