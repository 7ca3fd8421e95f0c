"""Probe of the GPU box's host side: memory, cgroup limits, disk, pinned-allocation and PCIe rates.
Run under gpurun; writes gpurun_out/probe_host.json.  Used to size the host-resident float32 store
of the bf16-primary configuration (100M rows on 2 GPUs)."""
import json, os, shutil, time, sys
out = {}
def rd(p):
    try:
        return open(p).read().strip()
    except Exception as e:
        return "ERR " + repr(e)
mi = {}
for line in open("/proc/meminfo"):
    k, v = line.split(":", 1)
    if k in ("MemTotal", "MemFree", "MemAvailable", "Shmem", "Mlocked", "Unevictable"):
        mi[k] = v.strip()
out["meminfo"] = mi
out["cgroup_memory_max"] = rd("/sys/fs/cgroup/memory.max")
out["cgroup_memory_current"] = rd("/sys/fs/cgroup/memory.current")
out["cgroup_v1_limit"] = rd("/sys/fs/cgroup/memory/memory.limit_in_bytes")
out["nproc"] = os.cpu_count()
try:
    out["affinity"] = len(os.sched_getaffinity(0))
except Exception:
    pass
for p in ("/tmp", "/root", "/dev/shm"):
    try:
        du = shutil.disk_usage(p)
        out["disk_" + p] = {"total_gb": du.total / 1e9, "free_gb": du.free / 1e9}
    except Exception as e:
        out["disk_" + p] = repr(e)
import resource
out["rlimit_memlock"] = resource.getrlimit(resource.RLIMIT_MEMLOCK)
import torch
out["gpus"] = torch.cuda.device_count()
free, total = torch.cuda.mem_get_info(0)
out["gpu_mem"] = {"free_gb": free / 1e9, "total_gb": total / 1e9}
gb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
t0 = time.perf_counter()
h = torch.empty(gb * (1 << 30), dtype=torch.uint8, pin_memory=True)
out["pin_alloc_s_per_%dGiB" % gb] = time.perf_counter() - t0
d = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for name, (src, dst) in {"d2h": (d, h[: 4 << 30]), "h2d": (h[: 4 << 30], d)}.items():
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    out[name + "_GBps"] = 4 * (4 << 30) / 1e9 / (time.perf_counter() - t0)
json.dump(out, open("gpurun_out/probe_host.json", "w"), indent=1)
print(json.dumps(out, indent=1))
