"""Episode-level data parallelism: one process per GPU, episodes sharded by index, and a single
all-reduce (sum, int64) of the `FewShotMetric` counts per evaluation round.  No data-path collective."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():       # every backend: each rank launches on its own GPU (gloo never sets a device)
            torch.cuda.set_device(local_rank % torch.cuda.device_count())
        if backend == "nccl":
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_indices(n_episodes, rank, world, first=0):
    """Episodes i with i = rank (mod world): the union over ranks is exactly range(first, first + n) for any
    world size, so sharded and single-GPU runs see the same episode set."""
    return list(range(first + rank, first + n_episodes, world))


def all_reduce_stat(stat):
    """stat [(C+1), 3] int64 tensor summed over ranks in place (exact: integer addition)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stat, op=dist.ReduceOp.SUM)
    return stat


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def bind_host_to_gpu(local_rank):
    """Pin this process (and therefore the pinned host buffers it first-touches and its copy-issuing threads) to the CPU cores
    NVML reports as nearest to its GPU, so that with 4-8 ranks on one box the host->device copies do not all cross one NUMA
    node / PCIe root complex.  Best effort: returns a short description of what was done for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, ((os.cpu_count() or 64) + 63) // 64)
        near = {64 * i + b for i, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = near & allowed
        if not target:
            return f"gpu-near cores {_span(near)} are outside this process's cpuset {_span(allowed)}; not bound"
        os.sched_setaffinity(0, target)
        return f"bound to {len(target)} gpu-near cores {_span(target)} of {len(allowed)} allowed"
    except Exception as exc:           # no NVML, no permission: run unbound
        return f"not bound ({type(exc).__name__})"


def _span(cpus):
    cpus = sorted(cpus)
    if not cpus:
        return "[]"
    runs, a, prev = [], cpus[0], cpus[0]
    for c in cpus[1:]:
        if c != prev + 1:
            runs.append((a, prev))
            a = c
        prev = c
    runs.append((a, prev))
    return ",".join(f"{x}-{y}" if x != y else str(x) for x, y in runs)
