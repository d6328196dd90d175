"""Host-side process settings that matter for the end-to-end path.

PyTorch sizes its intra-op thread pool from the number of CPUs it can SEE; inside a container with a CFS quota (cgroup
`cpu.max`) that is far more than the process may USE.  The staging copies of `ClipRunner.submit` then fan out over e.g.
64 spinning OpenMP threads on a 16-CPU quota, the cgroup is throttled for the rest of each 100 ms period, and every ~6th
clip stalls for ~21 ms while the GPU idles (measured on the B200 boxes: cpu.max = 1600000/100000 with 128 visible CPUs;
profiles/r02_e2e_breakdown.txt).  `tune_host_threads()` caps the pool at the quota (shared between the ranks of one host)."""
import os


def cpu_quota():
    """CPUs this process may use: min(visible CPUs, cgroup v2 / v1 CFS quota, scheduler affinity)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path in ("/sys/fs/cgroup/cpu.max",):
        try:
            q, p = open(path).read().split()[:2]
            if q != "max":
                n = min(n, max(1, int(q) // int(p)))
        except (OSError, ValueError):
            pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0 and p > 0:
            n = min(n, max(1, q // p))
    except (OSError, ValueError):
        pass
    return n


def tune_host_threads(ranks_on_host=1, cap=4):
    """Cap torch's intra-op pool at min(cap, quota / ranks_on_host) — the GPU path needs a few host threads for staging
    copies, never dozens.  Returns the thread count in effect."""
    import torch
    want = max(1, min(cap, cpu_quota() // max(1, ranks_on_host)))
    if torch.get_num_threads() > want:
        torch.set_num_threads(want)
    return torch.get_num_threads()
