"""TEST TOOL.  Runs bench.py's product arm (N = 1) on a machine without a GPU: the C ABI over the simulated CUDA runtime
(build it first: python -c 'from racon_b200 import build; build.build_simapi()'), torch.cuda replaced
by stand-ins (wall-clock events, no-op streams).  It exists to catch Python-level errors in bench.py's main path before a
GPU call is spent on them; the numbers it prints are meaningless.  Takes minutes even for a handful of windows:
    python tools/mock_bench.py --windows 6 --steps 1 --warmup 1 [--no-cpu-baseline]
    python tools/mock_bench.py --mock-world 2 --windows 4 --steps 1 --warmup 1     # two ranks over gloo"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RACON_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "racon_b200", "lib", "simapi", "libracon_b200.so")
os.environ["RP_BENCH_NO_BY_REFERENCE"] = "1"
import torch


class Ev:
    def __init__(self, enable_timing=False):
        self.t = None
    def record(self, stream=None):
        self.t = time.perf_counter()
    def elapsed_time(self, other):
        return 1e3 * (other.t - self.t)
    def synchronize(self):
        pass
    def query(self):
        return True


class St:
    cuda_stream = 0
    def wait_event(self, ev):
        pass
    def synchronize(self):
        pass


class Props:
    name = "fake"
    multi_processor_count = 2
    pci_bus_id = 0
    pci_device_id = 0
    pci_domain_id = 0
    total_memory = 8 << 30


c = torch.cuda
c.is_available = lambda: True
c.set_device = lambda d: None
c.get_device_properties = lambda d: Props()
c.current_stream = lambda *a: St()
c.Stream = lambda *a, **k: St()
c.Event = Ev
c.synchronize = lambda *a: None
c.empty_cache = lambda: None
c.mem_get_info = lambda *a: (4 << 30, 8 << 30)
_tensor = torch.tensor
torch.tensor = lambda data, device=None, **kw: _tensor(data, **kw)

# --mock-world N: N ranks of this script under gloo (launched below), device "cuda" meaning the CPU everywhere — exercises
# bench.py's multi-rank control flow (binding, process group, per-step gather, reductions, the wait for rank 0's baselines)
if "--mock-world" in sys.argv:
    k = sys.argv.index("--mock-world")
    n_ranks = int(sys.argv[k + 1])
    del sys.argv[k:k + 2]
    if "RANK" not in os.environ:
        import socket
        import subprocess
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        procs = []
        for r in range(n_ranks):
            env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(n_ranks), LOCAL_WORLD_SIZE=str(n_ranks),
                       MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RP_SIM_DEVICES=str(n_ranks))
            procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--mock-world", str(n_ranks)] +
                                          sys.argv[1:] + ["--gpus", str(n_ranks)], env=env))
        sys.exit(max(p.wait() for p in procs))
    import torch.distributed as dist
    _init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: _init("gloo")
    _empty = torch.empty
    torch.empty = lambda *a, pin_memory=False, device=None, **kw: _empty(
        *a, device=None if device is not None and str(device).startswith("cuda") else device, **kw)
    _from_numpy_to = torch.Tensor.to
    torch.Tensor.to = lambda self, *a, **kw: _from_numpy_to(
        self, *[x for x in a if not str(x).startswith("cuda")], **{k2: v for k2, v in kw.items() if k2 != "non_blocking"})

import bench
sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()
