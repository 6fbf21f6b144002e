"""TEST TOOL.  Runs bench.py's product arm (N = 1) on a machine without a GPU: the C ABI over the simulated CUDA runtime
(build it first: python -c 'from racon_b200 import build; build.build_simapi()'), torch.cuda replaced
by stand-ins (wall-clock events, no-op streams).  It exists to catch Python-level errors in bench.py's main path before a
GPU call is spent on them; the numbers it prints are meaningless.  Takes minutes even for a handful of windows:
    python tools/mock_bench.py --windows 6 --steps 1 --warmup 1 [--no-cpu-baseline]"""
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

import bench
sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()
