"""TEST INFRASTRUCTURE: lets the torch-side orchestration (bench.py, dealii_cuda_b200/distributed.py, the torchrun workers of the GPU
tests) run on the CPU against libmfgpu_emu.so.  The emulated library's device memory IS host memory, so a CPU tensor's data_ptr() is
a valid "device" pointer: tensors asked for on "cuda" are created on the CPU, torch.cuda's streams / events / synchronisation become
no-ops with wall-clock timing (every emulated launch completes before it returns), the process group is gloo instead of NCCL (gloo has
all_reduce, all_to_all_single, barrier and all_gather_object on CPU tensors).  CUDA graphs and symmetric memory do not exist here:
capture raises, which sends the callers down their documented eager / NCCL-exchange fallbacks."""
import contextlib
import time


def install():
    import torch
    import torch.distributed as dist

    def drop_device(fn):
        def f(*a, **k):
            k.pop("device", None)
            if isinstance(k.get("generator"), _Generator):
                k["generator"] = k["generator"].g
            return fn(*a, **k)
        return f

    for name in ("full", "zeros", "empty", "tensor", "rand", "ones", "arange", "randn"):
        setattr(torch, name, drop_device(getattr(torch, name)))

    class _Generator:
        def __init__(self, *a, **k):
            self.g = _RealGenerator()

        def manual_seed(self, s):
            self.g.manual_seed(s)
            return self

    _RealGenerator = torch.Generator
    torch.Generator = _Generator
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.is_cuda = property(lambda self: True)   # (GpuVector.wrap asserts it; the data is where the emulated library expects it)
    torch.Tensor.pin_memory = lambda self, *a, **k: self

    class Stream:
        def __init__(self, *a, **k):
            self.cuda_stream = 0

        def wait_event(self, e):
            pass

        def wait_stream(self, s):
            pass

        def synchronize(self):
            pass

    class Event:
        def __init__(self, enable_timing=False, **k):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def wait(self, stream=None):
            pass

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return max(1e-6, (other.t - self.t) * 1e3)

    class CUDAGraph:
        def replay(self):
            raise RuntimeError("no CUDA graphs in the CPU shim")

    @contextlib.contextmanager
    def graph(g, stream=None, **k):
        raise RuntimeError("no CUDA graph capture in the CPU shim")
        yield

    @contextlib.contextmanager
    def stream_ctx(s):
        yield

    state = {"current": Stream(), "default": Stream()}
    state["current"] = state["default"]
    cuda = torch.cuda
    cuda.is_available = lambda: True
    cuda.device_count = lambda: int(__import__("os").environ.get("WORLD_SIZE", "1"))
    cuda.set_device = lambda d: None
    cuda.current_device = lambda: 0
    cuda.synchronize = lambda *a, **k: None
    cuda.init = lambda: None
    cuda.empty_cache = lambda: None
    cuda.Stream = Stream
    cuda.Event = Event
    cuda.CUDAGraph = CUDAGraph
    cuda.graph = graph
    cuda.stream = stream_ctx
    cuda.set_stream = lambda s: state.__setitem__("current", s)
    cuda.current_stream = lambda *a, **k: state["current"]
    cuda.default_stream = lambda *a, **k: state["default"]

    def no_props(*a, **k):
        raise RuntimeError("no device properties in the CPU shim")
    cuda.get_device_properties = no_props

    real_init = dist.init_process_group

    def init_process_group(backend=None, *a, **k):
        k.pop("device_id", None)
        return real_init("gloo", *a, **k)
    dist.init_process_group = init_process_group
