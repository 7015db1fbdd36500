"""Device plumbing of the agents (streams, events, CUDA graphs, pinned staging, the native library), in one place.

`CudaRuntime` is the product's: CUDA streams / events / graphs of torch, the nvcc-built C-ABI library.  There is no CPU
implementation in the package; the test-suite substitutes a host runtime (tests/cusim/emu_agent.py: kernel emulator build,
no-op streams) so that the agents' host-side logic — sharding, gradient accumulation, graph-slot bookkeeping, pickling — is
exercised on machines without a GPU."""
import contextlib

import torch

from molgym_b200 import _lib


class CudaGraphStep:
    """A captured launch sequence.  `fn(stream_ptr)` enqueues the C-ABI calls; replay() re-issues them."""

    def __init__(self, runtime, fn):
        dev = runtime.device
        fn(torch.cuda.current_stream(dev).cuda_stream)   # eager warm-up: function attributes, lazy module loading
        torch.cuda.current_stream(dev).synchronize()
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev, priority=-5)   # the critical chain outranks the plan's (default-priority) side streams
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(self.graph, stream=side):
                fn(torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.current_stream(dev).wait_stream(side)

    def replay(self):
        self.graph.replay()


class CudaRuntime:
    is_cuda = True

    def __init__(self, device):
        self.device = _lib.require_cuda_device(device)

    def lib(self):
        return _lib.load()

    def device_ctx(self):
        return torch.cuda.device(self.device)

    def current_stream(self):
        return torch.cuda.current_stream(self.device)

    def stream_ptr(self, stream=None):
        return (stream if stream is not None else torch.cuda.current_stream(self.device)).cuda_stream

    def new_stream(self, priority=0):
        return torch.cuda.Stream(self.device, priority=priority)

    def new_event(self):
        return torch.cuda.Event()

    def stream_ctx(self, stream):
        return torch.cuda.stream(stream)

    def pinned(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)

    def capture(self, fn):
        return CudaGraphStep(self, fn)

    def total_memory(self):
        return torch.cuda.get_device_properties(self.device).total_memory


null_ctx = contextlib.nullcontext
