"""CUDA-graph replay of a whole model forward (opt-in).

Every forward here is a fixed sequence of kernel launches over cached workspaces, so after one eager
call per input geometry the sequence is captured into a CUDA graph and replayed: one graph launch
instead of 30-300 ctypes launches (LAFC and RAFT are host-launch bound otherwise; FGT is GPU bound and
gains little). TMA descriptors are encoded on the host at capture time and baked into the graph, which
is valid because all buffers are static. Inputs are copied into static tensors, outputs are cloned.
"""
import torch

from . import lib


class GraphedCall:
    def __init__(self, fn):
        self.fn = fn
        self.entries = {}

    def __call__(self, *tensors):
        key = tuple((tuple(t.shape), t.dtype, t.device.index) for t in tensors)
        e = self.entries.get(key)
        if e is None:
            self.entries[key] = {}
            return self.fn(*tensors)  # eager: packs weights, allocates workspaces, sets kernel attributes
        if "graph" not in e:
            e["inp"] = [t.clone() for t in tensors]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = lib.COUNTERS["launches"]
            with torch.cuda.graph(g):
                e["out"] = self.fn(*e["inp"])
            e["launches"] = lib.COUNTERS["launches"] - l0
            e["graph"] = g
        for s, t in zip(e["inp"], tensors):
            s.copy_(t, non_blocking=True)
        e["graph"].replay()
        lib.COUNTERS["launches"] += e["launches"]
        out = e["out"]
        if isinstance(out, torch.Tensor):
            return out.clone()
        return type(out)(o.clone() if isinstance(o, torch.Tensor) else o for o in out)
