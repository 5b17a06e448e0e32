"""CUDA-graph replay of a whole optimisation step (zero_grad -> loss -> backward -> Adam) of the PFCN / FairGo families.

A step of these models is ~200-400 small kernels; launched one by one from Python the host is the bottleneck.  The step
is captured ONCE per (phase, attribute subset, batch size) into a CUDA graph over static input buffers and replayed with
one launch per batch.  What makes the captured step correct on every replay: Adam's per-parameter step counts and the
dropout seed offset live in device memory and are advanced by kernels inside the graph (ops.AdamGroup, ops.bump_seed)."""
import torch

from . import ops
from .interaction import Interaction


class GraphedStep:
    def __init__(self, loss_fn, optimizer, sst_list, example, device):
        """loss_fn(interaction, sst_list) -> 0-dim loss; example: Interaction (host or device) fixing shapes / dtypes"""
        self.device = device
        self.static = {k: torch.empty(example[k].shape, dtype=example[k].dtype, device=device) for k in example.columns}
        for k in example.columns:
            self.static[k].copy_(example[k])
        self.inter = Interaction(self.static)
        optimizer.init_state()
        ops.init_autograd_thread(device)
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            optimizer.zero_grad()
            with torch.cuda.graph(self.graph, stream=side):
                ops.bump_seed(device)
                optimizer.zero_grad()
                loss = loss_fn(self.inter, sst_list)
                loss.backward()
                optimizer.step()
                self.loss = loss.detach()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

    def matches(self, interaction):
        return all(k in interaction and interaction[k].shape == v.shape for k, v in self.static.items())

    def run(self, interaction):
        """copy the batch into the static buffers (H2D when it lives in host memory) and replay; returns the device
        tensor holding the loss of this replay (overwritten by the next one)"""
        for k, v in self.static.items():
            src = interaction[k]
            if src.data_ptr() != v.data_ptr():
                v.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss


class GraphedSteps:
    """cache of GraphedStep objects keyed by (phase name, attribute subset, batch rows)"""

    def __init__(self, device):
        self.device, self.cache = device, {}

    def run(self, name, loss_fn, optimizer, sst_list, interaction):
        key = (name, tuple(sst_list) if sst_list is not None else None, len(interaction))
        g = self.cache.get(key)
        if g is None:
            g = self.cache[key] = GraphedStep(loss_fn, optimizer, sst_list, interaction, self.device)
        return g.run(interaction)       # (capturing does not execute anything: the first batch is replayed like the rest)
