"""Optimizer side of the training step: the `clip_grad_norm_` + `Adamax` +
TUNE_PARTIAL reset of `SDNetTrainer.update` (Models/SDNetTrainer.py:363-371) on ONE flat fp32 buffer,
with the data-parallel gradient mean (NCCL all-reduce of that buffer) in front of it.

The gradients come from the differentiable path (ruart_b200/autograd_ops.py: hand-written backward kernels
behind torch.autograd.Function); the reference's update rule is then two kernel launches and no host
synchronisation.
"""
import torch
import torch.distributed as dist

from ._lib import current_stream, ptr
from .ops import call


class FlatAdamax(object):
    """torch.optim.Adamax(params, lr) + clip_grad_norm_(params, max_norm), fused over a flat buffer.

    The parameters are re-pointed at views of one contiguous fp32 CUDA buffer (values preserved), so
    the update is one launch.  A parameter whose gradient is None in a step is treated as having a
    zero gradient (its moments still decay), whereas torch skips it: pass only parameters that get a
    gradient in every step — in the shipped conf the reference's 89 tensors, i.e. everything trainable
    except the GRUCell of GetFinalScores, which the forward never uses (Layers.py:395-397)."""

    def __init__(self, params, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev for p in self.params):
            raise RuntimeError("FlatAdamax needs fp32 CUDA parameters on one device (no CPU fallback)")
        self.lr, self.betas, self.eps, self.max_norm = float(lr), betas, float(eps), max_norm
        self.n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_inf = torch.zeros_like(self.flat)
        self._ws = torch.empty(1025, dtype=torch.float64, device=dev)   # partial sums + the result
        self.step_count = 0
        o = 0
        self._slices = []
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[o:o + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + k].view(p.shape)
                self._slices.append((o, k))
                o += k

    def load_grads(self, grads=None):
        """Copy `grads` (list aligned with params; default p.grad) into the flat gradient buffer."""
        with torch.no_grad():
            for i, (p, (o, k)) in enumerate(zip(self.params, self._slices)):
                g = p.grad if grads is None else grads[i]
                if g is None:
                    self.grad[o:o + k].zero_()
                else:
                    self.grad[o:o + k].copy_(g.reshape(-1))
        return self.grad

    def allreduce_mean(self):
        """Data-parallel mean of the flat gradient (one NCCL all-reduce; no-op for world size 1)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
            self.grad.div_(dist.get_world_size())

    def step(self, grads=None, reset=(), loaded=False):
        """One update.  `reset`: iterable of (param, first_row, rows_tensor) restored afterwards —
        the TUNE_PARTIAL rule (`weight.data[tune_partial:] = fixed_embedding`, SDNetTrainer.py:367-371).
        loaded=True: the caller already ran load_grads() + allreduce_mean() (e.g. to time the collective)."""
        if not loaded:
            self.load_grads(grads)
            self.allreduce_mean()
        st = current_stream()
        self.step_count += 1
        sq = None
        if self.max_norm is not None:
            call("ruart_grad_sqnorm", ptr(self.grad), self.n, ptr(self._ws), self._ws.data_ptr() + 8 * 1024, st)
            sq = self._ws.data_ptr() + 8 * 1024
        call("ruart_adamax_step", ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_inf), self.n,
             self.lr, self.betas[0], self.betas[1], self.eps, self.step_count, sq,
             float(self.max_norm) if self.max_norm is not None else 0.0, st)
        with torch.no_grad():
            for p, first, rows in reset:
                p.data[first:].copy_(rows)
            # the kernels wrote through raw pointers: tell autograd / the prepared-weight caches of the
            # inference path (keyed on data_ptr + _version, sdnet_ops.prep_weight, BertEngine._weights_key)
            # that every parameter changed
            for p in self.params:
                torch.autograd.graph.increment_version(p)

    def grad_norm(self):
        """Total gradient norm of the last step (a device read-back; for logging only)."""
        return float(self._ws[1024].sqrt())
