"""Drop-in for `transducer.functions.transducer` of awni/transducer, which the reference imports at
speech/models/transducer_model.py:11 (call site :46-52; dependency cloned by Makefile:10-12).

    loss = TransducerLoss()(log_probs, labels, x_lens, y_lens)      # 1-element tensor

    log_probs  (B, T, U+1, V+1) log-softmax on the CUDA device (the reference applies log_softmax
               itself, transducer_model.py:76); gradient flows to it
    labels     IntTensor flat (sum(y_lens),);  x_lens, y_lens IntTensor (B,)   (CPU or CUDA)
Blank = last class (transducer_model.py:28); reduction = sum over the minibatch.  Both are
constructor keywords because the un-vendored dependency could not be inspected.
"""
import ctypes

import torch

from .. import _lib


def rnnt_costs_and_grads(log_probs, labels, x_lens, y_lens, blank=None, need_grad=True):
    _lib.require_cuda(log_probs, "log_probs")
    lib = _lib.load()
    lp = log_probs.float().contiguous()
    B, T, U1, V = lp.shape
    if blank is None:
        blank = V - 1
    dev = lp.device
    lab = labels.detach().to("cpu", torch.int32).reshape(-1)
    ylen = y_lens.detach().to("cpu", torch.int32).reshape(-1)
    xlen = x_lens.detach().to("cpu", torch.int32).reshape(-1)
    if int(ylen.sum()) != lab.numel():
        raise ValueError("labels / y_lens mismatch")
    if B and int(ylen.max()) > U1 - 1:
        raise ValueError("a label sequence is longer than the lattice (U+1 = %d)" % U1)
    offs = torch.zeros(B, dtype=torch.int32)
    if B > 1:
        offs[1:] = torch.cumsum(ylen[:-1], 0)
    n = lab.numel()
    packed = torch.cat([lab, offs, ylen, xlen]).pin_memory().to(dev, non_blocking=True)
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.sb_rnnt_workspace_size(B, T, U1, ctypes.byref(nbytes)), "sb_rnnt_workspace_size")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    costs = torch.empty(B, dtype=torch.float32, device=dev)
    grads = torch.empty_like(lp) if need_grad else None
    from .. import ops
    with torch.cuda.device(dev):
        sp = _lib.stream_ptr()
        ops._launch("rnnt_fwd_bwd", 0.0,
                    lambda: lib.sb_rnnt_fwd_bwd(lp.data_ptr(), _lib.ptr(grads),
                                                packed[:n].data_ptr(),
                                                packed[n:n + B].data_ptr(),
                                                packed[n + B:n + 2 * B].data_ptr(),
                                                packed[n + 2 * B:].data_ptr(), B, T, U1, V,
                                                int(blank), costs.data_ptr(), ws.data_ptr(),
                                                nbytes.value, sp))
    return costs, grads


class _RNNTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, labels, x_lens, y_lens, blank, size_average):
        costs, grads = rnnt_costs_and_grads(log_probs, labels, x_lens, y_lens, blank,
                                            log_probs.requires_grad)
        loss = costs.sum().reshape(1)
        if size_average:
            loss = loss / log_probs.shape[0]
            if grads is not None:
                grads = grads / log_probs.shape[0]
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.grads
        if g is None:
            return (None,) * 6
        return (g * grad_out.reshape(1, 1, 1, 1),) + (None,) * 5


class TransducerLoss(torch.nn.Module):
    def __init__(self, blank=None, size_average=False):
        super().__init__()
        self.blank = blank
        self.size_average = size_average

    def forward(self, log_probs, labels, x_lens, y_lens):
        return _RNNTFunction.apply(log_probs, labels, x_lens, y_lens, self.blank,
                                   self.size_average)
