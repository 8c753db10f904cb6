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


# ------------------------------------------------------------------------------------------------
# Fused joint network + loss: the (B, T', U+1, H) and (B, T', U+1, V+1) tensors of the reference
# (transducer_model.py:71-76) never exist on the training path.
# ------------------------------------------------------------------------------------------------
JOINT_SLAB_FRAMES = 8     # frames whose hidden activations are materialised at a time in backward


def _pack_labels(labels, x_lens, y_lens, B, U1, dev):
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
    return packed[:n], packed[n:n + B], packed[n + B:n + 2 * B], packed[n + 2 * B:]


def joint_log_probs(fx, fy, fc2, ymat, blank):
    """Full (B, T, U1, V1) log-probabilities of the joint network (what `Transducer.infer` hands
    to the beam search), from fx = fc1(x) (B,T,H), fy = fc1(pred) (B,U1,H); no autograd."""
    from .. import ops
    _lib.require_cuda(fx, "fx")
    lib = _lib.load()
    B, T, H = fx.shape
    U1 = fy.shape[1]
    V1 = fc2.weight.shape[0]
    fxc, fyc = fx.detach().float().contiguous(), fy.detach().float().contiguous()
    w2 = fc2.weight.detach().to(torch.bfloat16).contiguous()
    b2 = fc2.bias.detach().float().contiguous()
    ym = ymat.detach().to(fx.device, torch.int32).contiguous()
    lat = torch.empty(T * B * U1, 2, dtype=torch.float32, device=fx.device)
    out = torch.empty(B, T, U1, V1, dtype=torch.float32, device=fx.device)
    sp = _lib.stream_ptr()
    ops._launch("rnnt_joint", 2.0 * B * T * U1 * H * V1,
                lambda: lib.sb_rnnt_joint_fwd(fxc.data_ptr(), fyc.data_ptr(), w2.data_ptr(),
                                              b2.data_ptr(), ym.data_ptr(), lat.data_ptr(),
                                              out.data_ptr(), B, T, U1, H, V1, int(blank), sp))
    return out


class _JointLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fx, fy, w2, b2, ymat, labels, x_lens, y_lens, blank, size_average):
        from .. import ops
        _lib.require_cuda(fx, "fx")
        lib = _lib.load()
        B, T, H = fx.shape
        U1 = fy.shape[1]
        V1 = w2.shape[0]
        dev = fx.device
        fxc, fyc = fx.detach().float().contiguous(), fy.detach().float().contiguous()
        w2b = w2.detach().to(torch.bfloat16).contiguous()
        b2c = b2.detach().float().contiguous()
        ym = ymat.detach().to(dev, torch.int32).contiguous()
        lab, offs, ylen, xlen = _pack_labels(labels, x_lens, y_lens, B, U1, dev)
        nodes = T * B * U1
        lat = torch.empty(nodes, 2, dtype=torch.float32, device=dev)
        sp = _lib.stream_ptr()
        ops._launch("rnnt_joint", 2.0 * nodes * H * V1,
                    lambda: lib.sb_rnnt_joint_fwd(fxc.data_ptr(), fyc.data_ptr(), w2b.data_ptr(),
                                                  b2c.data_ptr(), ym.data_ptr(), lat.data_ptr(),
                                                  None, B, T, U1, H, V1, int(blank), sp))
        need = any(ctx.needs_input_grad[:4])
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.sb_rnnt_workspace_size(B, T, U1, ctypes.byref(nbytes)), "ws")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        costs = torch.empty(B, dtype=torch.float32, device=dev)
        garc = torch.empty(nodes, 2, dtype=torch.float32, device=dev) if need else None
        ops._launch("rnnt_fwd_bwd", 0.0,
                    lambda: lib.sb_rnnt_fwd_bwd_compact(lat.data_ptr(), _lib.ptr(garc),
                                                        lab.data_ptr(), offs.data_ptr(),
                                                        ylen.data_ptr(), xlen.data_ptr(), B, T, U1,
                                                        int(blank), costs.data_ptr(), ws.data_ptr(),
                                                        nbytes.value, sp))
        loss = costs.sum().reshape(1)
        scale = 1.0 / B if size_average else 1.0
        ctx.saved = (fxc, fyc, w2b, b2c, ym, garc)
        ctx.dims = (B, T, U1, H, V1, int(blank), scale)
        return loss * scale

    @staticmethod
    def backward(ctx, grad_out):
        from .. import ops
        lib = _lib.load()
        fxc, fyc, w2b, b2c, ym, garc = ctx.saved
        B, T, U1, H, V1, blank, scale = ctx.dims
        dev = fxc.device
        if garc is None:
            return (None,) * 10
        garc = garc * (grad_out.reshape(1, 1).float() * scale)
        nodes = T * B * U1
        NV = 32 if V1 <= 32 else 64
        dlog = torch.empty(nodes, NV, dtype=torch.bfloat16, device=dev)
        db2 = torch.zeros(V1, dtype=torch.float32, device=dev)
        sp = _lib.stream_ptr()
        # recompute pass: logits -> softmax -> gradient w.r.t. the logits (bf16 rows)
        ops._launch("rnnt_joint", 2.0 * nodes * H * V1,
                    lambda: lib.sb_rnnt_joint_dlogits(fxc.data_ptr(), fyc.data_ptr(), w2b.data_ptr(),
                                                      b2c.data_ptr(), ym.data_ptr(), garc.data_ptr(),
                                                      dlog.data_ptr(), db2.data_ptr(), B, T, U1, H,
                                                      V1, blank, sp))
        # weight / input gradients, a slab of frames at a time: the hidden activations of Tc
        # frames (z, bf16) and their gradient (dz, f32) are the only (.., H)-wide temporaries
        Tc = max(1, min(JOINT_SLAB_FRAMES, T))
        rows_max = Tc * B * U1
        z = torch.empty(rows_max, H, dtype=torch.bfloat16, device=dev)
        dz = torch.empty(rows_max, H, dtype=torch.float32, device=dev)
        dw2 = torch.zeros(NV, H, dtype=torch.float32, device=dev)
        dfx = torch.empty(B, T, H, dtype=torch.float32, device=dev)
        dfy = torch.zeros(B, U1, H, dtype=torch.float32, device=dev)
        w2p = torch.zeros(NV, H, dtype=torch.bfloat16, device=dev)
        w2p[:V1] = w2b
        for t0 in range(0, T, Tc):
            tc = min(Tc, T - t0)
            rows = tc * B * U1
            dl = dlog[t0 * B * U1:t0 * B * U1 + rows]
            ops._launch("rnnt_joint_slab", 0.0,
                        lambda: lib.sb_rnnt_joint_build_slab(fxc.data_ptr(), fyc.data_ptr(),
                                                             z.data_ptr(), B, T, U1, H, t0, tc, sp))
            # dW2 (NV x H) += dlogits^T z   (token-major operands, MN-major UMMA)
            ops.gemm_bf16_tn(dl, z[:rows], out=dw2, accumulate=True, a_mn=True, b_mn=True)
            # dz (rows x H) = dlogits W2    (W2 is the [K][N] form of the B operand)
            ops.gemm_bf16_tn(dl, w2p, out=dz[:rows], b_mn=True)
            ops._launch("rnnt_joint_slab", 0.0,
                        lambda: lib.sb_rnnt_joint_reduce_slab(dz.data_ptr(), z.data_ptr(),
                                                              dfx.data_ptr(), dfy.data_ptr(), B, T,
                                                              U1, H, t0, tc, sp))
        return dfx, dfy, dw2[:V1], db2, None, None, None, None, None, None


class JointTransducerLoss(torch.nn.Module):
    """fc2 + log-softmax + transducer loss on fx = fc1(x), fy = fc1(pred) without the 4-D tensors:
    loss = JointTransducerLoss()(fx, fy, fc2.weight, fc2.bias, ymat, labels, x_lens, y_lens)."""

    def __init__(self, blank=None, size_average=False):
        super().__init__()
        self.blank = blank
        self.size_average = size_average

    def forward(self, fx, fy, w2, b2, ymat, labels, x_lens, y_lens):
        blank = w2.shape[0] - 1 if self.blank is None else self.blank
        return _JointLossFunction.apply(fx, fy, w2, b2, ymat, labels, x_lens, y_lens, blank,
                                        self.size_average)
