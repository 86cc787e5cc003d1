"""Encoder modules with the reference's signatures, kept THIN: the ResNet-152 trunk is
torchvision's (not on the hot path, run under no_grad exactly like the reference), only the tail that hands the
features to the decoder runs on libsn100 kernels.

  EncoderCNN(embed_size)               stylenet/model.py:11-27 / nic/model.py:10-27
        trunk (-> [B, 2048, 1, 1]) -> Linear(2048, E) -> BatchNorm1d(E, momentum=0.01)            -> [B, E]
  EncoderCNNAtt(encoded_image_size=14)  stylenet/model_att.py:11-29 / nic/model_att.py (class EncoderCNN there)
        trunk (-> [B, 2048, h, w]) -> AdaptiveAvgPool2d((S, S)) -> permute(0, 2, 3, 1)             -> [B, S, S, 2048]

Parameter / buffer names equal the reference's (``resnet.*``, ``linear.*``, ``bn.*``) so encoder checkpoints
interchange.  ``linear`` and ``bn`` are trained by the reference's ``optimizer`` (train_multitask.py:163-164): pass
them to ``FusedClampAdam(..., extra_params=...)``.

What the fused tail changes against the reference: the attention encoder returns a CONTIGUOUS NHWC tensor produced in
one pass over the trunk output (the reference returns a permuted view that the decoder has to copy), and attaches the
mean over pixels (``features.sn_mean``, what init_hidden_state needs) and, on request, the bf16 copy the tensor-core
GEMMs consume -- the decoder picks them up instead of re-reading the largest tensor on the path.
Data parallel note: BatchNorm1d statistics are per rank (like torch without SyncBatchNorm)."""
import torch
import torch.nn as nn

from . import ops


def _trunk(cut, pretrained, backbone):
    if backbone is not None:
        return backbone, None
    import torchvision.models as models
    try:
        weights = models.ResNet152_Weights.IMAGENET1K_V1 if pretrained else None
        resnet = models.resnet152(weights=weights)
    except Exception as e:          # no network / no cached checkpoint
        raise RuntimeError("EncoderCNN(pretrained=True) needs the torchvision ResNet-152 checkpoint in the local cache "
                           "(no network here); pass pretrained=False and load a state_dict, or pass backbone=") from e
    return nn.Sequential(*list(resnet.children())[:cut]), resnet.fc.in_features


class _LinearBNFn(torch.autograd.Function):
    """bn(linear(x)) on the kernels (stylenet/model.py:25-26): sn_gemm + sn_bn1d_fwd; backward sn_bn1d_bwd + GEMMs."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, run_mean, run_var, momentum, eps, training):
        x = x.detach().float().contiguous()
        B, E = x.shape[0], w.shape[0]
        lin = ops.linear_nt(x, w.detach(), b.detach())
        y = torch.empty_like(lin)
        sm = torch.empty(E, dtype=torch.float32, device=x.device)
        si = torch.empty(E, dtype=torch.float32, device=x.device)
        ops.bn1d_fwd(lin, gamma.detach(), beta.detach(), run_mean, run_var, momentum, eps, training, y, sm, si)
        ctx.save_for_backward(x, w, gamma, lin, sm, si)
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, gamma, lin, sm, si = ctx.saved_tensors
        dy = dy.contiguous()
        B, E = dy.shape
        K = x.shape[1]
        dlin = torch.empty_like(dy)
        dgamma = torch.empty(E, dtype=torch.float32, device=dy.device)
        dbeta = torch.empty(E, dtype=torch.float32, device=dy.device)
        ops.bn1d_bwd(lin, dy, gamma.detach(), sm, si, ctx.training, dlin, dgamma, dbeta)
        dw = torch.empty(E, K, dtype=torch.float32, device=dy.device)
        ops.gemm(ops.OP_TN, dlin, x, dw, E, K, B, E, K, K)
        db = torch.empty(E, dtype=torch.float32, device=dy.device)
        ops.colsum(dlin, B, E, E, db)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, K, dtype=torch.float32, device=dy.device)
            ops.gemm(ops.OP_NN, dlin, w.detach(), dx, B, K, E, E, K, K)
        return dx, dw, db, dgamma, dbeta, None, None, None, None, None


class EncoderCNN(nn.Module):
    """Signature of stylenet/model.py:13 (``EncoderCNN(embed_size)``); ``pretrained`` / ``backbone`` are additions:
    the reference always downloads the ImageNet checkpoint, which a box without network cannot."""

    def __init__(self, embed_size, pretrained=True, backbone=None, in_features=None):
        super().__init__()
        self.resnet, feat = _trunk(-1, pretrained, backbone)
        self.linear = nn.Linear(in_features or feat or 2048, embed_size)
        self.bn = nn.BatchNorm1d(embed_size, momentum=0.01)

    def forward(self, images):
        with torch.no_grad():
            features = self.resnet(images)
        features = features.reshape(features.size(0), -1)
        if not features.is_cuda:
            raise ops._lib.SnError("EncoderCNN tail runs on libsn100 kernels: inputs must be CUDA tensors (no CPU fallback)")
        bn = self.bn
        if bn.training and bn.track_running_stats:
            bn.num_batches_tracked += 1
        use_batch = bn.training or not bn.track_running_stats
        return _LinearBNFn.apply(features, self.linear.weight, self.linear.bias, bn.weight, bn.bias, bn.running_mean,
                                 bn.running_var, bn.momentum, bn.eps, use_batch)


class _PoolNHWCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, S, want_bf16):
        x = x.detach().float().contiguous()
        B, D, h, w = x.shape
        out = torch.empty(B, S, S, D, dtype=torch.float32, device=x.device)
        mean = torch.empty(B, D, dtype=torch.float32, device=x.device)
        outb = torch.empty(B, S, S, D, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
        ops.pool_nhwc_fwd(x, S, out, outb, mean)
        ctx.hw = (h, w)
        ctx.mark_non_differentiable(mean)
        if outb is not None:
            ctx.mark_non_differentiable(outb)
            return out, mean, outb
        return out, mean

    @staticmethod
    def backward(ctx, dout, *unused):
        h, w = ctx.hw
        B, S, _, D = dout.shape
        dx = torch.empty(B, D, h, w, dtype=torch.float32, device=dout.device)
        ops.pool_nhwc_bwd(dout.contiguous(), h, w, dx)
        return dx, None, None


class EncoderCNNAtt(nn.Module):
    """``EncoderCNN(encoded_image_size=14)`` of stylenet/model_att.py:11-29.  Returns [B, S, S, 2048] (contiguous)
    with ``.sn_mean`` [B, 2048] (and ``.sn_bf16`` when ``emit_bf16``) attached for the decoder."""

    def __init__(self, encoded_image_size=14, pretrained=True, backbone=None, emit_bf16=False):
        super().__init__()
        self.resnet, _ = _trunk(-2, pretrained, backbone)
        self.encoded_image_size = encoded_image_size
        self.adaptive_pool = nn.AdaptiveAvgPool2d((encoded_image_size, encoded_image_size))   # kept for its name / repr
        self.emit_bf16 = emit_bf16

    def forward(self, images):
        with torch.no_grad():
            features = self.resnet(images)
        if not features.is_cuda:
            raise ops._lib.SnError("EncoderCNNAtt tail runs on libsn100 kernels: inputs must be CUDA tensors (no CPU fallback)")
        res = _PoolNHWCFn.apply(features, self.encoded_image_size, self.emit_bf16)
        out = res[0]
        out.sn_mean = res[1]
        if self.emit_bf16:
            out.sn_bf16 = res[2]
        return out
