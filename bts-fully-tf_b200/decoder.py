"""decoder_model(...) of the reference (bts_decoder.py:26-105) wired around the fused LPG ops.

Scope note: the 3x3 / dilated convolutions, BatchNorm and up-sampling of the
decoder body are dense-conv, library territory -- they run on torch/cuDNN here exactly as they run
on TF/cuDNN in the reference and are NOT part of the hand-written hot path.  What this module
replaces is bts_decoder.py:79-81, 86-88, 93-94: each `reduction_NxN` Conv2D(3,1x1,sigmoid) +
`LocalPlanarGuidance` + down-sampling Lambda becomes ONE kernel forward and ONE backward
(layers.ReductionLPG -> libbtslpg.so), which also removes the ~thousand-node split/concat storm
that K.repeat_elements generates.

Interface = the reference's: `decoder_model(decoder_inputs, max_depth, num_filters, is_training)`
with NHWC tensors `[dense_features, skip_2, skip_4, skip_8, skip_16]` -> depth_est (B,H,W,1).
Internally activations are NCHW tensors in channels_last memory format, so the NHWC views the
LPG kernels need are zero-copy.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .layers import ReductionLPG

BN_EPS = 1.1e-5          # bts_decoder.py:27
BN_MOMENTUM = 1 - 0.99   # Keras momentum 0.99 == torch momentum 0.01


def _conv(cin, cout, k=3, dilation=1):
    pad = dilation * (k - 1) // 2          # padding='same', stride 1
    return nn.Conv2d(cin, cout, k, padding=pad, dilation=dilation, bias=False)


def _bn(c):
    return nn.BatchNorm2d(c, eps=BN_EPS, momentum=BN_MOMENTUM)


def _nhwc_view(x):
    """NCHW channels_last tensor -> contiguous NHWC view (no copy)."""
    return x.permute(0, 2, 3, 1)


def _upsample2x(x):
    """UpSampling2D(size=2, 'nearest') (bts_decoder.py:31, :38, :97) on an NCHW / channels_last tensor through
    the hand-written kernel (ops.upsample2x_nhwc); the result is again NCHW in channels_last memory."""
    x_nhwc = _nhwc_view(x.contiguous(memory_format=torch.channels_last))
    return _to_nchw(ops.upsample2x_nhwc(x_nhwc))


def _to_nchw(x_nhwc):
    """NHWC tensor -> NCHW view in channels_last memory format (no copy when x is contiguous)."""
    return x_nhwc.permute(0, 3, 1, 2)


TENSOR_CORE_WGRAD = True     # training: d kernel of the 3x3 convolutions with few channels on ops.conv3x3_wgrad (module-wide switch)


class Conv3x3Function(torch.autograd.Function):
    """Conv2D(3x3, stride 1, padding='same', no bias): forward and d input from the library, d kernel from the tcgen05 kernel
    (ops.conv3x3_wgrad) -- the library's weight-gradient kernels run at 2-8 times its time when the output has few channels and the
    reduction runs over millions of pixels (profiles/r02_wgrad_tcgen05.json)."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        return F.conv2d(x, weight, None, 1, 1, 1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous(memory_format=torch.channels_last)
        g_x = g_w = None
        if ctx.needs_input_grad[0]:
            g_x = torch.ops.aten.convolution_backward(g, x, weight, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False])[0]
        if ctx.needs_input_grad[1]:
            hwio = ops.conv3x3_wgrad(_nhwc_view(x.contiguous(memory_format=torch.channels_last)), _nhwc_view(g))
            g_w = hwio.permute(3, 2, 0, 1)                                   # HWIO -> OIHW (a view; AccumulateGrad adds it into .grad)
        return g_x, g_w


def _tc_wgrad_applies(x, weight):
    """TF32 like the library kernel it replaces (so only while torch's own cuDNN TF32 switch is on); channel counts the kernel takes and
    at which it measured faster than the library (Cin * Cout <= 12288: not 128 x 128), enough pixels to fill the machine."""
    cout, cin, kh, kw = weight.shape
    return (TENSOR_CORE_WGRAD and torch.is_grad_enabled() and weight.requires_grad and x.is_cuda and x.dtype == torch.float32
            and torch.backends.cudnn.allow_tf32 and (kh, kw) == (3, 3) and cin % 4 == 0 and cout % 4 == 0 and cin <= 256 and cout <= 128
            and cin * cout <= 12288 and x.shape[0] * x.shape[2] * x.shape[3] >= 32768)


def _conv3x3(x_nchw, weight):
    if _tc_wgrad_applies(x_nchw, weight):
        return Conv3x3Function.apply(x_nchw, weight)
    return F.conv2d(x_nchw, weight, None, 1, 1, 1)


def _conv_padded_input(conv, x_nchw, pad):
    """conv(x) for an input that carries `pad` extra zero channels: the kernel gets matching zero input channels
    (identical result; the gradient of the padding is dropped by autograd's slice)."""
    w = F.pad(conv.weight, (0, 0, 0, 0, 0, pad)) if pad else conv.weight
    if conv.kernel_size == (3, 3) and conv.dilation == (1, 1) and conv.stride == (1, 1) and conv.padding == (1, 1):
        return _conv3x3(x_nchw, w)
    return F.conv2d(x_nchw, w, None, conv.stride, conv.padding, conv.dilation)


def _bn_affine(bn):
    """Inference-mode BatchNormalization as y = x * scale + shift (float32 [C] each)."""
    scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
    return scale.contiguous(), (bn.bias - bn.running_mean * scale).contiguous()


class _ConvBlock(nn.Module):
    """conv_block / conv_block_no_lpg (bts_decoder.py:30-44): upsample x2, 3x3 conv + ELU, BN, concat, 3x3 conv + ELU.

    The memory-bound glue between the two convolutions runs on the hand-written kernels: the up-sampling
    (ops.upsample2x_nhwc) and ONE pass that writes the concat [up, skip, lpg_ds] (ops.concat_*), padded with zero
    channels to a multiple of 4 so that cuDNN does not re-copy the tensor; in inference mode that pass also applies
    the ELU of the upconv and the folded BatchNormalization."""

    def __init__(self, cin, cskip, clpg, nf, subpixel_inference=False):
        super().__init__()
        self.upconv = _conv(cin, nf)
        self.bn = _bn(nf)
        self.iconv = _conv(nf + cskip + clpg, nf)
        self.pad = ops.pad_to(nf + cskip + clpg)
        # inference: evaluate the upconv on the low-res input (ops.subpixel_kernel) where cuDNN gains from it (block2: 0.76 -> 0.57 ms)
        self.subpixel_inference = bool(subpixel_inference) and nf % 4 == 0
        self.fused_training_glue = True     # training: ELU + BatchNorm (batch statistics) + concat on the hand-written kernels

    def forward(self, x, skip_nhwc, lpg_nhwc=None):
        planes = [lpg_nhwc] if lpg_nhwc is not None else []            # order is load-bearing (bts_decoder.py:42)
        if not self.training and not torch.is_grad_enabled():
            scale, shift = _bn_affine(self.bn)
            if self.subpixel_inference:
                up = F.conv2d(x, ops.subpixel_kernel(self.upconv.weight), padding=1)
            else:
                up = self.upconv(_upsample2x(x))
            cat = ops.concat_forward(_nhwc_view(up.contiguous(memory_format=torch.channels_last)), planes, skip_nhwc.contiguous(), act=True,
                                     pad=self.pad, scale=scale, shift=shift, a_subpixel=self.subpixel_inference)
            return F.elu(_conv_padded_input(self.iconv, _to_nchw(cat), self.pad))
        raw = _conv3x3(_upsample2x(x), self.upconv.weight)
        nf = raw.shape[1]
        if self.training and self.fused_training_glue and ops.bn_glue_supported(nf, nf + skip_nhwc.shape[-1] + len(planes) + self.pad, raw.dtype):
            # training: ELU + BatchNormalization (batch statistics, moving averages) + concat as a statistics pass and ONE fused pass,
            # forward and backward (ops.conv_block_glue); the framework runs three read+write passes each way
            cat = ops.conv_block_glue(_nhwc_view(raw.contiguous(memory_format=torch.channels_last)), skip_nhwc, planes, self.bn, pad=self.pad)
            if self.bn.num_batches_tracked is not None:
                self.bn.num_batches_tracked.add_(1)
        else:
            up = self.bn(F.elu(raw))                                    # autograd in eval mode / unsupported channel counts: the framework's BatchNorm
            cat = ops.concat_nhwc(_nhwc_view(up.contiguous(memory_format=torch.channels_last)), planes, b=skip_nhwc, act=False, pad=self.pad)
        return F.elu(_conv_padded_input(self.iconv, _to_nchw(cat), self.pad))


class _DenseAspp(nn.Module):
    """dense_aspp_block (bts_decoder.py:46-54)."""

    def __init__(self, cin, nf, rate, batch_norm_first=True):
        super().__init__()
        self.bn_first = _bn(cin) if batch_norm_first else None
        self.conv1 = _conv(cin, nf, k=1)
        self.bn2 = _bn(nf)
        self.conv2 = _conv(nf, nf // 2, k=3, dilation=rate)

    def forward(self, x):
        if self.bn_first is not None:
            x = self.bn_first(x)
        x = self.conv1(F.relu(x))
        x = F.relu(self.bn2(x)).contiguous(memory_format=torch.channels_last)
        return _to_nchw(_conv_nhwc(_nhwc_view(x), self.conv2))

    def tail_inference(self, x_relu_nhwc, split=0):
        """conv1 -> BN -> ReLU -> dilated conv2 on an input that already went through (BN and) ReLU, inference mode.
        The 1x1 convolution, the folded second BatchNormalization and the ReLU are ONE library call (cuDNN's fused convolution +
        bias + ReLU with the BatchNormalization scale folded into the kernel: 55-111 us per block at B = 32, 60x80, where the plain
        convolution followed by one in-place ops.affine_act pass took 114-167 us, tools/exp_conv_bias_relu.py).
        split = s: x_relu_nhwc is in the sub-grid form of _s2b(., s) (the 1x1 convolution is pointwise, so its output is too) and so is
        the returned tensor -- the rate-18 / 24 convolutions, see _dilation_split."""
        scale, shift = _bn_affine(self.bn2)
        x = _to_nchw(x_relu_nhwc)
        if x.is_cuda and hasattr(torch, "cudnn_convolution_relu"):
            wf = (self.conv1.weight * scale[:, None, None, None]).contiguous(memory_format=torch.channels_last)
            y = torch.cudnn_convolution_relu(x, wf, shift, [1, 1], [0, 0], [1, 1], 1).contiguous(memory_format=torch.channels_last)
            y_nhwc = _nhwc_view(y)
        else:
            y_nhwc = _nhwc_view(self.conv1(x).contiguous(memory_format=torch.channels_last))
            ops.affine_act(y_nhwc, dst=y_nhwc, scale=scale, shift=shift, act=ops.ACT_RELU)
        return _conv_nhwc(y_nhwc, self.conv2, split=split)


SPLIT_DILATION_FROM = 16     # dilation rates from here on run as 2 x 2 interleaved sub-grids of half the rate (see _dilation_split)


def _dilation_split(conv, H, W):
    """The library's fast convolution kernels stop at a dilation rate between 12 and 18: the DenseASPP convolutions of rate 18 and 24
    (bts_decoder.py:53, :70, :73) fall to kernels 3.5x (forward) / 4.6x (weight gradient) slower than those of rate 3-12
    (tools/exp_dilated_wgrad.py).  A rate-d convolution is exactly s x s independent rate-d/s convolutions on the interleaved
    sub-grids x[:, i::s, j::s] (what TensorFlow's atrous convolution does with space_to_batch): with s = 2 the rates become 9 and 12.
    Returns s (1: run the layer as it is)."""
    d = conv.dilation[0]
    return 2 if (d >= SPLIT_DILATION_FROM and d % 2 == 0 and H % 2 == 0 and W % 2 == 0 and conv.dilation[1] == d) else 1


def _s2b(x_nhwc, s):
    """(B,H,W,C) -> (B*s*s, H/s, W/s, C): sub-grid (i, j) of image b becomes image (b*s + i)*s + j."""
    B, H, W, C = x_nhwc.shape
    return x_nhwc.reshape(B, H // s, s, W // s, s, C).permute(0, 2, 4, 1, 3, 5).reshape(B * s * s, H // s, W // s, C)


def _b2s(y_nhwc, s):
    Bs, h, w, C = y_nhwc.shape
    return y_nhwc.reshape(Bs // (s * s), s, s, h, w, C).permute(0, 3, 1, 4, 2, 5).reshape(Bs // (s * s), h * s, w * s, C)


def _conv_backward(g_nhwc, x_nhwc, conv, split=0):
    """(d input, d weight) of a bias-free stride-1 convolution, NHWC tensors in and out (cuDNN through the framework's own entry).
    split = s: g_nhwc, x_nhwc and the returned d input are all in the sub-grid form of _s2b(., s) (the caller re-orders inside passes
    it runs anyway, ops.affine_act(..., src_split / dst_split)); split = 0: normal layout, split internally where it pays."""
    s = split or _dilation_split(conv, x_nhwc.shape[1], x_nhwc.shape[2])
    if s > 1 and not split:
        g_nhwc, x_nhwc = _s2b(g_nhwc, s), _s2b(x_nhwc, s)
    dil = [conv.dilation[0] // s, conv.dilation[1] // s]
    pad = list(conv.padding) if s == 1 else dil
    g_in, g_w, _ = torch.ops.aten.convolution_backward(_to_nchw(g_nhwc), _to_nchw(x_nhwc), conv.weight, None, [1, 1], pad, dil, False, [0, 0], 1,
                                                       [True, True, False])
    g_in = _nhwc_view(g_in.contiguous(memory_format=torch.channels_last))
    return (_b2s(g_in, s) if (s > 1 and not split) else g_in), g_w


def _conv_nhwc(x_nhwc, conv, split=0):
    """conv(x) for a bias-free stride-1 'same' convolution on an NHWC tensor; differentiable (plain framework ops).
    split = s: x_nhwc and the result are in the sub-grid form of _s2b(., s); split = 0: normal layout, split internally where it pays."""
    s = split or _dilation_split(conv, x_nhwc.shape[1], x_nhwc.shape[2])
    if s > 1:
        d = conv.dilation[0] // s
        y = _nhwc_view(F.conv2d(_to_nchw(x_nhwc if split else _s2b(x_nhwc, s)), conv.weight, None, 1, d, d).contiguous(memory_format=torch.channels_last))
        return y if split else _b2s(y, s)
    return _nhwc_view(F.conv2d(_to_nchw(x_nhwc), conv.weight, None, 1, conv.padding, conv.dilation).contiguous(memory_format=torch.channels_last))


class DenseAsppTrainFunction(torch.autograd.Function):
    """The DenseASPP of bts_decoder.py:46-76 in TRAINING mode, iconv4 -> concat4_daspp, on ONE (B,h,w,nf + 5 nf/2) buffer the blocks
    append to.  The convolutions are cuDNN; everything between them is the hand-written glue of csrc/bnrelu_kernels.cuh and
    csrc/slice_kernels.cuh:
      forward   per-channel batch moments are taken once per appended piece and shared by every BatchNormalization over a concat that
                contains it; Concatenate + BatchNormalization + ReLU in front of a 1x1 conv = one affine_act pass over a channel slice
      backward  ReLU' + BatchNormalization' = a statistics pass and an apply pass that ADDS into the slice of the shared gradient
                buffer (the Concatenates' backward).
    Inputs: iconv4 (NCHW), the decoder (for the layers' hyper-parameters and moving averages), then the parameters in
    BtsDecoder._daspp_params() order (so that their gradients come back through autograd)."""

    @staticmethod
    def forward(ctx, iconv4, dec, *params):
        blocks = dec._daspp_blocks()
        B, nf, h, w = iconv4.shape
        half, n = nf // 2, B * h * w
        CT = nf + 5 * half
        dev = iconv4.device
        x4 = _nhwc_view(iconv4.contiguous(memory_format=torch.channels_last))
        buf = torch.empty((B, h, w, CT), dtype=torch.float32, device=dev)          # [iconv4 | d3 | d6 | d12 | d18 | d24]
        mean, var = torch.empty(CT, dtype=torch.float32, device=dev), torch.empty(CT, dtype=torch.float32, device=dev)
        ops.affine_act(x4, dst=buf[..., :nf])
        ops.bn_moments(x4, mean[:nf], var[:nf])

        def fold(bn, m, v):
            return ops.bn_fold(m, v, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.momentum, bn.eps, n)

        vec4 = fold(dec.bn4, mean[:nf], var[:nf])
        saved = [x4, buf, *vec4]
        for k, blk in enumerate(blocks):
            ck = nf + half * k
            if k == 0:
                vk = vec4
                xk = ops.affine_act(x4, scale=vec4[0], shift=vec4[1], act=ops.ACT_RELU)              # relu(iconv4_bn)
            else:
                vk = fold(blk.bn_first, mean[:ck], var[:ck])
                xk = ops.affine_act(buf[..., :ck], scale=vk[0], shift=vk[1], act=ops.ACT_RELU)       # relu(BN(concat4_k)), contiguous
            t = _conv_nhwc(xk, blk.conv1)
            m2, v2 = torch.empty(nf, dtype=torch.float32, device=dev), torch.empty(nf, dtype=torch.float32, device=dev)
            ops.bn_moments(t, m2, v2)
            v2k = fold(blk.bn2, m2, v2)
            s2 = _dilation_split(blk.conv2, h, w)
            s2 = s2 if s2 > 1 else 0
            # relu(BN(conv1)); for the split rate-18 / 24 convolutions written in the sub-grid form they read (and their backward reads)
            r2 = ops.affine_act(t, scale=v2k[0], shift=v2k[1], act=ops.ACT_RELU, dst_split=s2)
            d = _conv_nhwc(r2, blk.conv2, split=s2)
            ops.affine_act(d, dst=buf[..., ck:ck + half], src_split=s2)
            if k + 1 < len(blocks):
                ops.bn_moments(d, mean[ck:ck + half], var[ck:ck + half])      # per-channel sums: the pixel order does not matter
            saved += [xk, t, r2, *v2k] + ([] if k == 0 else list(vk))
        ops.affine_act(x4, dst=buf[..., :nf], scale=vec4[0], shift=vec4[1])                          # concat4_daspp starts with iconv4_bn (:75)
        ctx.save_for_backward(*saved)
        ctx.dec = dec
        return buf

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        dec = ctx.dec
        blocks = dec._daspp_blocks()
        sv = list(ctx.saved_tensors)
        x4, buf, vec4 = sv[0], sv[1], tuple(sv[2:6])
        pos = 6
        per_block = []
        for k in range(len(blocks)):
            xk, t, r2 = sv[pos:pos + 3]
            v2k = tuple(sv[pos + 3:pos + 7])
            pos += 7
            vk = vec4
            if k > 0:
                vk = tuple(sv[pos:pos + 4])
                pos += 4
            per_block.append((xk, t, r2, v2k, vk))
        B, h, w, CT = buf.shape
        nf = x4.shape[-1]
        half = nf // 2
        dev = buf.device
        # g_out, the gradient of concat4_daspp, is only READ; the blocks' contributions to the appended pieces d3 .. d18 are gathered in
        # gacc (channels [nf, nf + 4 half) of the buffer), started by the last block as g_out's slice + its contribution (dst_init)
        g_out = g_out.contiguous()
        last = len(blocks) - 1
        gacc = torch.empty((B, h, w, last * half), dtype=torch.float32, device=dev)
        g4 = torch.empty_like(x4)                      # d loss / d iconv4
        g4_written = False
        grads = {}

        def vec_slice(v, a, b):
            return tuple(u[a:b] for u in v)

        for k in range(len(blocks) - 1, -1, -1):
            blk = blocks[k]
            xk, t, r2, v2k, vk = per_block[k]
            ck = nf + half * k
            s2 = _dilation_split(blk.conv2, h, w)
            s2 = s2 if s2 > 1 else 0
            g_slot = g_out[..., ck:ck + half] if k == last else gacc[..., ck - nf:ck - nf + half]
            g_d = ops.affine_act(g_slot, dst_split=s2)                                  # contiguous copy of the block's complete gradient
            g_r2, grads[blk.conv2.weight] = _conv_backward(g_d, r2, blk.conv2, split=s2)
            if s2:
                g_r2 = ops.affine_act(g_r2, src_split=s2)                               # back to the pixel order of conv1's output
            gg2, gb2 = torch.empty(nf, dtype=torch.float32, device=dev), torch.empty(nf, dtype=torch.float32, device=dev)
            g_t = ops.bn_act_backward(g_r2, t, v2k, gg2, gb2, torch.empty_like(t))
            grads[blk.bn2.weight], grads[blk.bn2.bias] = gg2, gb2
            g_xk, grads[blk.conv1.weight] = _conv_backward(g_t, xk, blk.conv1)
            if k == 0:
                # iconv4_bn feeds relu -> conv1 of daspp_3 AND concat4_daspp directly (:58-59, :75): one BatchNormalization backward
                gg, gb = torch.empty(nf, dtype=torch.float32, device=dev), torch.empty(nf, dtype=torch.float32, device=dev)
                ops.bn_act_backward(g_xk, x4, vec4, gg, gb, g4, accumulate=g4_written, g2=g_out[..., :nf])
                grads[dec.bn4.weight], grads[dec.bn4.bias] = gg, gb
            else:
                gg, gb = torch.empty(ck, dtype=torch.float32, device=dev), torch.empty(ck, dtype=torch.float32, device=dev)
                ops.bn_act_backward(g_xk[..., :nf], x4, vec_slice(vk, 0, nf), gg[:nf], gb[:nf], g4, accumulate=g4_written)
                g4_written = True
                ops.bn_act_backward(g_xk[..., nf:], buf[..., nf:ck], vec_slice(vk, nf, ck), gg[nf:], gb[nf:], gacc[..., :ck - nf],
                                    accumulate=k != last, dst_init=g_out[..., nf:ck] if k == last else None)
                grads[blk.bn_first.weight], grads[blk.bn_first.bias] = gg, gb
        g_iconv4 = _to_nchw(g4) if ctx.needs_input_grad[0] else None
        return (g_iconv4, None) + tuple(grads[p] for p in dec._daspp_params())


class BtsDecoder(nn.Module):
    """The decoder graph of bts_decoder.py with the three LPG heads fused.  Sub-modules are created in
    the reference's layer-creation order, so `conv_kernels()` lines up with a Keras weight list."""

    def __init__(self, in_channels, max_depth, num_filters=256):
        """in_channels: channel counts of [dense_features, skip_2, skip_4, skip_8, skip_16]."""
        super().__init__()
        c_dense, c2, c4, c8, c16 = in_channels
        self.max_depth = float(max_depth)
        nf = num_filters
        self.block5 = _ConvBlock(c_dense, c16, 0, nf)             # iconv5, H/16
        nf //= 2
        self.block4 = _ConvBlock(nf * 2, c8, 0, nf)               # iconv4, H/8
        self.bn4 = _bn(nf)
        self.daspp_3 = _DenseAspp(nf, nf, 3, batch_norm_first=False)
        self.daspp_6 = _DenseAspp(nf + nf // 2, nf, 6)
        self.daspp_12 = _DenseAspp(nf + 2 * (nf // 2), nf, 12)
        self.daspp_18 = _DenseAspp(nf + 3 * (nf // 2), nf, 18)
        self.daspp_24 = _DenseAspp(nf + 4 * (nf // 2), nf, 24)
        self.daspp_feat = _conv(nf + 5 * (nf // 2), nf // 2)
        self.reduction_8x8 = ReductionLPG(nf // 2, 8, ds_stride=4, name="reduction_8x8")
        c_daspp = nf // 2
        nf //= 2
        self.block3 = _ConvBlock(c_daspp, c4, 1, nf)              # iconv3, H/4
        self.reduction_4x4 = ReductionLPG(nf, 4, ds_stride=2, name="reduction_4x4")
        nf //= 2
        self.block2 = _ConvBlock(nf * 2, c2, 1, nf, subpixel_inference=True)   # iconv2, H/2
        self.reduction_2x2 = ReductionLPG(nf, 2, ds_stride=0, name="reduction_2x2")
        c_iconv2 = nf
        nf //= 2
        self.upconv1 = _conv(c_iconv2, nf)
        self.iconv1 = _conv(nf + 3, nf)
        self.depth_conv = _conv(nf, 1)
        self.intermediates = {}
        self.fused_training_glue = True         # training: DenseASPP BatchNorm / ReLU / concat glue on the hand-written kernels
        self.tensor_core_iconv1 = True          # inference: iconv1 on the tcgen05 kernel (TF32) while torch's cuDNN TF32 switch is on
        self._loss_ws = None
        self.reset_parameters_keras()

    def reset_parameters_keras(self):
        """Keras defaults of the reference's layers (bts_decoder.py builds every Conv2D without an initializer argument):
        kernel_initializer = glorot_uniform; BatchNormalization gamma = 1, beta = 0 (torch's defaults already)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)

    # --- weights in Keras order / layout -----------------------------------------------------
    def conv_modules(self):
        """Convolution-like modules in the reference's creation order (25 for bts_decoder.py)."""
        out = [self.block5.upconv, self.block5.iconv, self.block4.upconv, self.block4.iconv]
        for d in (self.daspp_3, self.daspp_6, self.daspp_12, self.daspp_18, self.daspp_24):
            out += [d.conv1, d.conv2]
        out += [self.daspp_feat, self.reduction_8x8, self.block3.upconv, self.block3.iconv, self.reduction_4x4,
                self.block2.upconv, self.block2.iconv, self.reduction_2x2, self.upconv1, self.iconv1, self.depth_conv]
        return out

    def load_keras_kernels(self, kernels):
        """kernels: list of HWIO arrays/tensors in creation order (Keras `layer.kernel`)."""
        mods = self.conv_modules()
        assert len(kernels) == len(mods), (len(kernels), len(mods))
        with torch.no_grad():
            for m, k in zip(mods, kernels):
                k = torch.as_tensor(k)
                if isinstance(m, ReductionLPG):
                    m.kernel.copy_(k.to(m.kernel))
                else:
                    m.weight.copy_(k.permute(3, 2, 0, 1).to(m.weight))      # HWIO -> OIHW

    def keras_kernel_grads(self):
        """Gradients of the conv kernels, HWIO, creation order."""
        out = []
        for m in self.conv_modules():
            g = m.kernel.grad if isinstance(m, ReductionLPG) else m.weight.grad.permute(2, 3, 1, 0)
            out.append(g)
        return out

    # --- forward -------------------------------------------------------------------------------
    def forward(self, decoder_inputs, return_logit=False):
        """decoder_inputs: NHWC [dense_features, skip_2, skip_4, skip_8, skip_16] -> depth_est NHWC (B,H,W,1)
        (return_logit: the pre-activation of the last Conv2D instead, for the fused loss)."""
        dense = _to_nchw(decoder_inputs[0])
        s2, s4, s8, s16 = decoder_inputs[1:]                               # skips stay NHWC: the concat kernel reads them as they are
        iconv5 = self.block5(dense, s16)
        iconv4 = self.block4(iconv5, s8)
        if not self.training and not torch.is_grad_enabled():
            daspp_feat = self._daspp_inference(iconv4)
        else:
            daspp_feat = self._daspp(iconv4)
        return self._tail(daspp_feat, s2, s4, return_logit)

    def _daspp_inference(self, iconv4):
        """DenseASPP (bts_decoder.py:46-77) in inference mode on ONE (B,h,w,896) buffer that the blocks append to:
        the reference's Concatenate + BatchNormalization + ReLU in front of every 1x1 conv (three passes over a
        growing map) are one ops.affine_act pass over a channel slice; see csrc/slice_kernels.cuh."""
        B, nf, h, w = iconv4.shape
        half = nf // 2
        buf = torch.empty((B, h, w, nf + 5 * half), dtype=iconv4.dtype, device=iconv4.device)       # [iconv4 | d3 | d6 | d12 | d18 | d24]
        ops.affine_act(_nhwc_view(iconv4.contiguous(memory_format=torch.channels_last)), dst=buf[..., :nf])
        s4_, t4_ = _bn_affine(self.bn4)
        blocks = self._daspp_blocks()
        for k, blk in enumerate(blocks):
            ck = nf + half * k
            sc, sh = (s4_, t4_) if k == 0 else _bn_affine(blk.bn_first)
            s2 = _dilation_split(blk.conv2, h, w)
            s2 = s2 if s2 > 1 else 0
            # relu(iconv4_bn) / relu(BN(concat4_k)), contiguous -- in sub-grid form for the blocks whose dilated convolution is split
            x = ops.affine_act(buf[..., :ck], scale=sc, shift=sh, act=ops.ACT_RELU, dst_split=s2)
            d = blk.tail_inference(x, split=s2)
            ops.affine_act(d, dst=buf[..., ck:ck + half], src_split=s2)
        ops.affine_act(buf[..., :nf], dst=buf[..., :nf], scale=s4_, shift=t4_)                       # concat4_daspp starts with iconv4_bn (:75)
        return F.elu(self.daspp_feat(_to_nchw(buf)))

    def _daspp_blocks(self):
        return (self.daspp_3, self.daspp_6, self.daspp_12, self.daspp_18, self.daspp_24)

    def _daspp_bns(self):
        out = [self.bn4]
        for blk in self._daspp_blocks():
            out += ([blk.bn_first] if blk.bn_first is not None else []) + [blk.bn2]
        return out

    def _daspp_params(self):
        """Parameters DenseAsppTrainFunction differentiates, in the order it returns their gradients."""
        out = [self.bn4.weight, self.bn4.bias]
        for blk in self._daspp_blocks():
            if blk.bn_first is not None:
                out += [blk.bn_first.weight, blk.bn_first.bias]
            out += [blk.conv1.weight, blk.bn2.weight, blk.bn2.bias, blk.conv2.weight]
        return out

    def _daspp(self, iconv4):
        if self.training and self.fused_training_glue and ops.bn_slices_supported(iconv4.shape[1], iconv4.dtype):
            buf = DenseAsppTrainFunction.apply(iconv4, self, *self._daspp_params())
            torch._foreach_add_([bn.num_batches_tracked for bn in self._daspp_bns()], 1)
            return F.elu(self.daspp_feat(_to_nchw(buf)))
        iconv4_bn = self.bn4(iconv4)
        d3 = self.daspp_3(iconv4_bn)
        c2 = torch.cat([iconv4, d3], 1)
        d6 = self.daspp_6(c2)
        c3 = torch.cat([c2, d6], 1)
        d12 = self.daspp_12(c3)
        c4 = torch.cat([c3, d12], 1)
        d18 = self.daspp_18(c4)
        c5 = torch.cat([c4, d18], 1)
        d24 = self.daspp_24(c5)
        return F.elu(self.daspp_feat(torch.cat([iconv4_bn, d3, d6, d12, d18, d24], 1)))

    def _tail(self, daspp_feat, s2, s4, return_logit):
        # bts_decoder.py:79-81: reduction_8x8 -> depth_8x8_scaled -> ds  (one kernel)
        red8, d8, d8_ds = self.reduction_8x8(_nhwc_view(daspp_feat.contiguous(memory_format=torch.channels_last)))
        iconv3 = self.block3(daspp_feat, s4, d8_ds)
        red4, d4, d4_ds = self.reduction_4x4(_nhwc_view(iconv3.contiguous(memory_format=torch.channels_last)))
        iconv2 = self.block2(iconv3, s2, d4_ds)
        red2, d2 = self.reduction_2x2(_nhwc_view(iconv2.contiguous(memory_format=torch.channels_last)))

        # bts_decoder.py:98-99: upconv1's ELU and concat1 = [upconv1, d2, d4, d8] as ONE pass (ops.concat_nhwc):
        # the raw conv output is read once and the F/16+3 channel NHWC pixel written once, LPG planes in their slots
        # bts_decoder.py:97-98 without the up-sampled tensor: UpSampling2D(2) + Conv2D(3x3) == the 3x3 conv on the LOW-RES
        # iconv2 with 4*Cout combined kernels (ops.subpixel_kernel) + a pixel shuffle, which the concat kernel does in its
        # addressing.  cuDNN: forward 0.77 -> 0.48 ms, forward+backward 3.0 -> 1.7 ms (B = 16, 352x1216).
        nf1 = self.upconv1.weight.shape[0]
        pad1 = ops.pad_to(nf1 + 3)                                         # 35 -> 36 channels
        self.intermediates = {"reduction_8x8": red8, "reduction_4x4": red4, "reduction_2x2": red2,
                              "depth_8x8_scaled": d8, "depth_4x4_scaled": d4, "depth_2x2_scaled": d2}
        if (not torch.is_grad_enabled() and nf1 in (16, 32) and self.tensor_core_iconv1 and torch.backends.cudnn.allow_tf32
                and iconv2.dtype == torch.float32):
            # bts_decoder.py:98-103, inference, WITHOUT concat1: iconv1's convolution is a tcgen05 implicit GEMM that stages
            # elu(upconv1) and the three LPG planes from their own buffers (ops.iconv1_forward), followed by the fused last
            # convolution.  TF32 operands like the library convolution it replaces -- hence only while the framework's own
            # TF32 switch (torch.backends.cudnn.allow_tf32, on by default) is on; with it off the float32 path below runs.
            up4 = F.conv2d(iconv2, ops.subpixel_kernel(self.upconv1.weight), padding=1)       # (B, 4*nf1, H/2, W/2)
            up4_nhwc = _nhwc_view(up4.contiguous(memory_format=torch.channels_last))
            x = ops.iconv1_forward(up4_nhwc, [d2, d4, d8], ops.kernel_hwio(self.iconv1.weight), a_subpixel=True)
            return ops.depthconv_forward(x, ops.kernel9c(self.depth_conv.weight), act_in=True,
                                         sigmoid_scale=None if return_logit else self.max_depth)
        if nf1 % 4 == 0:
            up4 = _conv3x3(iconv2, ops.subpixel_kernel(self.upconv1.weight))                  # (B, 4*nf1, H/2, W/2)
            up4_nhwc = _nhwc_view(up4.contiguous(memory_format=torch.channels_last))
            concat1 = _to_nchw(ops.concat_nhwc(up4_nhwc, [d2, d4, d8], act=True, pad=pad1, a_subpixel=True))
        else:
            up1_raw = self.upconv1(_upsample2x(iconv2))
            up1_nhwc = _nhwc_view(up1_raw.contiguous(memory_format=torch.channels_last))
            concat1 = _to_nchw(ops.concat_nhwc(up1_nhwc, [d2, d4, d8], act=True, pad=pad1))
        iconv1_raw = _conv_padded_input(self.iconv1, concat1, pad1)
        fused_tail = iconv1_raw.shape[1] in (16, 32)
        if fused_tail and not torch.is_grad_enabled():
            # bts_decoder.py:100-103 in ONE pass over the raw conv output: iconv1's ELU, the last Conv2D(1, 3x3) and,
            # unless the logit is asked for, sigmoid * max_depth (ops.depthconv_forward)
            x = _nhwc_view(iconv1_raw.contiguous(memory_format=torch.channels_last))
            return ops.depthconv_forward(x, ops.kernel9c(self.depth_conv.weight), act_in=True,
                                         sigmoid_scale=None if return_logit else self.max_depth)
        if fused_tail:
            # bts_decoder.py:100-102 with autograd: iconv1's ELU inside the hand-written forward, ELU' inside the ONE
            # hand-written pass that yields both gradients (ops.depth_conv): no ELU tensor, no ELU passes
            logit = _to_nchw(ops.depth_conv(_nhwc_view(iconv1_raw.contiguous(memory_format=torch.channels_last)), self.depth_conv.weight,
                                            act_in=True))
        else:
            logit = self.depth_conv(F.elu(iconv1_raw))                                 # (B,1,H,W): same memory as NHWC (B,H,W,1)
        if return_logit:
            return _nhwc_view(logit)
        if not torch.is_grad_enabled():                                                # bts_decoder.py:102-103 in one pass
            depth, _, _ = ops.silog_forward(_nhwc_view(logit.contiguous(memory_format=torch.channels_last)).contiguous(), None,
                                            self.max_depth, 0.0)
            return depth
        depth = torch.sigmoid(logit) * self.max_depth                                  # bts_decoder.py:102-103
        return _nhwc_view(depth)

    def forward_loss(self, decoder_inputs, y_true, dataset):
        """Training-step form: (depth_est, loss) with the last activation, the depth_est Lambda
        (bts_decoder.py:102-103) and si_log_loss (bts.py:27-41) as ONE kernel forward and ONE backward
        (losses.depth_silog); gradient flows from `loss` into the decoder."""
        from . import losses
        logit = self.forward(decoder_inputs, return_logit=True)
        ws = self._loss_ws
        if ws is None or ws.device != logit.device:
            ws = self._loss_ws = ops.tail_workspace(logit.device)          # one scratch for the life of the module, not one per step
        return losses.depth_silog(logit, y_true, self.max_depth, dataset, ws)


def decoder_model(decoder_inputs, max_depth, num_filters=256, is_training=False, decoder=None):
    """Same call as the reference's decoder_model (bts_decoder.py:26).  A BtsDecoder is built from the
    input channel counts on first use (pass `decoder=` to reuse weights); returns depth_est NHWC."""
    if decoder is None:
        decoder = BtsDecoder([t.shape[-1] for t in decoder_inputs], max_depth, num_filters).to(decoder_inputs[0].device)
    decoder.train(bool(is_training))
    return decoder(decoder_inputs)


def si_log_loss(y_true, y_pred, dataset="nyu"):
    """bts.py:27-41 (torch restatement for the training-step benches; not a hot-path kernel)."""
    th = {"nyu": 0.1, "kitti": 1.0, "matterport": 0.1}[dataset]
    mask = y_true > th
    d = torch.log(y_true[mask] + 1e-7) - torch.log(y_pred[mask] + 1e-7)
    return torch.sqrt((d * d).mean() - 0.85 * d.mean() ** 2) * 10.0
