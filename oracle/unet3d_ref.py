"""fp32 PyTorch restatement of ``deepsulci.deeptools.models.UNet3D``.

Test infrastructure (see oracle/__init__.py) — "parity unpinned".

Anchors in the reference (paths relative to /root/reference):
  * ctor kwargs / defaults ........ training.py:65-67, pattern_class.py:352-356,
                                    transfer_learning/transfer_learning.py:155-157
  * head = nn.Conv3d(f, out, 1) ... pattern_class.py:364
  * parameter-name prefixes ....... transfer_learning/transfer_learning.py:62-69
                                    ('final_conv', 'decoders.0|1|2')
  * outputs fed to CrossEntropyLoss / torch.max(out, 1) ... training.py:206-208
  * eval outputs used as per-class scores ................. pattern_class.py:275
North-star (BASELINE.json): in 1, out 56, 'crg', init 64, trilinear
interpolate, softmax head, state_dict loads into the reference unchanged.

Frozen upstream ambiguities:
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

# --- frozen choices ---------------------------------------------------------
UPSAMPLE_MODE = "trilinear"       # north_star says trilinear (binding)
UPSAMPLE_ALIGN_CORNERS = False
POOL_KERNEL = 2
POOL_STRIDE = 2
POOL_PADDING = 0                  # 96 -> 48 -> 24 -> 12
CONV_BIAS_WHEN_NORM = False       # Conv3d(bias=False) when 'g' or 'b' in order
GN_MAX_GROUPS = 32                # num_groups = min(init_channel_number // 2, 32)
GN_EPS = 1e-5
SOFTMAX_ONLY_IN_EVAL = True       # logits in train(), softmax in eval()


def _double_conv(in_ch, out_ch, encoder, order, num_groups):
    """(Conv3d 3x3x3 -> ReLU -> GroupNorm) x 2, names conv1/relu1/norm1/conv2/..."""
    if encoder:
        c1_in, c1_out = in_ch, max(out_ch // 2, in_ch)
        c2_in, c2_out = c1_out, out_ch
    else:
        c1_in, c1_out = in_ch, out_ch
        c2_in, c2_out = out_ch, out_ch
    seq = nn.Sequential()
    for pos, (ci, co) in ((1, (c1_in, c1_out)), (2, (c2_in, c2_out))):
        for i, ch in enumerate(order):
            if ch == "c":
                bias = not (("g" in order or "b" in order) and not CONV_BIAS_WHEN_NORM)
                seq.add_module("conv%d" % pos, nn.Conv3d(ci, co, 3, padding=1, bias=bias))
            elif ch == "r":
                seq.add_module("relu%d" % pos, nn.ReLU(inplace=False))
            elif ch == "g":
                nf = co if i > order.index("c") else ci
                seq.add_module("norm%d" % pos, nn.GroupNorm(num_groups, nf, eps=GN_EPS))
            else:
                raise ValueError("unsupported layer type %r (oracle covers 'c','r','g')" % ch)
    return seq


class _Encoder(nn.Module):
    def __init__(self, in_ch, out_ch, order, num_groups, is_max_pool):
        super().__init__()
        self.max_pool = (nn.MaxPool3d(POOL_KERNEL, POOL_STRIDE, POOL_PADDING)
                         if is_max_pool else None)
        self.double_conv = _double_conv(in_ch, out_ch, True, order, num_groups)

    def forward(self, x):
        if self.max_pool is not None:
            x = self.max_pool(x)
        return self.double_conv(x)


class _Decoder(nn.Module):
    def __init__(self, in_ch, out_ch, order, num_groups):
        super().__init__()
        self.double_conv = _double_conv(in_ch, out_ch, False, order, num_groups)

    def forward(self, skip, x):
        x = F.interpolate(x, size=skip.shape[2:], mode=UPSAMPLE_MODE,
                          align_corners=UPSAMPLE_ALIGN_CORNERS)
        x = torch.cat((skip, x), dim=1)          # skip channels first
        return self.double_conv(x)


def _rb(t):
    """bf16 storage emulation with a straight-through gradient."""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


def _double_conv_emulated(seq, x, first_layer_fp32):
    """Same arithmetic as the nn.Sequential, with the B200 path's bf16 storage points made explicit:
    3x3x3 weights (except the Cin=1 first conv), relu(conv) and the GroupNorm output are rounded to bf16."""
    for pos in (1, 2):
        conv = getattr(seq, "conv%d" % pos)
        norm = getattr(seq, "norm%d" % pos)
        w = conv.weight if (first_layer_fp32 and pos == 1) else _rb(conv.weight)
        x = F.conv3d(x, w, conv.bias, padding=1)
        x = _rb(F.relu(x))
        x = _rb(F.group_norm(x, norm.num_groups, norm.weight, norm.bias, norm.eps))
    return x


class UNet3DRef(nn.Module):
    """Oracle network.  state_dict keys (44 tensors for 'crg'):
    {encoders.{0-3},decoders.{0-2}}.double_conv.{conv,norm}{1,2}.*, final_conv.*

    ``emulate_bf16_storage = True`` keeps the fp32 arithmetic but rounds tensors to bf16 at exactly the points
    where the B200 path stores them (weights, relu(conv), GroupNorm output, upsample output).  Parity against
    this variant isolates kernel bugs from the precision budget of bf16 storage: ReLU masks then agree, so
    gradients can be compared tightly."""
    emulate_bf16_storage = False

    def __init__(self, in_channels, out_channels, final_sigmoid=False, interpolate=True,
                 dropout=0.0, conv_layer_order="crg", init_channel_number=64):
        super().__init__()
        if not interpolate:
            raise ValueError("oracle covers interpolate=True only (north_star)")
        if dropout not in (0, 0.0, None):
            raise ValueError("oracle covers dropout=0 only (training.py:66)")
        f = init_channel_number
        g = min(f // 2, GN_MAX_GROUPS)
        o = conv_layer_order
        self.encoders = nn.ModuleList([
            _Encoder(in_channels, f, o, g, False),
            _Encoder(f, 2 * f, o, g, True),
            _Encoder(2 * f, 4 * f, o, g, True),
            _Encoder(4 * f, 8 * f, o, g, True)])
        self.decoders = nn.ModuleList([
            _Decoder(4 * f + 8 * f, 4 * f, o, g),
            _Decoder(2 * f + 4 * f, 2 * f, o, g),
            _Decoder(f + 2 * f, f, o, g)])
        self.final_conv = nn.Conv3d(f, out_channels, 1)
        self.final_activation = nn.Sigmoid() if final_sigmoid else nn.Softmax(dim=1)

    def _forward_emulated(self, x):
        feats = []
        for i, enc in enumerate(self.encoders):
            if enc.max_pool is not None:
                x = enc.max_pool(x)
            x = _double_conv_emulated(enc.double_conv, x, first_layer_fp32=(i == 0))
            feats.insert(0, x)
        for dec, skip in zip(self.decoders, feats[1:]):
            x = _rb(F.interpolate(x, size=skip.shape[2:], mode=UPSAMPLE_MODE,
                                  align_corners=UPSAMPLE_ALIGN_CORNERS))
            x = _double_conv_emulated(dec.double_conv, torch.cat((skip, x), dim=1), False)
        return x

    def forward(self, x):
        if self.emulate_bf16_storage:
            x = self._forward_emulated(x)
        else:
            feats = []
            for enc in self.encoders:
                x = enc(x)
                feats.insert(0, x)
            for dec, skip in zip(self.decoders, feats[1:]):
                x = dec(skip, x)
        x = self.final_conv(x)                    # read lazily: callers replace it
        if SOFTMAX_ONLY_IN_EVAL and not self.training:
            x = self.final_activation(x)
        return x
