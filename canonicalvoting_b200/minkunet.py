"""MinkUNet34C -- the sparse-voxel U-Net of the reference (utils/minkunet.py:36-245 + utils/resnet.py:109-154),
built on canonicalvoting_b200.sparse.

Same module API and the same attribute names as the reference model, hence the same state-dict keys and
shapes (188 tensors, 37 860 320 parameters for MinkUNet34C(3, 64); `conv0p1s1.kernel [125,3,32]`,
`block5.0.conv1.kernel [27,384,256]`, `final.bias [1,64]`): a checkpoint of the reference loads with
`load_state_dict`.  The reference's own `utils/minkunet.py` also runs unchanged on the `MinkowskiEngine/`
compat package of this repository; this file exists because /root/reference is not present where the
benchmark runs.  The network is described by a stage table instead of a hand-written layer list.
"""
import torch.nn as nn

from . import sparse as ME

# (conv name, bn name, block name) per encoder stage / decoder stage, in forward order (utils/minkunet.py:122-180)
_ENCODER = [("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"), ("conv3p4s2", "bn3", "block3"),
            ("conv4p8s2", "bn4", "block4")]
_DECODER = [("convtr4p16s2", "bntr4", "block5"), ("convtr5p8s2", "bntr5", "block6"), ("convtr6p4s2", "bntr6", "block7"),
            ("convtr7p2s2", "bntr7", "block8")]


class MinkUNetBase(nn.Module):
    BLOCK = ME.BasicBlock
    PLANES = None
    LAYERS = (2, 2, 2, 2, 2, 2, 2, 2)
    INIT_DIM = 32
    OUT_TENSOR_STRIDE = 1

    def __init__(self, in_channels, out_channels, D=3):
        super().__init__()
        self.D = D
        P, Ls, ex = self.PLANES, self.LAYERS, self.BLOCK.expansion
        self.inplanes = self.INIT_DIM
        self.conv0p1s1 = ME.MinkowskiConvolution(in_channels, self.inplanes, kernel_size=5, dimension=D)   # :53
        self.bn0 = ME.MinkowskiBatchNorm(self.inplanes)
        for i, (conv, bn, block) in enumerate(_ENCODER):                                                   # :58-83
            setattr(self, conv, ME.MinkowskiConvolution(self.inplanes, self.inplanes, kernel_size=2, stride=2, dimension=D))
            setattr(self, bn, ME.MinkowskiBatchNorm(self.inplanes))
            setattr(self, block, self._make_layer(self.BLOCK, P[i], Ls[i]))
        skips = [P[2] * ex, P[1] * ex, P[0] * ex, self.INIT_DIM]                                           # :89,96,103,110
        for i, (conv, bn, block) in enumerate(_DECODER):                                                   # :85-112
            setattr(self, conv, ME.MinkowskiConvolutionTranspose(self.inplanes, P[4 + i], kernel_size=2, stride=2, dimension=D))
            setattr(self, bn, ME.MinkowskiBatchNorm(P[4 + i]))
            self.inplanes = P[4 + i] + skips[i]
            setattr(self, block, self._make_layer(self.BLOCK, P[4 + i], Ls[4 + i]))
        self.final = ME.MinkowskiConvolution(P[7], out_channels, kernel_size=1, bias=True, dimension=D)    # :114
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.weight_initialization()

    def weight_initialization(self):
        """utils/resnet.py:109-116: kaiming-normal (fan_out, relu) on MinkowskiConvolution kernels only (the
        isinstance test skips the transposed convolutions), BN weight 1 / bias 0."""
        for m in self.modules():
            if isinstance(m, ME.MinkowskiConvolution):
                ME.utils.kaiming_normal_(m.kernel, mode="fan_out", nonlinearity="relu")
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1, bn_momentum=0.1):
        """utils/resnet.py:118-154."""
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                ME.MinkowskiConvolution(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, dimension=self.D),
                ME.MinkowskiBatchNorm(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride=stride, dilation=dilation, downsample=downsample, dimension=self.D)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes, stride=1, dilation=dilation, dimension=self.D) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        """utils/minkunet.py:122-180."""
        out = self.relu(self.bn0(self.conv0p1s1(x)))
        skips = [out]
        for conv, bn, block in _ENCODER:
            out = getattr(self, block)(self.relu(getattr(self, bn)(getattr(self, conv)(out))))
            skips.append(out)
        skips.pop()                                   # the stride-16 tensor is not a skip
        for conv, bn, block in _DECODER:
            out = self.relu(getattr(self, bn)(getattr(self, conv)(out)))
            out = getattr(self, block)(ME.cat(out, skips.pop()))
        return self.final(out)


# The model family of utils/minkunet.py:183-245 as a table: depth (blocks per stage) x width variant (planes per stage).
# Only MinkUNet34C is instantiated by the reference's scripts; MinkUNet14A is the small net of the tests and smoke().
_DEPTHS = {"14": (1, 1, 1, 1, 1, 1, 1, 1), "18": (2, 2, 2, 2, 2, 2, 2, 2), "34": (2, 3, 4, 6, 2, 2, 2, 2)}
_ENC = (32, 64, 128, 256)
_WIDTHS = {"14A": (128, 128, 96, 96), "14B": (128, 128, 128, 128), "14C": (192, 192, 128, 128), "14D": (384, 384, 384, 384),
           "18A": (128, 128, 96, 96), "18B": (128, 128, 128, 128), "18D": (384, 384, 384, 384),
           "34A": (256, 128, 64, 64), "34B": (256, 128, 64, 32), "34C": (256, 128, 96, 96)}
for _d, _layers in _DEPTHS.items():
    globals()["MinkUNet" + _d] = type("MinkUNet" + _d, (MinkUNetBase,), {"LAYERS": _layers, "__module__": __name__})
for _v, _dec in _WIDTHS.items():
    globals()["MinkUNet" + _v] = type("MinkUNet" + _v, (globals()["MinkUNet" + _v[:2]],), {"PLANES": _ENC + _dec, "__module__": __name__})


def decode_heads(feats, nclasses=9, log_scale=True):
    """Head decode of the joint model (eval_joint.py:173-190) in plain torch ops: [N, 6*C + C + 1] ->
    xyz_pred [N,3], scale_pred [N,3], class_pred [N] int64, prob_pred [N]."""
    import torch
    xyz = feats[:, :3 * nclasses].reshape(-1, nclasses, 3)
    scale = feats[:, 3 * nclasses:6 * nclasses].reshape(-1, nclasses, 3)
    cls = feats[:, 6 * nclasses:]
    k = cls.argmax(-1)
    k = torch.where(k == nclasses, torch.zeros_like(k), k)
    idx = k.view(-1, 1, 1).expand(-1, 1, 3)
    xyz_pred = torch.gather(xyz, 1, idx)[:, 0].contiguous()
    scale_pred = torch.gather(scale, 1, idx)[:, 0]
    scale_pred = (torch.exp(scale_pred) if log_scale else scale_pred).contiguous()
    class_pred = torch.argmax(cls[:, :-1], -1)
    prob_pred = torch.softmax(cls, -1)[:, :-1].max(-1)[0].contiguous()
    return xyz_pred, scale_pred, class_pred, prob_pred
