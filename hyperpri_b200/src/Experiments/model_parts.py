"""U-Net building blocks with the reference's names, constructor signatures and state-dict
layout (reference src/Experiments/model_parts.py:14-99), re-hosted on the B200 kernels.

The blocks are parameter containers: the arithmetic of a whole network is scheduled by
``hyperpri_b200.engine`` (conv+BN-statistics GEMM epilogues, fused BN/ReLU/pool passes, concat by
placement), so a block does not run layer-by-layer torch ops.  `use_attention=True` (skip * up,
model_parts.py:84-85) and `bilinear=True` (nn.Upsample + mid-channel DoubleConvs, :56-61) are built,
although every reference config leaves them off (params_HyperPRI.py:53-54,210-211).
"""
import torch.nn as nn

_STANDALONE = ("{} is scheduled as part of UNet/CubeNET by hyperpri_b200.engine; "
               "standalone block forward is not on the B200 hot path")


def _conv_bn_relu(cin, cout):
    return [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]


class DoubleConv(nn.Module):
    """[conv3x3 -> BatchNorm -> ReLU] twice; keys double_conv.{0,1,3,4}.* (model_parts.py:14-31)."""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid = mid_channels or out_channels
        self.double_conv = nn.Sequential(*_conv_bn_relu(in_channels, mid), *_conv_bn_relu(mid, out_channels))

    def forward(self, x):
        raise NotImplementedError(_STANDALONE.format("DoubleConv"))


class Down(nn.Module):
    """MaxPool2d(2) then DoubleConv; keys maxpool_conv.1.double_conv.* (model_parts.py:34-45)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        raise NotImplementedError(_STANDALONE.format("Down"))


class Up(nn.Module):
    """ConvTranspose2d(k2,s2) -> pad to the skip -> cat([skip, up]) (use_attention: skip * up, a DoubleConv over
    in_channels // 2) -> DoubleConv (model_parts.py:48-90); keys up.{weight,bias}, conv.double_conv.*."""

    def __init__(self, in_channels, out_channels, bilinear=True, use_attention=False):
        super().__init__()
        self.use_attention = use_attention
        cin = in_channels // 2 if use_attention else in_channels
        if bilinear:                                                     # model_parts.py:56-61
            self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
            self.conv = DoubleConv(cin, out_channels // 2, in_channels // 2)
        else:                                                            # model_parts.py:62-68
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.conv = DoubleConv(cin, out_channels)

    def forward(self, x1, x2):
        raise NotImplementedError(_STANDALONE.format("Up"))


class OutConv(nn.Module):
    """1x1 conv head; keys conv.{weight,bias} (model_parts.py:93-99)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)

    def forward(self, x):
        raise NotImplementedError(_STANDALONE.format("OutConv"))
