"""Sequential CNNs with a Keras-compatible layer/weight naming contract.

``MedCNN`` is the reference model (FLPyfhelin.py:118-136): six [Conv 3x3 valid + ReLU ->
MaxPool 2x2] stages with 32,32,32,64,64,128 filters, Flatten, Dense 128, Dense 64, Dense 2
(softmax folded into the loss). 222,722 parameters in 18 tensors.

``keras_layers()`` reproduces Keras' ``model.layers`` indexing (conv at 0,2,..,10, pools in
between, Flatten at 12, Dense at 13-15) so ciphertext dictionaries use the reference's keys
``c_{layer}_{0|1}`` (FLPyfhelin.py:221) and Keras layouts (conv HWIO, dense [in,out],
Flatten in H,W,C order).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class _SeqCNN(nn.Module):
    conv_channels: Tuple[int, ...] = ()
    dense_units: Tuple[int, ...] = ()

    def __init__(self, in_channels: int, num_classes: int, image_size: int):
        super().__init__()
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.image_size = image_size
        chans = [in_channels, *self.conv_channels]
        self.convs = nn.ModuleList(nn.Conv2d(chans[i], chans[i + 1], 3) for i in range(len(self.conv_channels)))
        s = image_size
        for _ in self.conv_channels:
            s = (s - 2) // 2
        if s < 1:
            raise ValueError(f"image size {image_size} too small for {len(self.conv_channels)} conv stages")
        self.final_hw = s
        units = [s * s * self.conv_channels[-1], *self.dense_units, num_classes]
        self.fcs = nn.ModuleList(nn.Linear(units[i], units[i + 1]) for i in range(len(units) - 1))
        self.reset_parameters_keras()

    def reset_parameters_keras(self) -> None:
        """Keras defaults: glorot_uniform kernels, zero biases."""
        for m in [*self.convs, *self.fcs]:
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)

    def features(self, x: torch.Tensor) -> torch.Tensor:
        for conv in self.convs:
            x = F.max_pool2d(F.relu(conv(x)), 2)
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.features(x)
        x = x.permute(0, 2, 3, 1).flatten(1)  # Keras Flatten: H, W, C order
        for fc in self.fcs[:-1]:
            x = F.relu(fc(x))
        return self.fcs[-1](x)  # logits; softmax lives in the loss (FLPyfhelin.py:136,141)

    # ---- Keras naming contract ---------------------------------------------------------
    def keras_layers(self) -> List[Tuple[str, Optional[nn.Module]]]:
        layers: List[Tuple[str, Optional[nn.Module]]] = []
        for conv in self.convs:
            layers.append(("conv", conv))
            layers.append(("pool", None))
        layers.append(("flatten", None))
        for fc in self.fcs:
            layers.append(("dense", fc))
        return layers


class MedCNN(_SeqCNN):
    conv_channels = (32, 32, 32, 64, 64, 128)
    dense_units = (128, 64)

    def __init__(self, in_channels: int = 3, num_classes: int = 2, image_size: int = 256):
        super().__init__(in_channels, num_classes, image_size)


class SmallCNN(_SeqCNN):
    """2-conv CNN for the 28x28 CPU plumbing config (BASELINE.json configs[0])."""

    conv_channels = (8, 16)
    dense_units = ()

    def __init__(self, in_channels: int = 1, num_classes: int = 10, image_size: int = 28):
        super().__init__(in_channels, num_classes, image_size)
