"""Model zoo: the reference's 6-conv medical CNN (FLPyfhelin.py:118-136), the 2-layer CNN of
the CPU plumbing config, and ResNet-18/50 (BASELINE.json configs)."""
from .cnn import MedCNN, SmallCNN  # noqa: F401
from .pack import ParamPack  # noqa: F401
from .resnet import resnet18, resnet50  # noqa: F401


def create_model(name: str, in_channels: int = 3, num_classes: int = 2, image_size: int = 256):
    name = name.lower()
    if name == "medcnn":
        return MedCNN(in_channels, num_classes, image_size)
    if name == "cnn2":
        return SmallCNN(in_channels, num_classes, image_size)
    if name == "resnet18":
        return resnet18(num_classes=num_classes, in_channels=in_channels)
    if name == "resnet50":
        return resnet50(num_classes=num_classes, in_channels=in_channels)
    raise ValueError(f"unknown model {name}")
