"""Model zoo: the reference's ConvNet (flagship) and ResNet-18/34 (bucket/overlap stress)."""
from .convnet import ConvNet
from .resnet import BasicBlock, ResNet, resnet18, resnet34

__all__ = ["ConvNet", "ResNet", "BasicBlock", "resnet18", "resnet34"]
