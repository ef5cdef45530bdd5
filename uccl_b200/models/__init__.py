from .moe import ExpertParallelMoE  # noqa: F401
from .resnet import ResNet, resnet18, resnet50  # noqa: F401
