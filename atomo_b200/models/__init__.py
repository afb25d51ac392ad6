"""Model zoo + the ``--network`` name resolver.

Parity: the ``build_model`` dispatch in
``/root/reference/src/sync_replicas_master_nn.py:146-171`` /
``distributed_worker.py:139-164`` (``LeNet, ResNet18, ResNet34, FC, DenseNet,
VGG11, AlexNet``), widened with every depth/variant the model files define.
"""
from .lenet import LeNet
from .fc_nn import FC_NN
from .resnet import ResNet, ResNet18, ResNet34, ResNet50, ResNet101, ResNet152, BasicBlock, Bottleneck
from .vgg import VGG, vgg11, vgg11_bn, vgg13, vgg13_bn, vgg16, vgg16_bn, vgg19, vgg19_bn
from .alexnet import AlexNet, alexnet
from .densenet import DenseNet
from .split import SplitModel, LeNetSplit, FC_NN_Split, ResNetSplit18

NETWORKS = ("LeNet", "FC", "ResNet18", "ResNet34", "ResNet50", "ResNet101", "ResNet152",
            "VGG11", "VGG13", "VGG16", "VGG19", "AlexNet", "DenseNet", "DenseNetSmall")


def build_model(network: str, num_classes: int = 10, dataset: str = ""):
    """Construct the network named by ``--network``.

    ``dataset == 'ImageNet'`` selects the ImageNet stem for ResNets.
    """
    imagenet = dataset.lower() == "imagenet"
    if network == "LeNet":
        return LeNet(num_classes)
    if network == "FC":
        return FC_NN(num_classes)
    if network.startswith("ResNet"):
        ctor = {"ResNet18": ResNet18, "ResNet34": ResNet34, "ResNet50": ResNet50,
                "ResNet101": ResNet101, "ResNet152": ResNet152}[network]
        return ctor(num_classes=num_classes, imagenet_stem=imagenet)
    if network == "VGG11":
        return vgg11_bn(num_classes)
    if network == "VGG13":
        return vgg13_bn(num_classes)
    if network == "VGG16":
        return vgg16_bn(num_classes)
    if network == "VGG19":
        return vgg19_bn(num_classes)
    if network == "AlexNet":
        return alexnet(num_classes=num_classes)
    if network == "DenseNet":  # DenseNet-BC-190-40, master:156-158
        return DenseNet(growthRate=40, depth=190, reduction=0.5, bottleneck=True, nClasses=num_classes)
    if network == "DenseNetSmall":  # DenseNet-BC-100-12, for tests / small GPUs
        return DenseNet(growthRate=12, depth=100, reduction=0.5, bottleneck=True, nClasses=num_classes)
    raise ValueError("unknown --network %r (choose from %s)" % (network, ", ".join(NETWORKS)))


def input_shape(network: str, dataset: str = ""):
    """(C, H, W) the network expects for a dataset name."""
    d = dataset.lower()
    if network in ("LeNet", "FC") or d == "mnist":
        return (1, 28, 28)
    if network == "AlexNet" or d == "imagenet":
        return (3, 224, 224) if network != "AlexNet" else (3, 227, 227)
    return (3, 32, 32)
