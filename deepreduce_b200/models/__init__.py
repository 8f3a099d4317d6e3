from .resnet import ResNet, ResNetCifar, resnet20, resnet50
from .zoo import VGG16, DenseNet40, MobileNet, NeuMF, NextWordLSTM, bert_large, count_params

__all__ = ["ResNet", "ResNetCifar", "resnet20", "resnet50", "VGG16", "DenseNet40", "MobileNet", "NeuMF",
           "NextWordLSTM", "bert_large", "count_params"]
