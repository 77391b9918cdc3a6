from .bert_dataset import GlueDataset
from .data_generator import BaseGenerator, DataloaderGenerator, RandomTensorGenerator
from .dataset import CIFAR10Dataset, RandomImageDataset, RandomMlpDataset, SynthMNLIDataset

__all__ = ["GlueDataset", "BaseGenerator", "DataloaderGenerator", "RandomTensorGenerator",
           "CIFAR10Dataset", "RandomImageDataset", "RandomMlpDataset", "SynthMNLIDataset"]
