from .processor import (PROCESSORS, ColaProcessor, InputExample, InputFeatures, MnliProcessor,
                        MrpcProcessor, Sst2Processor, convert_examples_to_features)
from .tokenization import BasicTokenizer, BertTokenizer, WordpieceTokenizer, load_vocab

__all__ = ["PROCESSORS", "ColaProcessor", "MnliProcessor", "MrpcProcessor", "Sst2Processor",
           "InputExample", "InputFeatures", "convert_examples_to_features", "BertTokenizer",
           "BasicTokenizer", "WordpieceTokenizer", "load_vocab"]
