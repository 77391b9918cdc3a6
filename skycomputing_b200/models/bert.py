"""BertConfig (capability parity with scaelum/model/bert.py:6-99).

A plain attribute container that can be built from a vocab size, a dict or a Google-style
``bert_config.json``; layers receive ``config.__dict__`` in the layer-list config and rebuild it
with ``BertConfig.from_dict`` (reference convention, e.g. scaelum/model/bert_layers.py:177).
"""
from __future__ import annotations

import copy
import json
from typing import Any, Dict, Union


class BertConfig(dict):
    def __init__(
        self,
        vocab_size_or_config_json_file: Union[int, str] = 30522,
        hidden_size: int = 768,
        num_hidden_layers: int = 12,
        num_attention_heads: int = 12,
        intermediate_size: int = 3072,
        hidden_act: str = "gelu",
        hidden_dropout_prob: float = 0.1,
        attention_probs_dropout_prob: float = 0.1,
        max_position_embeddings: int = 512,
        type_vocab_size: int = 2,
        initializer_range: float = 0.02,
        output_all_encoded_layers: bool = False,
    ):
        super().__init__()
        if isinstance(vocab_size_or_config_json_file, str):
            with open(vocab_size_or_config_json_file, "r", encoding="utf-8") as reader:
                json_config = json.loads(reader.read())
            for key, value in json_config.items():
                self.__dict__[key] = value
        elif isinstance(vocab_size_or_config_json_file, int):
            self.vocab_size = vocab_size_or_config_json_file
            self.hidden_size = hidden_size
            self.num_hidden_layers = num_hidden_layers
            self.num_attention_heads = num_attention_heads
            self.hidden_act = hidden_act
            self.intermediate_size = intermediate_size
            self.hidden_dropout_prob = hidden_dropout_prob
            self.attention_probs_dropout_prob = attention_probs_dropout_prob
            self.max_position_embeddings = max_position_embeddings
            self.type_vocab_size = type_vocab_size
            self.initializer_range = initializer_range
            self.output_all_encoded_layers = output_all_encoded_layers
        else:
            raise ValueError(
                "First argument must be either a vocabulary size (int) "
                "or the path to a pretrained model config file (str)"
            )

    # BERT-large geometry used by the reference experiment
    # (wwm_uncased_L-24_H-1024_A-16/bert_config.json, experiment/config.py:22-24)
    @classmethod
    def bert_large(cls, **overrides: Any) -> "BertConfig":
        cfg = cls(30522, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                  intermediate_size=4096)
        for k, v in overrides.items():
            setattr(cfg, k, v)
        return cfg

    @classmethod
    def from_dict(cls, json_object: Dict[str, Any]) -> "BertConfig":
        config = BertConfig(vocab_size_or_config_json_file=-1)
        for key, value in json_object.items():
            config.__dict__[key] = value
        return config

    @classmethod
    def from_json_file(cls, json_file: str) -> "BertConfig":
        with open(json_file, "r", encoding="utf-8") as reader:
            return cls.from_dict(json.loads(reader.read()))

    def __repr__(self) -> str:
        return str(self.to_json_string())

    def to_dict(self) -> Dict[str, Any]:
        return copy.deepcopy(self.__dict__)

    def to_json_string(self) -> str:
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"
