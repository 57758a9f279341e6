"""BertConfig: the config object the reference worker builds (worker.py:495, 506-522).

Same surface as the [UPSTREAM] ``vilbert.vilbert.BertConfig`` the worker imports (worker.py:45):
``from_json_file`` / ``from_dict`` / attribute access / ``to_dict`` / ``to_json_string``; the worker then
mutates ``v_target_size``, ``predict_feature``, ``task_specific_tokens``, ``dynamic_attention`` and
``visualization`` in place (worker.py:509-522).  Defaults are the values of
``config/bert_base_6layer_6conect.json`` (file name pinned by worker.py:472).
"""
from __future__ import annotations

import copy
import json

_DEFAULTS = {
    "vocab_size": 30522, "hidden_size": 768, "num_hidden_layers": 12, "num_attention_heads": 12,
    "intermediate_size": 3072, "hidden_act": "gelu", "hidden_dropout_prob": 0.1,
    "attention_probs_dropout_prob": 0.1, "max_position_embeddings": 512, "type_vocab_size": 2,
    "initializer_range": 0.02, "v_feature_size": 2048, "v_target_size": 1601, "v_hidden_size": 1024,
    "v_num_hidden_layers": 6, "v_num_attention_heads": 8, "v_intermediate_size": 1024,
    "bi_hidden_size": 1024, "bi_num_attention_heads": 8, "bi_intermediate_size": 1024,
    "bi_attention_type": 1, "v_attention_probs_dropout_prob": 0.1, "v_hidden_act": "gelu",
    "v_hidden_dropout_prob": 0.1, "v_initializer_range": 0.02,
    "v_biattention_id": [0, 1, 2, 3, 4, 5], "t_biattention_id": [6, 7, 8, 9, 10, 11],
    "pooling_method": "mul", "fusion_method": "mul", "predict_feature": False, "fast_mode": False,
    "fixed_v_layer": 0, "fixed_t_layer": 0, "in_batch_pairs": False, "fusion_method_": None,
    "dynamic_attention": False, "with_coattention": True, "objective": 0, "num_negative": 128,
    "model": "bert", "task_specific_tokens": False, "visualization": False, "num_task_tokens": 20,
}


class BertConfig(object):
    def __init__(self, vocab_size_or_config_json_file=None, **kwargs):
        d = copy.deepcopy(_DEFAULTS)
        d.pop("fusion_method_")
        if isinstance(vocab_size_or_config_json_file, str):
            with open(vocab_size_or_config_json_file, "r", encoding="utf-8") as reader:
                d.update(json.loads(reader.read()))
        elif isinstance(vocab_size_or_config_json_file, int):
            d["vocab_size"] = vocab_size_or_config_json_file
        elif vocab_size_or_config_json_file is not None:
            raise ValueError("First argument must be either a vocabulary size (int) "
                             "or the path to a pretrained model config file (str)")
        d.update(kwargs)
        self.__dict__.update(d)

    @classmethod
    def from_dict(cls, json_object):
        config = cls()
        config.__dict__.update(copy.deepcopy(dict(json_object)))
        return config

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r", encoding="utf-8") as reader:
            return cls.from_dict(json.loads(reader.read()))

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def __repr__(self):
        return str(self.to_json_string())
