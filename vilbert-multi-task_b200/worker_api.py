"""Task / prediction API of the reference worker, minus Django, RabbitMQ and the detector.

Same call signatures, same validation, same per-task output dictionaries as
/root/reference/worker.py:

    custom_prediction(query, task, features, infos, task_id)                      worker.py:388-458
    prediction(question, features, spatials, segment_ids, input_mask, image_mask,
               co_attention_mask, task_tokens, task_id, infos)                    worker.py:248-386
    shape_result(task_id, answer, image_path)   (the `result` dicts of callback)  worker.py:564-645

Differences that do not change results: the model call goes to the sm_100a engine; only the heads a task
reads are computed (the reference computes all ten outputs and drops nine); top-N uses `topk` instead of a
full sort + per-element `.item()`; label maps are loaded once instead of per request (worker.py:299-300,
311-315); `eval(task_id)` (worker.py:562) is `int(task_id)`.
"""
from __future__ import annotations

import os
import pickle
import re
import unicodedata
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L

# ------------------------------------------------------------------------------------------ module state
# (the reference keeps `model`, `tokenizer`, `feature_extractor` as module globals, worker.py:464-466)
model = None
tokenizer = None
label_maps: Dict[str, Optional[list]] = {"vqa": None, "gqa": None}

MAX_LENGTH = 37                     # worker.py:408
SINGLE_IMAGE_TASKS = ("1", "15", "13", "11", "4", "16")   # worker.py:256
PAIR_TASKS = ("12",)                # worker.py:258
RETRIEVAL_TASKS = ("7",)            # worker.py:260
GROUNDING_TASKS = ("11", "4", "16")  # worker.py:374
# task id -> which element of the 10-tuple the decode reads (worker.py:295-374)
TASK_OUTPUT = {"1": L.OUT_VIL_PREDICTION, "2": L.OUT_VIL_PREDICTION, "15": L.OUT_VIL_PREDICTION_GQA,
               "12": L.OUT_VIL_BINARY_PREDICTION, "13": L.OUT_VIL_TRI_PREDICTION, "7": L.OUT_VIL_LOGIT,
               "11": L.OUT_VISION_LOGIT, "4": L.OUT_VISION_LOGIT, "16": L.OUT_VISION_LOGIT}
# task-id -> name table of the demo page (demo/templates/vilbert_multitask/result.html:321-335)
TASK_NAMES = {"1": "VQA", "2": "VG-QA", "15": "GQA", "12": "NLVR2", "13": "SNLI-VE", "7": "Image Retrieval",
              "11": "RefCOCO", "4": "Visual7W", "16": "GuessWhat"}


def load_vilbert_model(from_pretrained, config_file, num_labels: int = 3129, device: int = 0, vocab_file=None,
                       label2ans: Optional[Dict[str, str]] = None, tokenizer_obj=None, **engine_kw):
    """worker.py:463-539 without the detector: config mutation (worker.py:509-522), from_pretrained, eval, cuda."""
    global model, tokenizer
    from .config import BertConfig
    from .model import VILBertForVLTasks
    config = config_file if isinstance(config_file, BertConfig) else BertConfig.from_json_file(config_file)
    config.v_target_size = 1601            # predict_feature=False, worker.py:512-514
    config.predict_feature = False
    config.task_specific_tokens = True     # worker.py:516-517
    config.visualization = True            # worker.py:522
    m = VILBertForVLTasks.from_pretrained(from_pretrained, config=config, num_labels=num_labels, default_gpu=True,
                                          **engine_kw)
    m.eval()
    m = m.cuda(device)
    model = m
    tokenizer = tokenizer_obj if tokenizer_obj is not None else (WordpieceTokenizer(vocab_file) if vocab_file else None)
    for key, path in (label2ans or {}).items():
        with open(path, "rb") as f:
            label_maps[key] = pickle.load(f)
    return m


def _label(kind: str, idx: int):
    m = label_maps.get(kind)
    if m is None:       # no trainval_label2ans.pkl on this box (worker.py:299, 311 read them from save/...)
        return f"<{kind}:{idx}>"
    return m[idx]


def _top(prob_1d: torch.Tensor, n: int):
    val, idx = torch.topk(prob_1d, n)      # == sort descending + first n (worker.py:297, 314, 330, 346, 360, 376)
    return val.tolist(), idx.tolist()


def prediction(question, features, spatials, segment_ids, input_mask, image_mask, co_attention_mask, task_tokens,
               task_id, infos):
    """Validate, expand the text for pair / retrieval tasks, run the model once, decode.  worker.py:248-386."""
    N = len(infos) if task_id == "7" else 3                      # worker.py:250-253
    if task_id in SINGLE_IMAGE_TASKS:
        assert len(infos) == 1, "task require 1 image"
    elif task_id in PAIR_TASKS:
        assert len(infos) == 2, "task require 2 images"
    elif task_id in RETRIEVAL_TASKS:
        assert len(infos) > 1 and len(infos) <= 10, "task require 2-10 images"
    else:
        raise ValueError("task not valid.")

    rep = 2 if task_id == "12" else (features.size(0) if task_id == "7" else 1)     # worker.py:266-284
    if rep > 1:
        question = question.repeat(rep, 1)
        input_mask = input_mask.repeat(rep, 1)
        segment_ids = segment_ids.repeat(rep, 1)
        task_tokens = task_tokens.repeat(rep, 1)

    out = model(question, features, spatials, segment_ids, input_mask, image_mask, co_attention_mask, task_tokens,
                output_all_attention_masks=True, select=TASK_OUTPUT[task_id])
    (vil_prediction, vil_prediction_gqa, vil_logit, vil_binary_prediction, vil_tri_prediction, _vision_prediction,
     vision_logit, _linguisic_prediction, _linguisic_logit, _attn) = out

    if task_id in ("1", "2"):
        conf, idx = _top(torch.softmax(vil_prediction.view(-1), dim=0), N)
        return {"top3_answer": [_label("vqa", i) for i in idx], "top3_confidence": conf}
    if task_id == "15":
        conf, idx = _top(torch.softmax(vil_prediction_gqa.view(-1), dim=0), N)
        return {"top3_answer": [_label("gqa", i) for i in idx], "top3_confidence": conf}
    if task_id == "12":
        names = {0: "False", 1: "True"}
        conf, idx = _top(torch.softmax(vil_binary_prediction.view(-1), dim=0), 2)
        return {"top3_answer": [names[i] for i in idx], "top3_confidence": conf}
    if task_id == "13":
        names = {0: "contradiction (false)", 1: "neutral", 2: "entailment (true)"}
        conf, idx = _top(torch.softmax(vil_tri_prediction.view(-1), dim=0), 3)
        return {"top3_answer": [names[i] for i in idx], "top3_confidence": conf}
    if task_id == "7":
        conf, idx = _top(torch.softmax(vil_logit.view(-1), dim=0), N)
        return {"top3_answer": idx, "top3_confidence": conf}
    # grounding: softmax over ALL rows of vision_logit, the synthetic global box included (worker.py:374-385)
    image_w, image_h = infos[0]["image_width"], infos[0]["image_height"]
    conf, idx = _top(torch.softmax(vision_logit.view(-1), dim=0), N)
    boxes = spatials[0][torch.as_tensor(idx, device=spatials.device)][:, :4].tolist()
    return [{"y1": int(b[1] * image_h), "y2": int(b[3] * image_h), "x1": int(b[0] * image_w), "x2": int(b[2] * image_w),
             "confidence": c * 100} for b, c in zip(boxes, conf)]


def build_inputs(query, task, features, infos, device, tok=None):
    """The tensor construction half of custom_prediction (worker.py:402-455), returned instead of consumed."""
    tok = tok if tok is not None else tokenizer
    if tok is None:
        raise L.VilbertB200Error("no tokenizer: pass vocab_file= (bert-base-uncased vocab.txt) to load_vilbert_model")
    tokens = tok.encode(query)
    tokens = tok.add_special_tokens_single_sentence(tokens)
    segment_ids = [0] * len(tokens)
    input_mask = [1] * len(tokens)
    if len(tokens) < MAX_LENGTH:                       # pad, never truncate (worker.py:408-414)
        pad = [0] * (MAX_LENGTH - len(tokens))
        tokens, input_mask, segment_ids = tokens + pad, input_mask + pad, segment_ids + pad
    text = torch.tensor(tokens, dtype=torch.long, device=device).unsqueeze(0)
    input_mask = torch.tensor(input_mask, dtype=torch.long, device=device).unsqueeze(0)
    segment_ids = torch.tensor(segment_ids, dtype=torch.long, device=device).unsqueeze(0)
    task_t = torch.tensor(np.array(task), dtype=torch.long, device=device).unsqueeze(0)

    feats, locs, masks = [], [], []
    for feature, info in zip(features, infos):
        w, h = float(info["image_width"]), float(info["image_height"])
        feature = torch.as_tensor(feature).to(device)
        n = feature.shape[0]
        g_feat = feature.sum(dim=0) / n                                   # mean-pooled global row (worker.py:432-434)
        feats.append(torch.cat([g_feat.view(1, -1), feature], dim=0))
        boxes = np.asarray(info["bbox"], dtype=np.float32)
        loc = np.zeros((boxes.shape[0], 5), dtype=np.float32)
        loc[:, :4] = boxes
        loc[:, 4] = (loc[:, 3] - loc[:, 1]) * (loc[:, 2] - loc[:, 0]) / (w * h)
        loc[:, [0, 2]] /= w
        loc[:, [1, 3]] /= h
        loc = np.concatenate([np.array([[0, 0, 1, 1, 1]], dtype=np.float32), loc], axis=0)   # worker.py:443-444
        locs.append(torch.from_numpy(loc))
        masks.append(torch.ones(n + 1, dtype=torch.uint8))
    features_t = torch.stack(feats, dim=0).float().to(device)              # requires equal n across images
    spatials = torch.stack(locs, dim=0).float().to(device)
    image_mask = torch.stack(masks, dim=0).to(device)
    co_attention_mask = torch.zeros((len(infos), features_t.shape[1], text.shape[1]), device=device)
    return text, features_t, spatials, segment_ids, input_mask, image_mask, co_attention_mask, task_t


def custom_prediction(query, task, features, infos, task_id):
    """worker.py:388-458 (the GuessWhat dialog rewrite at 391-400 builds `tokens` and then discards it)."""
    device = torch.device("cuda", model._device)
    text, feats, spatials, segment_ids, input_mask, image_mask, co_mask, task_t = build_inputs(query, task, features,
                                                                                              infos, device)
    return prediction(text, feats, spatials, segment_ids, input_mask, image_mask, co_mask, task_t, task_id, infos)


def shape_result(task_id: str, answer, image_path: Sequence[str], image_names: Optional[List[str]] = None):
    """The JSON-able `result` dict callback() pushes to the WebSocket, per task family (worker.py:564-645).
    Box drawing / file writing (cv2, worker.py:596-600) is left to the caller; pass the names it chose."""
    if task_id in ("1", "15", "2", "13"):
        return {"task_id": task_id, "result": [{"answer": answer["top3_answer"][i],
                                                "confidence": round(answer["top3_confidence"][i] * 100, 2)}
                                               for i in range(3)]}
    if task_id in ("4", "16", "11"):
        names = image_names if image_names is not None else [str(i) for i in range(len(answer))]
        return {"task_id": task_id, "image_name_list": names[:3],
                "confidence_list": [round(a["confidence"], 2) for a in answer[:3]]}
    if task_id == "12":
        return {"task_id": task_id, "result": [{"answer": answer["top3_answer"][i],
                                                "confidence": round(answer["top3_confidence"][i] * 100, 2)}
                                               for i in range(2)]}
    if task_id == "7":
        prefix = "demo/" if "demo" in image_path[0].split("/") else "test2014/"
        ext = str(image_path[0].split("/")[-1].split(".")[1])
        names = [prefix + os.path.split(image_path[i])[1].split(".")[0] + "." + ext for i in answer["top3_answer"]]
        return {"task_id": task_id, "image_name_list": names,
                "confidence_list": [round(c * 100, 2) for c in answer["top3_confidence"]]}
    raise ValueError("task not valid.")


def handle_request(body: dict, features, infos):
    """callback() minus transport (worker.py:556-563): body = {image_path, question, socket_id, task_id}."""
    task_id = str(body["task_id"])
    answer = custom_prediction(body["question"], [int(task_id)], features, infos, task_id)
    return shape_result(task_id, answer, body["image_path"])


# ------------------------------------------------------------------------------------------ tokenizer
class WordpieceTokenizer(object):
    """bert-base-uncased style tokenizer (lower-case, accent strip, punctuation split, greedy WordPiece) with the two
    methods the worker calls (worker.py:402-403).  Needs the user's vocab.txt: none ships with the reference."""

    def __init__(self, vocab, do_lower_case: bool = True, unk_token="[UNK]", cls_token="[CLS]", sep_token="[SEP]"):
        if isinstance(vocab, (str, os.PathLike)):
            with open(vocab, "r", encoding="utf-8") as f:
                vocab = {line.rstrip("\n"): i for i, line in enumerate(f)}
        self.vocab = dict(vocab)
        self.lower = do_lower_case
        self.unk, self.cls, self.sep = unk_token, cls_token, sep_token

    @staticmethod
    def _is_punct(ch):
        cp = ord(ch)
        if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
            return True
        return unicodedata.category(ch).startswith("P")

    def _basic(self, text):
        if self.lower:
            text = unicodedata.normalize("NFD", text.lower())
            text = "".join(c for c in text if unicodedata.category(c) != "Mn")
        out = []
        for word in text.split():
            cur = ""
            for ch in word:
                if self._is_punct(ch):
                    if cur:
                        out.append(cur)
                    out.append(ch)
                    cur = ""
                else:
                    cur += ch
            if cur:
                out.append(cur)
        return out

    def tokenize(self, text):
        pieces = []
        for word in self._basic(text):
            if len(word) > 100:
                pieces.append(self.unk)
                continue
            start, sub = 0, []
            while start < len(word):
                end, cur = len(word), None
                while start < end:
                    s = word[start:end] if start == 0 else "##" + word[start:end]
                    if s in self.vocab:
                        cur = s
                        break
                    end -= 1
                if cur is None:
                    sub = [self.unk]
                    break
                sub.append(cur)
                start = end
            pieces.extend(sub)
        return pieces

    def encode(self, text):
        return [self.vocab.get(t, self.vocab.get(self.unk, 0)) for t in self.tokenize(text)]

    def add_special_tokens_single_sentence(self, ids):
        return [self.vocab[self.cls]] + list(ids) + [self.vocab[self.sep]]
