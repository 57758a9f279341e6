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
# One engine handle serves one host thread at a time (plan cache and last_error are per handle, unlocked): every model call
# of this module -- prediction(), prediction_batch(), the MicroBatchWorker thread -- takes this lock.
import threading as _threading
model_lock = _threading.RLock()
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


def _validate(task_id, infos):
    """Image-count rules of worker.py:255-262."""
    if task_id in SINGLE_IMAGE_TASKS:
        assert len(infos) == 1, "task require 1 image"
    elif task_id in PAIR_TASKS:
        assert len(infos) == 2, "task require 2 images"
    elif task_id in RETRIEVAL_TASKS:
        assert len(infos) > 1 and len(infos) <= 10, "task require 2-10 images"
    else:
        raise ValueError("task not valid.")


def _expand_text(task_id, n_rows, question, input_mask, segment_ids, task_tokens):
    """The text is repeated per image for NLVR2 pairs and retrieval candidates (worker.py:266-284)."""
    rep = 2 if task_id == "12" else (n_rows if task_id == "7" else 1)
    if rep > 1:
        question, input_mask = question.repeat(rep, 1), input_mask.repeat(rep, 1)
        segment_ids, task_tokens = segment_ids.repeat(rep, 1), task_tokens.repeat(rep, 1)
    return question, input_mask, segment_ids, task_tokens


def _decode(task_id, out, spatials, infos):
    """Logits of ONE request -> the reference's answer (worker.py:295-386)."""
    N = len(infos) if task_id == "7" else 3                      # worker.py:250-253
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


def prediction(question, features, spatials, segment_ids, input_mask, image_mask, co_attention_mask, task_tokens,
               task_id, infos):
    """Validate, expand the text for pair / retrieval tasks, run the model once, decode.  worker.py:248-386."""
    _validate(task_id, infos)
    question, input_mask, segment_ids, task_tokens = _expand_text(task_id, features.size(0), question, input_mask,
                                                                  segment_ids, task_tokens)
    with model_lock:
        out = model(question, features, spatials, segment_ids, input_mask, image_mask, co_attention_mask, task_tokens,
                    output_all_attention_masks=True, select=TASK_OUTPUT[task_id])
    return _decode(task_id, out, spatials, infos)


def prediction_batch(requests, bucket: int = 8, select: Optional[int] = None):
    """Micro-batching across requests (the reference handles one message at a time, worker.py:664-673).

    `requests` is a list of argument tuples of `prediction()`.  Requests whose tensors have the same text length, region count
    and feature width go through ONE model call; every pair's forward is independent of its batch neighbours, so each answer is
    what `prediction()` returns for that request alone.  NLVR2 requests (task 12) are placed on even rows -- the binary head
    consumes adjacent rows as one pair (worker.py:266-276) -- with ONE filler row in front when needed.  A request that fails
    validation gets its exception as its answer; the others are unaffected.

    The engine keeps one plan (workspace + CUDA graph) per (batch, text length, regions, select): so that a long-running worker
    sees a small, fixed set of plans, the batch is padded with filler rows to a multiple of `bucket` (even, so the binary
    head always has whole pairs) and `select` defaults to the seven task heads whatever tasks happen to be pending."""
    if bucket < 1:
        raise ValueError("bucket must be >= 1")
    fixed_select = L.OUT_TASK_HEADS if select is None else int(select)
    answers = [None] * len(requests)
    groups = {}
    for i, r in enumerate(requests):
        (question, features, spatials, segment_ids, input_mask, image_mask, co_mask, task_tokens, task_id, infos) = r
        try:
            _validate(task_id, infos)
        except (AssertionError, ValueError) as e:
            answers[i] = e
            continue
        key = (question.shape[1], features.shape[1], features.shape[2])
        groups.setdefault(key, []).append(i)
    for idxs in groups.values():
        rows, spans, need = [], {}, 0
        n = 0
        for i in idxs:
            (question, features, spatials, segment_ids, input_mask, image_mask, co_mask, task_tokens, task_id, infos) = requests[i]
            q, im, seg, tk = _expand_text(task_id, features.size(0), question, input_mask, segment_ids, task_tokens)
            if task_id == "12" and n % 2 == 1:            # ONE filler row so that the pair starts on an even row
                rows.append(tuple(t[-1:] for t in rows[-1]))
                n += 1
            if co_mask is None:
                co_mask = torch.zeros(features.size(0), features.size(1), question.size(1), device=features.device)
            rows.append((q, features, spatials, seg, im, image_mask, co_mask, tk))
            spans[i] = (n, n + features.size(0))
            n += features.size(0)
            need |= TASK_OUTPUT[task_id]
        sel = fixed_select | need
        target = -(-n // bucket) * bucket
        if target % 2 == 1 and (sel & L.OUT_VIL_BINARY_PREDICTION):
            target += 1                                    # the binary head needs an even number of rows
        if target > n:                                     # filler rows: copies of the last row, results discarded
            last = tuple(t[-1:] for t in rows[-1])
            rows.append(tuple(t.expand(target - n, *t.shape[1:]) for t in last))
            n = target
        cat = [torch.cat([r[k] for r in rows], dim=0) for k in range(8)]
        with model_lock:
            out = model(*cat, output_all_attention_masks=True, select=sel)
        for i in idxs:
            lo, hi = spans[i]
            task_id, infos, spatials = requests[i][8], requests[i][9], requests[i][2]
            sl = [None if o is None or isinstance(o, list) else o[lo:hi] for o in out[:9]] + [[]]
            if out[3] is not None:
                sl[3] = out[3][lo // 2:hi // 2] if task_id == "12" else out[3][0:0]
            answers[i] = _decode(task_id, tuple(sl), spatials, infos)
    return answers


def build_inputs(query, task, features, infos, device, tok=None):
    """The tensor construction half of custom_prediction (worker.py:402-455), returned instead of consumed."""
    tokens, input_mask, segment_ids = tokenize_query(query, tok)
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


def tokenize_query(query, tok=None):
    """worker.py:402-414: encode, add [CLS]/[SEP], pad (never truncate) to 37 -> (tokens, input_mask, segment_ids) lists."""
    tok = tok if tok is not None else tokenizer
    if tok is None:
        raise L.VilbertB200Error("no tokenizer: pass vocab_file= (bert-base-uncased vocab.txt) to load_vilbert_model")
    tokens = tok.add_special_tokens_single_sentence(tok.encode(query))
    segment_ids, input_mask = [0] * len(tokens), [1] * len(tokens)
    if len(tokens) < MAX_LENGTH:
        pad = [0] * (MAX_LENGTH - len(tokens))
        tokens, input_mask, segment_ids = tokens + pad, input_mask + pad, segment_ids + pad
    return tokens, input_mask, segment_ids


def custom_prediction(query, task, features, infos, task_id):
    """worker.py:388-458 (the GuessWhat dialog rewrite at 391-400 builds `tokens` and then discards it).

    The image half of the reference's tensor construction (global mean row, cat, box normalisation, [0,0,1,1,1], masks,
    stack -- worker.py:422-453) runs inside the engine's region-pack kernel: the detector's box features go to the device
    as they are and are written straight into the image-embedding GEMM's 16-bit operand (`model.forward_regions`).  Images
    with different box counts -- which `torch.stack` at worker.py:452 cannot take -- are padded and masked."""
    _validate(task_id, infos)
    device = torch.device("cuda", model._device)
    tokens, input_mask, segment_ids = tokenize_query(query)
    n_img = len(infos)
    rep = n_img if task_id in PAIR_TASKS + RETRIEVAL_TASKS else 1          # text repeated per image (worker.py:266-284)
    text = torch.tensor([tokens] * rep, dtype=torch.long)
    mask_t = torch.tensor([input_mask] * rep, dtype=torch.long)
    seg_t = torch.tensor([segment_ids] * rep, dtype=torch.long)
    task_t = torch.tensor(np.array(task), dtype=torch.long).view(1, -1)[:, :1].repeat(rep, 1)
    counts = [int(f.shape[0]) for f in features]
    n = max(counts)
    feats = torch.zeros(n_img, n, features[0].shape[1], dtype=torch.float32)
    boxes = torch.zeros(n_img, n, 4, dtype=torch.float32)
    for i, (f, info) in enumerate(zip(features, infos)):
        feats[i, :counts[i]] = torch.as_tensor(f, dtype=torch.float32)
        boxes[i, :counts[i]] = torch.as_tensor(np.asarray(info["bbox"], dtype=np.float32))[:counts[i], :4]
    wh = torch.tensor([[float(info["image_width"]), float(info["image_height"])] for info in infos], dtype=torch.float32)
    num_boxes = None if min(counts) == n else torch.tensor(counts, dtype=torch.int32)
    with model_lock:
        out, spatials = model.forward_regions(text.to(device), seg_t.to(device), mask_t.to(device), task_t.to(device),
                                              feats.to(device), boxes.to(device), wh.to(device), num_boxes,
                                              select=TASK_OUTPUT[task_id], output_all_attention_masks=True)
    return _decode(task_id, out, spatials, infos)


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
class MicroBatchWorker(object):
    """Dependency-light stand-in for the reference's consumer loop (`callback` / `main`, worker.py:542-673) with micro-batching.

    `submit(body, features, infos)` takes the JSON message of demo/sender.py:19-24 (`image_path, question, socket_id, task_id`)
    plus the detector output for its images (the detector itself, worker.py:59-223, is out of scope) and returns a
    `concurrent.futures.Future` whose result is `{"socket_id": ..., "result": <the dict callback() pushes to the WebSocket,
    worker.py:564-649>}`.  A background thread drains the queue: it waits at most `max_wait_ms` after the first pending message
    for up to `max_rows` image rows, runs them through `prediction_batch` (one model call per shape group) and resolves the
    futures; a failing request resolves to its exception, like the try/except of callback (worker.py:653-655)."""

    def __init__(self, max_rows: int = 64, max_wait_ms: float = 2.0, bucket: int = 8):
        import queue
        import threading
        self.max_rows, self.max_wait, self.bucket = int(max_rows), float(max_wait_ms) * 1e-3, int(bucket)
        self._q = queue.Queue()
        self._stop = False
        self.batches = []                      # rows per model call group, for tests / monitoring
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def submit(self, body: dict, features, infos):
        from concurrent.futures import CancelledError, Future
        fut = Future()
        if self._stop:
            fut.set_exception(CancelledError("MicroBatchWorker closed"))
            return fut
        self._q.put((body, features, infos, fut))
        return fut

    def close(self):
        """Stop the consumer thread.  Messages still queued (behind the sentinel, or submitted after close) are cancelled:
        their futures raise `concurrent.futures.CancelledError` instead of blocking their callers for ever."""
        import queue
        from concurrent.futures import CancelledError
        self._stop = True
        self._q.put(None)
        self._thread.join(timeout=30.0)
        while True:
            try:
                item = self._q.get_nowait()
            except queue.Empty:
                break
            if item is not None and not item[3].done():
                item[3].set_exception(CancelledError("MicroBatchWorker closed"))

    def _loop(self):
        import queue
        import time
        while not self._stop:
            item = self._q.get()
            if item is None:
                break
            pending, rows = [item], len(item[2])
            deadline = time.monotonic() + self.max_wait
            while rows < self.max_rows:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                try:
                    nxt = self._q.get(timeout=left)
                except queue.Empty:
                    break
                if nxt is None:
                    self._stop = True
                    break
                pending.append(nxt)
                rows += len(nxt[2])
            self._run(pending)

    def _run(self, pending):
        device = torch.device("cuda", model._device) if torch.cuda.is_available() else torch.device("cpu")
        reqs, live = [], []
        for body, features, infos, fut in pending:
            try:
                task_id = str(body["task_id"])
                reqs.append(build_inputs(body["question"], [int(task_id)], features, infos, device) + (task_id, infos))
                live.append((body, fut, task_id))
            except Exception as e:             # noqa: BLE001 -- mirrors callback()'s catch-all (worker.py:653-655)
                fut.set_exception(e)
        if not reqs:
            return
        try:
            answers = prediction_batch(reqs, bucket=self.bucket)
        except Exception as e:                 # noqa: BLE001
            for _, fut, _ in live:
                fut.set_exception(e)
            return
        self.batches.append(sum(r[1].shape[0] for r in reqs))
        for (body, fut, task_id), ans in zip(live, answers):
            if isinstance(ans, Exception):
                fut.set_exception(ans)
            else:
                try:
                    fut.set_result({"socket_id": body.get("socket_id"), "result": shape_result(task_id, ans, body["image_path"])})
                except Exception as e:         # noqa: BLE001
                    fut.set_exception(e)


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
