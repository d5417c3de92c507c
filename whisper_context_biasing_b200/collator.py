"""Device-side collator: same batch dict as the reference's `DataCollatorSpeechSeq2SeqWithPadding`
(REF/data_utils/data_collator.py:27-127), but the log-mel features are computed for the whole batch
in ONE call of the B200 extractor, in the main process, and stay in HBM.

Differences from the reference, on purpose:
  * items may carry raw PCM under "audio" (1-D float32 / int16 numpy) instead of precomputed
    "input_features" -- the dataset then skips REF/data_utils/data_loader.py:171-172 and the two
    host copies of the features disappear; items that already carry "input_features" are stacked
    exactly like `feature_extractor.pad(..., padding="longest", return_tensors="pt")` (:71-76);
  * label / bias-span logic (:78-125) is restated with plain torch ops and needs only the pad token
    id, not a tokenizer object.  It is bit-identical to the reference on the committed golden batch
    (`tests/golden/collator_golden.npz`, produced by the reference's own class).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np


def pad_labels(label_lists: List[List[int]], pad_token_id: int):
    """`tokenizer.pad({"input_ids": ...}, padding="longest", return_tensors="pt")` (REF :78-84):
    right-pad with the pad token, attention_mask 1 on real tokens."""
    import torch

    T = max((len(x) for x in label_lists), default=0)
    ids = torch.full((len(label_lists), T), int(pad_token_id), dtype=torch.long)
    mask = torch.zeros((len(label_lists), T), dtype=torch.long)
    for b, x in enumerate(label_lists):
        n = len(x)
        if n:
            ids[b, :n] = torch.as_tensor(list(x), dtype=torch.long)
            mask[b, :n] = 1
    return ids, mask


def collate_labels(features: List[Dict[str, Any]], pad_token_id: int, decoder_start_token_id: int,
                   decoder_prev_token_id: Optional[int]) -> Dict[str, Any]:
    """REF/data_utils/data_collator.py:78-125 -- shift, -100 on padding and on the prompt before
    `<|startoftranscript|>`, bias spans padded with the literal 50256."""
    import torch

    ids, mask = pad_labels([f["labels"] for f in features], pad_token_id)
    decoder_input_ids = ids[:, :-1]                                    # :90
    labels = ids[:, 1:]                                                # :91
    labels_mask = mask[:, 1:]                                          # :94
    labels = labels.masked_fill(labels_mask.ne(1), -100)               # :96
    if decoder_prev_token_id is not None:                              # :98-102
        bos_index = torch.argmax((labels == decoder_start_token_id).long(), dim=1)
        prompt_mask = torch.arange(labels.shape[1]) < bos_index[:, None]
        labels = torch.where(prompt_mask, -100, labels)
    out: Dict[str, Any] = {"labels": labels, "decoder_input_ids": decoder_input_ids}
    if features and "bias_spans" in features[0]:                       # :107-125
        raw_spans = [f["bias_spans"] for f in features]
        max_span_len = max((len(span) for sample in raw_spans for span in sample), default=0)
        max_n_spans = max((len(sample) for sample in raw_spans), default=0)
        if max_span_len == 0 or max_n_spans == 0:
            out["bias_spans"] = torch.zeros((len(raw_spans), 1, 1), dtype=torch.long)
        else:
            padded = [[list(span) + [50256] * (max_span_len - len(span)) for span in sample]
                      + [[50256] * max_span_len] * (max_n_spans - len(sample)) for sample in raw_spans]
            out["bias_spans"] = torch.tensor(padded, dtype=torch.long)
    return out


class B200DataCollatorSpeechSeq2SeqWithPadding:
    """Drop-in for the reference collator with the feature extraction moved onto the B200.

    Args mirror the reference dataclass (`processor`, `decoder_start_token_id`,
    `decoder_prev_token_id`); `processor.feature_extractor` must be a `B200WhisperFeatureExtractor`
    and `processor.tokenizer.pad_token_id` supplies the pad id (or pass `pad_token_id=`).
    """

    def __init__(self, processor: Any = None, decoder_start_token_id: int = None, decoder_prev_token_id: Optional[int] = None,
                 feature_extractor: Any = None, pad_token_id: Optional[int] = None, labels_on_device: bool = False):
        self.processor = processor
        self.feature_extractor = feature_extractor if feature_extractor is not None else processor.feature_extractor
        if pad_token_id is None:
            pad_token_id = processor.tokenizer.pad_token_id
        if decoder_start_token_id is None:
            raise ValueError("decoder_start_token_id is required")
        self.pad_token_id = int(pad_token_id)
        self.decoder_start_token_id = int(decoder_start_token_id)
        self.decoder_prev_token_id = decoder_prev_token_id
        self.labels_on_device = labels_on_device

    def __call__(self, features: List[Dict[str, Any]]):
        from .feature_extraction import LogMelBatch

        fe = self.feature_extractor
        name = fe.model_input_names[0]
        if features and "audio" in features[0]:
            clips = [np.ascontiguousarray(np.asarray(f["audio"]).reshape(-1)) for f in features]
            if clips[0].dtype not in (np.float32, np.int16):
                clips = [c.astype(np.float32) for c in clips]
            feats = fe.extract_host(clips)                       # one batched call; stays in HBM
            batch = LogMelBatch({name: feats})
        else:
            batch = fe.pad({name: [f[name] for f in features]}, padding="longest", return_tensors="pt")
        lab = collate_labels(features, self.pad_token_id, self.decoder_start_token_id, self.decoder_prev_token_id)
        if self.labels_on_device:
            lab = {k: v.to(fe.device, non_blocking=True) for k, v in lab.items()}
        for k, v in lab.items():
            batch[k] = v
        return batch
