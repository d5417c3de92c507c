"""Label / bias-span half of the collator against the golden batch produced by the reference's own
`DataCollatorSpeechSeq2SeqWithPadding` (tests/golden/make_collator_golden.py).  Integer work: exact."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_collate_labels_matches_reference_collator():
    from whisper_context_biasing_b200.collator import collate_labels

    z = np.load(os.path.join(ROOT, "tests", "golden", "collator_golden.npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    assert len(meta["cases"]) == 4
    for c in meta["cases"]:
        feats = []
        for b in range(c["B"]):
            f = {"labels": c["labels_in"][b]}
            if c["bias_spans_in"] is not None:
                f["bias_spans"] = c["bias_spans_in"][b]
            feats.append(f)
        out = collate_labels(feats, c["pad"], c["sot"], c["prev"] if c["with_prev"] else None)
        assert np.array_equal(out["labels"].numpy(), z[c["name"] + "_labels_out"]), c["name"]
        assert np.array_equal(out["decoder_input_ids"].numpy(), z[c["name"] + "_decoder_input_ids"]), c["name"]
        if c["bias_spans_in"] is not None:
            assert np.array_equal(out["bias_spans"].numpy(), z[c["name"] + "_bias_spans_out"]), c["name"]
            assert out["bias_spans"].dtype.is_floating_point is False
        else:
            assert "bias_spans" not in out


def test_collate_labels_edge_cases():
    from whisper_context_biasing_b200.collator import collate_labels

    out = collate_labels([{"labels": [5, 9, 1], "bias_spans": []}, {"labels": [9, 2], "bias_spans": []}], 0, 9, 5)
    assert out["bias_spans"].shape == (2, 1, 1) and int(out["bias_spans"].sum()) == 0      # REF :113-116
    assert out["labels"].tolist() == [[9, 1], [2, -100]]
    assert out["decoder_input_ids"].tolist() == [[5, 9], [9, 2]]
