"""The gradient source around the hot path, with the reference's signatures
(src/core/loss_helpers.py:7-32).  wav2vec2 + CTC stay on PyTorch/cuDNN untouched; only the WER counters
run in libpaa.so (the reference uses HF evaluate/jiwer, which this image does not ship)."""
import re

import torch

try:
    from .. import paa_lib as L
except ImportError:
    import paa_lib as L

# the 32-token character vocabulary of wav2vec2 CTC checkpoints
VOCAB = ["<pad>", "<s>", "</s>", "<unk>", "|"] + list("ETAONIHSRDLUMWCFGYPBVK'XJQZ")
_TOKEN = {c: i for i, c in enumerate(VOCAB)}


def clean_transcripts(texts):
    """Drop <unk>, lower-case, collapse whitespace (loss_helpers.py:7-9)."""
    return [re.sub(r"\s+", " ", t.replace("<unk>", "").lower()).strip() for t in texts]


def encode_labels(texts, device):
    """What ``processor(text=..., padding=True).input_ids`` gives for the vocabulary above, pads as -100."""
    rows = [[_TOKEN.get("|" if c == " " else c, _TOKEN["<unk>"]) for c in t.upper()] for t in texts]
    width = max((len(r) for r in rows), default=0)
    return torch.tensor([r + [-100] * (width - len(r)) for r in rows], dtype=torch.long, device=device)


def greedy_decode(pred_ids):
    """CTC collapse of argmax ids -> text, special tokens skipped (processor.batch_decode(skip_special_tokens=True))."""
    out = []
    for row in pred_ids.tolist():
        chars, last = [], None
        for tok in row:
            if tok != last and tok > 3:
                chars.append(" " if tok == 4 else VOCAB[tok])
            last = tok
        out.append(" ".join("".join(chars).split()))
    return out


def get_loss_for_training(model, data, target_texts, processor, args):
    """CTC loss (sum) and logits for the perturbed batch (loss_helpers.py:12-23)."""
    if args.attack_mode == "targeted":
        target_texts = [" ".join([args.target] * args.target_reps)] * len(data)
    target_texts = clean_transcripts(target_texts)
    if processor is None:
        labels = encode_labels(target_texts, args.device)
    else:
        labels = processor(text=target_texts, return_tensors="pt", padding=True).input_ids.to(args.device)
        labels[labels == processor.tokenizer.pad_token_id] = -100
    outputs = model(input_values=data, labels=labels)
    return outputs.loss, outputs.logits


class WerMetric:
    """``.compute(predictions=, references=)`` like HF evaluate's "wer": sum(S+D+I) / sum(reference words),
    counted by libpaa.so's paa_wer_counts.  The raw counters are kept for the cross-GPU all-reduce."""

    def __init__(self):
        self.errors = 0
        self.words = 0

    def compute(self, predictions, references):
        e, w = L.wer_counts(list(references), list(predictions))
        self.errors += e
        self.words += w
        return e / w if w else 0.0


def compute_wer(logits, target_texts, processor, wer_metric):
    """argmax -> CTC decode -> WER against the cleaned references (loss_helpers.py:25-32)."""
    return wer_from_ids(torch.argmax(logits, dim=-1), target_texts, processor, wer_metric)


def wer_from_ids(pred_ids, target_texts, processor, wer_metric):
    """The decode + WER half of compute_wer, for greedy ids that were kept on the device until the epoch ended."""
    if processor is None:
        pred_texts = greedy_decode(pred_ids)
    else:
        pred_texts = processor.batch_decode(pred_ids, skip_special_tokens=True)
    pred_texts = [p.strip().lower() for p in pred_texts]
    ref_texts = [t.lower() for t in clean_transcripts(target_texts)]
    return wer_metric.compute(predictions=pred_texts, references=ref_texts)
