"""Drop-in for the inference half of the reference's functions/train.py: ``evaluate_full`` with the same signature
and return value (/root/reference/functions/train.py:192-238), running on the B200 kernels.

``model`` is an ``irp_b200.classifier.B200Classifier`` (built from the reference's AnimalClassifier instance);
``test_loader`` yields ``(inputs [B,3,224,224] float, labels [B] int)`` exactly like the reference's DataLoader.
The per-batch loss / correct counts are computed on the device and read back ONCE at the end (the reference
synchronises on ``loss.item()`` every batch).  Training functions of that module are outside the hot path.
"""
from __future__ import annotations

import torch
from tqdm.auto import tqdm

from irp_b200.classifier import B200Classifier, batch_stats


def evaluate_full(model, test_loader, criterion, disable_progress=False):
    """Evaluate model on the full dataset without batch limits -> (epoch_loss, epoch_acc, all_preds, all_labels)."""
    if not isinstance(model, B200Classifier):
        raise TypeError("evaluate_full of the B200 path needs a B200Classifier (there is no CPU fallback)")
    model.eval()
    stats, sizes, preds, labs = [], [], [], []
    print("Evaluating model on full test set (no batch limit)...")
    with torch.no_grad():
        for inputs, labels in tqdm(test_loader, desc="Full Evaluation", disable=disable_progress):
            logits = model(inputs)
            labels_dev = labels.to(logits.device, torch.int64)
            stats.append(batch_stats(logits, labels_dev, criterion))
            sizes.append(int(labels.shape[0]))
            preds.append(torch.argmax(logits, 1))
            labs.append(labels_dev)
    total = sum(sizes)
    if total > 0:
        s = torch.stack(stats).cpu().numpy()  # one device->host read for the whole evaluation
        # per batch: loss = sum(w*ce)/sum(w) (CrossEntropyLoss mean), running_loss += loss * batch size
        running_loss = float(sum((row[0] / row[1] if row[1] > 0 else 0.0) * b for row, b in zip(s, sizes)))
        correct = int(round(float(s[:, 2].sum())))
        epoch_loss = running_loss / total
        epoch_acc = 100 * correct / total
        all_preds = list(torch.cat(preds).cpu().numpy())
        all_labels = list(torch.cat(labs).cpu().numpy())
    else:
        epoch_loss, epoch_acc, all_preds, all_labels = 0, 0, [], []
    print(f"Evaluated on {total} samples: Loss={epoch_loss:.4f}, Accuracy={epoch_acc:.2f}%")
    return epoch_loss, epoch_acc, all_preds, all_labels
