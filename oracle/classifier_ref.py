"""ORACLE (test infrastructure, never on the product path): the reference's classifier inference on the CPU.

SURVEY.md section 8(f) row N1.  Restates, with the same torch modules,

* ``AnimalClassifier`` (/root/reference/functions/model.py:9-41): torchvision ResNet-50 with ``fc = Identity``
  followed by ``Sequential(Dropout, Linear(2048, 512), ReLU, Dropout, Linear(512, num_classes))``; in ``eval()``
  mode the dropouts are identities;
* ``evaluate_full`` (/root/reference/functions/train.py:192-238): per batch ``outputs = model(inputs)``,
  ``loss = criterion(outputs, labels)`` (mean cross-entropy), ``running_loss += loss * batch``, ``argmax`` ->
  predictions; returns ``(sum loss / total, 100 * correct / total, preds, labels)``.

The reference constructor downloads IMAGENET1K_V2 weights (no network here); like the embedding oracle
(stage_ref.full_resnet50) the weights are torch's seeded random init instead, built in the SAME order as the
reference constructor (backbone, then the two Linear layers) so that, under the same seed, this module and the
patched reference class hold identical parameters.

Pinned: tests/test_oracle.py compares ``logits`` / ``evaluate_full`` with tests/golden/classifier.npz, which
oracle/make_golden.py produced by running the reference's own AnimalClassifier and evaluate_full (imported from
/root/reference, ``resnet50`` patched to skip the download).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def build_classifier(num_classes: int = 10, seed: int = 1234, dropout_rate: float = 0.2) -> nn.Module:
    """Same module tree and parameter-creation order as functions/model.py:10-35."""
    from torchvision.models import resnet50

    torch.manual_seed(seed)
    backbone = resnet50(weights=None)
    in_features = backbone.fc.in_features
    backbone.fc = nn.Identity()
    classifier = nn.Sequential(nn.Dropout(dropout_rate), nn.Linear(in_features, 512), nn.ReLU(),
                               nn.Dropout(dropout_rate), nn.Linear(512, num_classes))
    model = nn.Module()
    model.backbone = backbone
    model.classifier = classifier
    model.forward = lambda x: classifier(backbone(x))  # functions/model.py:37-40
    return model.eval()


def synthetic_eval_set(n: int, num_classes: int = 10, seed: int = 0):
    """(images, labels): HWC uint8 arrays of assorted sizes and round-robin labels (what a decoded shard yields)."""
    from . import synth

    images, _ = synth.config1_images(n, num_classes, seed=seed)
    rng = np.random.default_rng(seed + 1)
    out = []
    for i, im in enumerate(images):
        if i % 3 == 1:      # a third of the set is not square / not 224: Resize((256,256)) changes the aspect ratio
            h, w = int(rng.integers(180, 420)), int(rng.integers(180, 520))
            ys = (np.arange(h) * im.shape[0] // h)
            xs = (np.arange(w) * im.shape[1] // w)
            im = np.ascontiguousarray(im[ys][:, xs])
        out.append(im)
    labels = np.arange(n, dtype=np.int64) % num_classes
    return out, labels


def val_batches(images, labels, batch_size: int):
    """List of (float32 [b,3,224,224], int64 [b]) -- what the reference's DataLoader yields after val_transform."""
    from . import pil_resample

    batches = []
    for s in range(0, len(images), batch_size):
        x = np.stack([pil_resample.val_transform(im) for im in images[s:s + batch_size]])
        batches.append((torch.from_numpy(x), torch.from_numpy(np.asarray(labels[s:s + batch_size], np.int64))))
    return batches


@torch.no_grad()
def logits(model: nn.Module, x: torch.Tensor) -> np.ndarray:
    return model.forward(x).numpy()


@torch.no_grad()
def evaluate_full(model: nn.Module, batches, criterion=None):
    """train.py:192-238 on a list of batches; returns (epoch_loss, epoch_acc, preds, labels)."""
    criterion = criterion or nn.CrossEntropyLoss()
    running_loss, correct, total = 0.0, 0, 0
    preds, labs = [], []
    for inputs, labels in batches:
        outputs = model.forward(inputs)
        loss = criterion(outputs, labels)
        running_loss += loss.item() * inputs.size(0)
        _, predicted = torch.max(outputs, 1)
        total += labels.size(0)
        correct += (predicted == labels).sum().item()
        preds.extend(predicted.numpy())
        labs.extend(labels.numpy())
    if total == 0:
        return 0, 0, [], []
    return running_loss / total, 100 * correct / total, preds, labs
