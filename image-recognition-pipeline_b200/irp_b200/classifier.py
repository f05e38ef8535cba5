"""Classifier inference on the B200 kernels.

Mirrors the inference side of the reference's ``AnimalClassifier`` (/root/reference/functions/model.py:9-41):
a torchvision ResNet-50 with ``fc = Identity`` followed by ``Sequential(Dropout, Linear(2048, 512), ReLU, Dropout,
Linear(512, num_classes))``.  ``B200Classifier`` takes such a module (``.backbone`` / ``.classifier``), folds the
backbone into the library's tcgen05 trunk and keeps the two Linear layers in fp32 on the device; calling it on a
normalised ``[B,3,224,224]`` tensor returns the logits, like ``model(inputs)`` in eval mode.  ``predict_packed``
is the fused route for decoded uint8 images: the validation transform of functions/dataload.py:51-56
(Resize((256,256)), CenterCrop(224), ToTensor, Normalize) runs on the device, bit-exact with Pillow.

Training (optimizer, dropout in train mode, layer4 fine-tuning) is outside this path; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, ops
from .stage import PackedImages, ResNet50Trunk


class B200Classifier:
    def __init__(self, torch_model: torch.nn.Module, device="cuda:0", max_batch: int = 256):
        backbone = getattr(torch_model, "backbone", None)
        head = getattr(torch_model, "classifier", None)
        if backbone is None or head is None:
            raise ValueError("expected a module with .backbone (ResNet-50, fc=Identity) and .classifier (the head)")
        linears = [m for m in head if isinstance(m, torch.nn.Linear)]
        others = [m for m in head if not isinstance(m, (torch.nn.Linear, torch.nn.ReLU, torch.nn.Dropout))]
        if len(linears) != 2 or others:
            raise ValueError("expected the head Dropout, Linear, ReLU, Dropout, Linear (functions/model.py:29-35)")
        self.device = torch.device(device)
        self.trunk = ResNet50Trunk(backbone, self.device, max_batch=max_batch)
        f32 = dict(device=self.device, dtype=torch.float32)
        self.w1 = linears[0].weight.detach().to(**f32).contiguous()
        self.b1 = linears[0].bias.detach().to(**f32).contiguous()
        self.w2 = linears[1].weight.detach().to(**f32).contiguous()
        self.b2 = linears[1].bias.detach().to(**f32).contiguous()
        self.num_classes = self.w2.shape[0]
        self.max_batch = int(max_batch)

    # nn.Module-flavoured no-ops so the object can be handed to code written for the reference model
    def eval(self):
        return self

    def to(self, *_args, **_kwargs):
        return self

    def head(self, features: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """fp32 [B,2048] pooled features -> (logits [B,C], argmax int32 [B])."""
        return ops.classifier_head(features.contiguous(), self.w1, self.b1, self.w2, self.b2)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """Normalised [B,3,224,224] (any device/dtype) -> fp32 logits [B,C] on the device (functions/model.py:37-40)."""
        logits, _ = self.head(self.trunk.embed_nchw(x))
        return logits

    forward = __call__

    @torch.no_grad()
    def predict_packed(self, packed: PackedImages, from_host: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """Decoded uint8 images (stage.pack_images(images, transform=_lib.TRANSFORM_VAL_256)) -> (logits, argmax):
        val_transform + backbone + head, every step a libirp_b200 kernel."""
        from .stage import taps_for
        n = len(packed)
        # the tap bound of THIS transform (a PackedImages built for the embedding transform carries a smaller one)
        taps = max((taps_for(int(h), int(w), _lib.TRANSFORM_VAL_256) for h, w in packed.hw_np), default=3)
        logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        pred = torch.empty((n,), dtype=torch.int32, device=self.device)
        for lo in range(0, n, self.max_batch):
            hi = min(lo + self.max_batch, n)
            part = packed.slice(lo, hi)
            if from_host or not part.pixels.is_cuda:
                part = part.to(self.device)
            x = ops.preprocess_ex(part.pixels, part.offsets, part.hw, taps, _lib.LAYOUT_NHWC4P,
                                  _lib.TRANSFORM_VAL_256)
            lg, pr = self.head(self.trunk.embed(x))
            logits[lo:hi] = lg
            pred[lo:hi] = pr
        return logits, pred

    def close(self):
        self.trunk.close()


def batch_stats(logits: torch.Tensor, labels: torch.Tensor, criterion=None) -> torch.Tensor:
    """Device fp64 [3] = (sum of weighted cross-entropies, sum of weights, correct) for one batch; `criterion` may
    be an ``nn.CrossEntropyLoss`` whose class ``weight`` is honoured (functions/model.py:47-52)."""
    weight: Optional[torch.Tensor] = None
    if criterion is not None:
        if not isinstance(criterion, torch.nn.CrossEntropyLoss):
            raise TypeError("only nn.CrossEntropyLoss is supported (functions/model.py:47-52)")
        if criterion.reduction != "mean" or criterion.label_smoothing != 0.0:
            raise ValueError("only CrossEntropyLoss(reduction='mean', label_smoothing=0) is supported")
        weight = criterion.weight
    return ops.cross_entropy_stats(logits, labels.to(logits.device, torch.int64), weight)
