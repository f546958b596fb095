"""Zero-shot evaluation and the frozen-CLIP feature records.

Mirrors ``src/training/zero_shot.py:14-52, 138-145`` and ``src/training/train.py:1128-1138,
1345-1381`` of the reference.  The class logits ``100 * I @ classifier`` are never stored:
``latte_nxc_topk`` (csrc/nxc.cu, csrc/nxc_tc.cu) streams the image features once and keeps the
sorted top-k per row in registers.  The reference's towers stay the reference's PyTorch; only
the logits / top-k / accuracy part runs here.  There is no CPU fallback.

The record format written by ``save_feature_records`` is the reference's
``clip_features_{split}.pkl`` (train.py:1365-1381), which its data pipeline reads back with
``load_key_to_clip_prediction`` (data.py:393-396) to supply ``zeroshot_classnames``
(data.py:416, 448) -- the ``zs`` ids of the prototype step (train.py:412-417).
"""

from __future__ import annotations

import contextlib
import os
import pickle
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from . import prototypes as P


def classifier_from_bank(memory_bank, class_names: Sequence[str]) -> torch.Tensor:
    """zero_shot.py:138-145: stack the memory bank in ``class_names`` order, L2-normalise
    rows, return the transposed view [D, C] the reference's ``run`` takes."""
    return P.build_classifier(P.stack_bank(memory_bank, class_names)).T


def _class_rows(classifier: torch.Tensor, dim: int) -> torch.Tensor:
    # the reference's classifier is [D, C] (zero_shot.py:145); the kernel wants class rows
    if classifier.dim() != 2 or classifier.shape[0] != dim:
        raise RuntimeError(f"classifier must be [D={dim}, C], got {tuple(classifier.shape)}")
    return classifier.T.contiguous()


def accuracy(image_features: torch.Tensor, classifier: torch.Tensor, target: torch.Tensor,
             topk: Sequence[int] = (1,)) -> Tuple[List[float], torch.Tensor, torch.Tensor]:
    """train.py:1128-1138 applied to ``100.0 * image_features @ classifier`` (zero_shot.py:40,
    train.py:1352) without forming it.  Returns ``(accs, top_logits, top_class_ids)``:
    ``accs[k]`` is the NUMBER of rows whose target is within the first k classes (a python
    float, as in the reference), the other two are [B, max(topk)] sorted by logit; equal
    logits are ordered by class id (lowest first)."""
    kmax = int(max(topk))
    idx, val = _lib.nxc_topk(image_features, _class_rows(classifier, image_features.shape[1]),
                             kmax, scale=100.0)
    hit = idx.eq(target.to(idx.device).view(-1, 1))
    counts = torch.stack([hit[:, :k].sum() for k in topk]).cpu()      # one D2H for all k
    return [float(c) for c in counts], val, idx


def _autocast(precision: Optional[str]):
    # training/precision.py:5-12
    if precision == "amp":
        return lambda: torch.autocast("cuda", dtype=torch.float16)
    if precision in ("amp_bfloat16", "amp_bf16"):
        return lambda: torch.autocast("cuda", dtype=torch.bfloat16)
    return contextlib.nullcontext


def _input_dtype(precision: Optional[str]):
    # open_clip/model.py:215-221
    if precision in ("bf16", "pure_bf16"):
        return torch.bfloat16
    if precision in ("fp16", "pure_fp16"):
        return torch.float16
    return None


def run(model, classifier: torch.Tensor, dataloader: Iterable, args) -> Tuple[float, float, float]:
    """zero_shot.py:23-52: top-1 / top-5 / top-10 rates over a dataloader of
    ``(image_id, images, target)`` batches.  ``model(image=images)`` is the reference's tower."""
    autocast = _autocast(getattr(args, "precision", None))
    input_dtype = _input_dtype(getattr(args, "precision", None))
    top = torch.zeros(3, dtype=torch.float64)
    n = 0
    with torch.no_grad():
        for _, images, target in dataloader:
            images = images.to(device=args.device, dtype=input_dtype)
            target = target.to(args.device)
            with autocast():
                output = model(image=images)
                feats = output["image_features"] if isinstance(output, dict) else output[0]
            accs, _, _ = accuracy(feats, classifier, target, topk=(1, 5, 10))
            top += torch.tensor(accs, dtype=torch.float64)
            n += images.size(0)
    top = top / n
    return float(top[0]), float(top[1]), float(top[2])


def feature_records(image_ids: Sequence[str], image_features: torch.Tensor,
                    top_class_ids: torch.Tensor, top_logits: torch.Tensor, target: torch.Tensor,
                    class_names: Sequence[str]) -> Dict[str, dict]:
    """train.py:1365-1374: the per-image dicts of ``clip_features_{split}.pkl``.  One bulk
    D2H copy per tensor instead of the reference's five per sample."""
    feats = image_features.detach().cpu().numpy()
    ids = top_class_ids.detach().cpu().numpy()
    logits = top_logits.detach().cpu().numpy()
    gts = target.detach().cpu().numpy()
    records = {}
    for k, image_id in enumerate(image_ids):
        records[image_id] = {
            "image": feats[k],
            "top_class_ids": ids[k],
            "class_names": [class_names[i] for i in ids[k]],
            "top_logit": logits[k],
            "gt_classname": class_names[int(gts[k])],
            "gt_class_id": int(gts[k]),
        }
    return records


def extract_feature_records(model, classifier: torch.Tensor, dataloader: Iterable, args,
                            class_names: Sequence[str]) -> Tuple[Dict[str, dict], Tuple[float, float, float]]:
    """train.py:1336-1376: encode every batch with the reference's image tower, take the
    top-10 zero-shot classes with the fused kernel, build the records and the accuracy rates."""
    autocast = _autocast(getattr(args, "precision", None))
    input_dtype = _input_dtype(getattr(args, "precision", None))
    records: Dict[str, dict] = {}
    top = torch.zeros(3, dtype=torch.float64)
    n = 0
    with torch.no_grad():
        for image_ids, images, target in dataloader:
            images = images.to(device=args.device, dtype=input_dtype, non_blocking=True)
            target = target.to(args.device)
            with autocast():
                feats = model.encode_image(images, normalize=True)
            accs, top_logits, top_ids = accuracy(feats, classifier, target, topk=(1, 5, 10))
            top += torch.tensor(accs, dtype=torch.float64)
            n += images.size(0)
            records.update(feature_records(image_ids, feats, top_ids, top_logits, target, class_names))
    top = top / max(n, 1)
    return records, (float(top[0]), float(top[1]), float(top[2]))


def save_feature_records(records: Dict[str, dict], directory: str, split: str) -> str:
    """train.py:1377-1381: ``{directory}/clip_features_{split}.pkl``."""
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, f"clip_features_{split}.pkl")
    with open(path, "wb") as f:
        pickle.dump(records, f)
    return path


def load_key_to_clip_prediction(path: str) -> Dict[str, dict]:
    """data.py:393-396."""
    with open(path, "rb") as f:
        return pickle.load(f)


def zeroshot_classnames(record: dict, class_per_image: int) -> List[str]:
    """data.py:412-416, 448: the first ``class_per_image`` predicted class names of a record."""
    return [record["class_names"][c] for c in range(class_per_image)]


def zeroshot_class_ids(batch_classnames: Sequence[Sequence[str]], class_names: Sequence[str],
                       device=None) -> torch.Tensor:
    """train.py:390, 412-417: ``classname2id[zeroshot_classnames[i][0]]`` for the whole batch
    as ONE int64 tensor and one H2D copy (the reference writes B device scalars in a loop)."""
    name2id = {c: i for i, c in enumerate(class_names)}
    ids = torch.tensor([name2id[z[0]] for z in batch_classnames], dtype=torch.int64)
    return ids.to(device) if device is not None else ids
