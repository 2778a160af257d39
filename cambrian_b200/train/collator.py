"""Static-shape batch preparation — mirror of the collator functions the hot path's callers use
(cambrian/train/train_fsdp.py:1039-1165: get_padding_offset, prepare_image_info, prepare_multimodal_data).

These run on the host per batch and produce exactly what `CambrianLlamaForCausalLM.forward` consumes: ids with the
single <image> indicator expanded to 576 + 24 placeholders, labels, the text+image attention mask (letter-boxed image
tokens masked out), position ids that skip masked image tokens, and the per-tower window masks [B*576, r_i^2] for the
SVA.  Pure integer/bool index arithmetic; checked bit-exactly against fixtures generated from the reference functions
(tests/golden/collator.npz).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Sequence

import torch

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


def get_padding_offset(cur_size, original_size):
    """(left, right, top, bottom) padded token rows/cols of a `cur_size` grid holding an aspect-preserving letter-boxed
    image of `original_size` (train_fsdp.py:1039-1055)."""
    cur_w, cur_h = cur_size
    ow, oh = original_size
    if ow / oh > cur_w / cur_h:
        pad = (cur_h - int(oh * (cur_w / ow))) // 2
        return 0, 0, pad, pad
    pad = (cur_w - int(ow * (cur_h / oh))) // 2
    return pad, pad, 0, 0


def prepare_image_info(image_size, image_token_len, newline=False):
    """Validity mask over the token grid (plus a newline column when newline=True) and running position ids
    (train_fsdp.py:1057-1085)."""
    side = int(image_token_len ** 0.5)
    left, right, top, bottom = get_padding_offset((side, side), image_size)
    rows = torch.arange(side)[:, None]
    cols = torch.arange(side + (1 if newline else 0))[None, :]
    ok = (rows >= top) & (rows < side - bottom)
    grid_cols = (cols >= left) & (cols < side - right)
    if newline:
        grid_cols = grid_cols | (cols == side)  # the newline token of every kept row stays visible
    mask = (ok & grid_cols).flatten()
    position_ids = mask.long().cumsum(0) - 1
    return mask, position_ids


def prepare_multimodal_data(input_ids, labels, attention_mask, image_sizes, image_token_len=576,
                            image_aux_token_len_list=(192 * 192,), max_length=2048):
    """train_fsdp.py:1089-1165.  One image per sample (asserted, as in the reference)."""
    side = int(image_token_len ** 0.5)
    span = image_token_len + side
    aux_sides = [int(t ** 0.5) for t in image_aux_token_len_list]
    out_ids, out_labels, out_mask, out_pos = [], [], [], []
    aux_masks = [[] for _ in aux_sides]
    for b in range(input_ids.shape[0]):
        row, lab, msk = input_ids[b], labels[b], attention_mask[b]
        where = torch.where(row == IMAGE_TOKEN_INDEX)[0]
        assert len(where) == 1, len(where)
        p = int(where[0])
        n_after = row.shape[0] - p - 1
        im_mask, im_pos = prepare_image_info(image_sizes[b], image_token_len, newline=True)
        for i, a_side in enumerate(aux_sides):
            assert a_side >= side
            r = a_side // side
            am, _ = prepare_image_info(image_sizes[b], a_side * a_side)
            am = am.view(side, r, side, r).permute(0, 2, 1, 3).reshape(side * side, r * r).clone()
            am[am.sum(1) == 0] = True  # fully padded windows attend to everything (avoids NaN rows)
            aux_masks[i].append(am)
        ids = torch.cat([row[:p + 1], torch.zeros(span - 1, dtype=row.dtype), row[p + 1:]])
        lbs = torch.cat([lab[:p], torch.full((span,), IGNORE_INDEX, dtype=lab.dtype), lab[p + 1:]])
        if bool(msk[p]):
            seg_mask = im_mask.to(msk.dtype)
            seg_pos = (im_pos + p).long()
            nxt = int(seg_pos.max()) + 1
        else:
            seg_mask = torch.zeros(span, dtype=msk.dtype)
            seg_pos = torch.zeros(span, dtype=torch.long)
            nxt = p
        am_full = torch.cat([msk[:p], seg_mask, msk[p + 1:]])
        pos = torch.cat([torch.arange(p), seg_pos, torch.arange(nxt, nxt + n_after)])
        out_ids.append(ids[:max_length])
        out_labels.append(lbs[:max_length])
        out_mask.append(am_full[:max_length])
        out_pos.append(pos[:max_length])
    return (torch.stack(out_ids), torch.stack(out_labels), torch.stack(out_mask), torch.stack(out_pos),
            [torch.stack(m) for m in aux_masks])


def valid_label_ranges(labels, ignore_index: int = IGNORE_INDEX):
    """Host-side hint for the fused lm_head + loss (extension): maximal runs of rows of the flattened [B*S] batch whose
    SHIFTED label (labels[b, s+1] at row (b, s); cambrian_llama.py:411-415) is not ignore_index.  Returns a list of
    (row_start, row_end) python ints and the number of valid labels."""
    lab = labels.detach().to("cpu")
    B, S = lab.shape
    shift = torch.full_like(lab, ignore_index)
    shift[:, :-1] = lab[:, 1:]
    valid = (shift != ignore_index).reshape(-1).numpy()
    import numpy as np
    edges = np.flatnonzero(np.diff(np.concatenate([[0], valid.astype(np.int8), [0]])))
    ranges = [(int(a), int(b)) for a, b in zip(edges[0::2], edges[1::2])]
    return ranges, int(valid.sum())


@dataclass
class DataCollatorForSupervisedDataset(object):
    """Mirror of train_fsdp.py:1168-1236: pads / truncates every sample to `tokenizer.model_max_length`, inserts a dummy
    <image> indicator at `image_position` when a sample has none, expands the indicator into the 576 + 24 slot span and
    emits ids, labels, attention mask, position ids, the per-tower window masks and the stacked per-tower images.

    `emit_hints` (extension; default on): three host-side facts the collator has anyway travel with the batch as extra
    keys of the dict that is splatted into `model(**batch)` — `num_valid_labels` (int), `image_positions` (list[int]) and
    `label_ranges` (list[(row0, row1)] of flattened rows whose SHIFTED label is not ignore_index).  They save the model two
    device->host syncs per step and let the fused lm_head + loss skip rows that cannot contribute; results are identical
    with or without them (tests/test_modules_gpu.py::test_fused_loss_with_label_ranges_matches_full_rows)."""

    tokenizer: object
    image_token_len: int
    image_aux_token_len_list: list
    image_position: int
    emit_hints: bool = True

    def __call__(self, instances: Sequence[Dict]) -> Dict[str, torch.Tensor]:
        input_ids, labels = tuple([inst[key] for inst in instances] for key in ("input_ids", "labels"))
        max_length = self.tokenizer.model_max_length
        pad_id = self.tokenizer.pad_token_id
        left = self.tokenizer.padding_side == "left"

        def fit(t, fill):
            if t.shape[0] >= max_length:
                return t[:max_length]
            n = max_length - t.shape[0]
            return torch.nn.functional.pad(t, (n, 0) if left else (0, n), "constant", fill)

        input_ids = torch.stack([fit(t, pad_id) for t in input_ids])
        labels = torch.stack([fit(t, IGNORE_INDEX) for t in labels])
        attention_mask = input_ids.ne(pad_id)
        p = self.image_position
        for i in range(len(input_ids)):                     # dummy image for text-only samples (:1201-1217)
            if (input_ids[i] == IMAGE_TOKEN_INDEX).sum() == 0:
                for t, fill in ((input_ids, IMAGE_TOKEN_INDEX), (labels, IGNORE_INDEX), (attention_mask, False)):
                    tmp = t[i].clone()
                    tmp[p + 1:] = t[i, p:-1]
                    tmp[p] = fill
                    t[i] = tmp
        image_sizes = [inst["image_size"] for inst in instances]
        ids, labs, mask, pos, aux_masks = prepare_multimodal_data(input_ids, labels, attention_mask, image_sizes,
                                                                  self.image_token_len, self.image_aux_token_len_list,
                                                                  max_length)
        batch = dict(input_ids=ids, labels=labs, attention_mask=mask, position_ids=pos,
                     image_aux_attention_masks_list=aux_masks)
        if "image_aux_list" in instances[0]:
            per_tower = [list(x) for x in zip(*[inst["image_aux_list"] for inst in instances])]
            if all(x is not None and x.shape == per_tower[0][0].shape for x in per_tower[0]):
                batch["images"] = [torch.stack(x) for x in per_tower]
            else:
                batch["images"] = per_tower
        if self.emit_hints:
            ranges, n_valid = valid_label_ranges(labs)
            batch["num_valid_labels"] = n_valid
            batch["label_ranges"] = ranges
            batch["image_positions"] = [int(torch.where(r == IMAGE_TOKEN_INDEX)[0][0]) for r in ids]
        return batch
