"""Static-shape batch preparation — mirror of the collator functions the hot path's callers use
(cambrian/train/train_fsdp.py:1039-1165: get_padding_offset, prepare_image_info, prepare_multimodal_data).

These run on the host per batch and produce exactly what `CambrianLlamaForCausalLM.forward` consumes: ids with the
single <image> indicator expanded to 576 + 24 placeholders, labels, the text+image attention mask (letter-boxed image
tokens masked out), position ids that skip masked image tokens, and the per-tower window masks [B*576, r_i^2] for the
SVA.  Pure integer/bool index arithmetic; checked bit-exactly against fixtures generated from the reference functions
(tests/golden/collator.npz).
"""
from __future__ import annotations

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
