"""Multimodal glue — B200-native mirror of the reference's `cambrian/model/cambrian_arch.py`.

`CambrianMetaModel` builds the same sub-modules with the same attribute / state-dict names (mm_projector,
mm_projector_aux_{i}, vision_sampler_{g}, vision_sampler_layers, vision_query, image_newline; cambrian_arch.py:35-87,
:99-200) and `CambrianMetaForCausalLM.prepare_inputs_labels_for_multimodal` reproduces the static-shape branch the
reference trains with (cambrian_arch.py:340-490, the `IS_XLA_AVAILABLE` side): towers -> aux projectors -> global
context -> connector SVA -> mm_projector -> newline column -> splice into the token embeddings.

Differences that are purely layout, not semantics:
  * the per-query window gather (`rearrange_vision_tower_features_train`, :271-287) is NOT materialised: the aux
    feature grids stay in their natural [B, N_i, 1024] layout and the SVA kernels do the index arithmetic;
  * embedding lookup + image-span replacement + newline append are one gather kernel (`EmbedSpliceFn`).
The per-sample dynamic-shape branch (non-square `image_sizes`, :289-330, :422-451, :493-609) is the next row of
SURVEY.md §8f and raises NotImplementedError for non-square images; square images give identical results on both
branches.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
import torch.nn as nn

from ..autograd import EmbedSpliceFn, ExpandRowsFn, MeanTokensFn
from .multimodal_encoder.builder import build_vision_tower_aux_list
from .multimodal_projector.builder import CBGELU, CBLayerNorm, CBLinear, build_vision_projector
from .vision_sampler import VisionTokenSampler

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


def _kv_sizes(aux_token_lens, query_num):
    return [int(t ** 0.5) // int(query_num ** 0.5) for t in aux_token_lens]


class CambrianMetaModel:
    def __init__(self, config):
        super(CambrianMetaModel, self).__init__(config)
        if hasattr(config, "mm_vision_tower_aux_list"):
            self._build_vision_modules(config, delay_load=True)

    def _build_vision_modules(self, config, delay_load):
        projector_type = getattr(config, "mm_projector_type", "linear")
        self.vision_tower_aux_list = build_vision_tower_aux_list(config, delay_load=delay_load)
        if projector_type == "sva":
            vh = config.vision_hidden_size
            G = config.num_query_group
            self.mm_projector = nn.Sequential(CBLinear(vh * G, config.hidden_size), CBGELU(),
                                              CBLinear(config.hidden_size, config.hidden_size))
            aux_lens = config.mm_vision_tower_aux_token_len_list
            for i, tower in enumerate(self.vision_tower_aux_list):
                setattr(self, f"mm_projector_aux_{i}",
                        nn.Sequential(CBLinear(tower.hidden_size, vh), CBGELU(), CBLinear(vh, vh), CBLayerNorm(vh)))
            T = len(self.vision_tower_aux_list)
            for g in range(G):
                setattr(self, f"vision_sampler_{g}",
                        VisionTokenSampler(vh, vh, [vh] * T, _kv_sizes(aux_lens, config.query_num_list[g]), vh,
                                           config.connector_depth))
            if not config.connector_only:
                sizes = _kv_sizes(aux_lens, config.image_token_len)
                self.vision_sampler_layers = nn.ModuleList([
                    VisionTokenSampler(config.hidden_size, vh, [vh] * T, sizes, vh, 1)
                    for _ in range(config.num_of_vision_sampler_layers)])
            self.vision_query = nn.Parameter(torch.randn((G, vh)) * (1.0 / vh) ** 0.5)      # :161-164
            self.image_newline = nn.Parameter(torch.randn(config.hidden_size) * (1.0 / config.hidden_size) ** 0.5)
        else:
            config.mm_hidden_size = sum(t.hidden_size for t in self.vision_tower_aux_list)
            self.mm_projector = build_vision_projector(config)
            self.image_newline = nn.Parameter(torch.randn(config.hidden_size) * (1.0 / config.hidden_size) ** 0.5)

    def get_vision_tower_aux_list(self):
        return getattr(self, "vision_tower_aux_list", None)

    def initialize_vision_modules(self, model_args, fsdp=None):
        """cambrian_arch.py:99-200: copy the SVA hyper-parameters onto config, build towers and connector modules."""
        cfg = self.config
        cfg.mm_vision_tower_aux_list = model_args.vision_tower_aux_list
        cfg.mm_vision_tower_aux_token_len_list = model_args.vision_tower_aux_token_len_list
        cfg.image_token_len = getattr(model_args, "image_token_len", 576)
        cfg.mm_projector_type = getattr(model_args, "mm_projector_type", "sva")
        cfg.mm_vision_select_layer = getattr(model_args, "mm_vision_select_layer", -2)
        cfg.mm_vision_select_feature = getattr(model_args, "mm_vision_select_feature", "patch")
        cfg.use_mm_proj = True
        if cfg.mm_projector_type == "sva":
            cfg.vision_hidden_size = model_args.vision_hidden_size
            cfg.num_query_group = model_args.num_query_group
            cfg.query_num_list = model_args.query_num_list
            assert cfg.num_query_group == len(cfg.query_num_list)                              # :117
            cfg.connector_depth = model_args.connector_depth
            cfg.connector_only = model_args.connector_only
            cfg.num_of_vision_sampler_layers = getattr(model_args, "num_of_vision_sampler_layers", 0)
            cfg.start_of_vision_sampler_layers = getattr(model_args, "start_of_vision_sampler_layers", 0)
            cfg.stride_of_vision_sampler_layers = getattr(model_args, "stride_of_vision_sampler_layers", 1)
        self._build_vision_modules(cfg, delay_load=False)


class CambrianMetaForCausalLM(ABC):
    @abstractmethod
    def get_model(self):
        pass

    def get_vision_tower_aux_list(self):
        return self.get_model().get_vision_tower_aux_list()

    def encode_images(self, image_aux_list):
        """cambrian_arch.py:332-338."""
        return [tower(img) for img, tower in zip(image_aux_list, self.get_model().get_vision_tower_aux_list())]

    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels,
                                             images, image_aux_attention_masks_list=None, image_sizes=None,
                                             image_positions=None):
        """`image_positions` (extension, host ints): index of the <image> indicator of every sample of an ALREADY
        expanded batch (what the collator knows); skips the device->host scan of input_ids (a stream sync)."""
        model = self.get_model()
        towers = model.get_vision_tower_aux_list()
        if towers is None or images is None or input_ids.shape[1] == 1:                          # :345-346
            return input_ids, position_ids, attention_mask, past_key_values, None, labels, None, None, None, None
        cfg = model.config
        bs = images[0].shape[0]
        q_num = cfg.image_token_len
        q_side = int(q_num ** 0.5)
        if image_sizes is not None and any(int(w) != int(h) for (w, h) in image_sizes):
            raise NotImplementedError("non-square image_sizes need the per-sample dynamic branch "
                                      "(cambrian_arch.py:289-330,:422-451) — SURVEY.md §8f rank 3")
        span = q_num + q_side
        # --- locate the image span of every sample; expand a bare <image> indicator the way the collator does
        #     (train_fsdp.py:1089-1165) when the caller passes un-expanded ids (inference path)
        if image_positions is not None:
            starts = [int(p) for p in image_positions]
        else:
            ids_cpu = input_ids.detach().to("cpu")
            if any(int((row == IMAGE_TOKEN_INDEX).sum()) > 1 for row in ids_cpu):
                raise NotImplementedError("exactly one image per sample (train_fsdp.py:1100)")
            needs_expand = False
            for row in ids_cpu:
                pos = torch.where(row == IMAGE_TOKEN_INDEX)[0]
                if len(pos) == 1:
                    p0 = int(pos[0])
                    if p0 + span > row.shape[0] or bool((row[p0 + 1:p0 + span] != 0).any()):
                        needs_expand = True
            if needs_expand:
                input_ids, labels, attention_mask, position_ids = _expand_image_tokens(
                    ids_cpu, labels, attention_mask, span, input_ids.device)
                ids_cpu = input_ids.detach().to("cpu")
            starts = []
            for row in ids_cpu:
                pos = torch.where(row == IMAGE_TOKEN_INDEX)[0]
                starts.append(int(pos[0]) if len(pos) else -1)
        img_start = torch.tensor(starts, dtype=torch.int32).to(input_ids.device, non_blocking=True)

        feats = self.encode_images(images)                                                      # :366
        feats_final = masks_final = ctx_final = None
        if cfg.mm_projector_type == "sva":
            aux = [getattr(model, f"mm_projector_aux_{i}")(f.to(torch.bfloat16)) for i, f in enumerate(feats)]   # :372-379
            ctx = MeanTokensFn.apply(aux[0])                                                    # :377  [B, vh]
            T = len(aux)
            outs = []
            for g, query_num in enumerate(cfg.query_num_list):
                qs = int(query_num ** 0.5)
                n = bs * query_num
                queries = ExpandRowsFn.apply(model.vision_query[g:g + 1].to(torch.bfloat16), n)      # :383
                ctx_g = ExpandRowsFn.apply(ctx, query_num)                                           # :384
                masks = _masks_for(image_aux_attention_masks_list, aux, qs, n)
                sampler = getattr(model, f"vision_sampler_{g}")
                qf = sampler(queries.view(n, 1, -1), ctx_g.view(n, 1, -1), *aux, *masks, natural_layout=(bs, qs))
                if qs != q_side:
                    raise NotImplementedError("query groups with a side different from the final grid "
                                              "(bilinear resize of the query grid, cambrian_arch.py:394-401)")
                outs.append(qf.view(bs, query_num, -1))
            image_features = outs[0] if len(outs) == 1 else torch.cat(outs, -1)
            feats_final = aux                                                                    # natural layout
            masks_final = _masks_for(image_aux_attention_masks_list, aux, q_side, bs * q_num)
            ctx_final = ExpandRowsFn.apply(ctx, q_num).view(bs * q_num, 1, -1)                   # :406
        else:
            image_features = feats[0] if len(feats) == 1 else torch.cat(feats, -1)              # :408
            image_features = image_features.to(torch.bfloat16)
        image_features = model.mm_projector(image_features)                                     # :410-411
        meta = dict(ids=input_ids.contiguous(), img_start=img_start, q_side=q_side,
                    params=(model.embed_tokens.weight, model.image_newline))
        inputs_embeds = EmbedSpliceFn.apply(meta, model.embed_tokens.weight, image_features.contiguous(),
                                            model.image_newline)
        final_size = [(q_side, q_side)] * bs
        return (None, position_ids, attention_mask, past_key_values, inputs_embeds, labels, feats_final, masks_final,
                final_size, ctx_final)


def _masks_for(mask_list, aux, q_side, n):
    """image_aux_attention_masks_list entries are [B*q^2, r_i^2] bool (collator, train_fsdp.py:1122-1137) or absent."""
    if mask_list is None:
        return [None] * len(aux)
    out = []
    for m, a in zip(mask_list, aux):
        r = int(a.shape[1] ** 0.5) // q_side
        out.append(None if m is None else m.reshape(n, r * r))
    return out


def _expand_image_tokens(ids_cpu, labels, attention_mask, span, device):
    """Host-side mirror of the collator's placeholder expansion (train_fsdp.py:1102-1150) for un-expanded inputs:
    the single -200 id becomes -200 followed by span-1 zero ids; labels get IGNORE_INDEX; positions run on."""
    new_ids, new_labels, new_mask = [], [], []
    lab_cpu = labels.detach().to("cpu") if labels is not None else None
    msk_cpu = attention_mask.detach().to("cpu") if attention_mask is not None else None
    for b, row in enumerate(ids_cpu):
        pos = torch.where(row == IMAGE_TOKEN_INDEX)[0]
        if len(pos) == 0:
            pad = span - 1
            new_ids.append(torch.cat([row, torch.zeros(pad, dtype=row.dtype)]))
            new_labels.append(None if lab_cpu is None else torch.cat([lab_cpu[b], torch.full((pad,), IGNORE_INDEX)]))
            m = torch.ones_like(row, dtype=torch.bool) if msk_cpu is None else msk_cpu[b].bool()
            new_mask.append(torch.cat([m, torch.zeros(pad, dtype=torch.bool)]))
            continue
        p0 = int(pos[0])
        new_ids.append(torch.cat([row[:p0 + 1], torch.zeros(span - 1, dtype=row.dtype), row[p0 + 1:]]))
        if lab_cpu is not None:
            new_labels.append(torch.cat([lab_cpu[b][:p0], torch.full((span,), IGNORE_INDEX), lab_cpu[b][p0 + 1:]]))
        m = torch.ones_like(row, dtype=torch.bool) if msk_cpu is None else msk_cpu[b].bool()
        new_mask.append(torch.cat([m[:p0], torch.ones(span, dtype=torch.bool), m[p0 + 1:]]))
    ids = torch.stack(new_ids).to(device)
    lab = torch.stack(new_labels).to(device) if lab_cpu is not None else None
    mask = torch.stack(new_mask).to(device)
    pos_ids = (mask.long().cumsum(1) - 1).clamp_(min=0)
    return ids, lab, mask, pos_ids
