"""Lazy JSON datasets + collators that produce the batch layout the trainers (and, through them, the CUDA path) consume
(SURVEY section 8b "Batch layout", 8f row N1).

Reference: llavamod/data/dataset.py -- LazySupervisedDataset :25-165, DataCollatorForSupervisedDataset :168-232,
LazyDPODataset :253-432, DataCollatorForDPODataset :435-505, make_*_data_module :235-246,508-517.

Layout kept: mimic batches {input_ids [B,Tt] int64 (image placeholder -200, pad = tokenizer.pad_token_id), labels [B,Tt] int64 (-100
ignored), attention_mask [B,Tt] bool, images: flat Python list of [3,H,W] float tensors}; preference batches the chosen_* / rejected_*
triples + one shared `images` list.  One design instead of two copies: a single sample loader parameterised by which conversation
fields a record carries, and a single collator parameterised by the field prefixes.  Video records need the LanguageBind towers, which
are outside the hot path: they raise."""
import json
import os
import random
from dataclasses import dataclass
from typing import Dict, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

from ..constants import IGNORE_INDEX, MAX_IMAGE_LENGTH
from .data_utils import deep_copy_turns, expand2square, preprocess, preprocess_multimodal

local_rank = None


def rank0_print(*args):
    if local_rank in (0, None) and int(os.environ.get("RANK", "0")) == 0:
        print(*args)


def order_pick_k(items, k):
    """Random subset of k items in their original order (reference llavamod/utils.py:17-28)."""
    if len(items) <= k:
        return items
    keep = sorted(np.argsort(np.random.random(len(items)))[:k])
    print(f"WARNING: total file: {len(items)}, random pick: {k}. (ignored)")
    return [items[i] for i in keep]


class _LazyConversationDataset(Dataset):
    """Records are read up front, tokenised / image-processed on access.  `fields` maps output prefix -> record key holding the turns."""
    fields: Dict[str, str] = {}

    def __init__(self, data_path, tokenizer, data_args):
        super().__init__()
        records = []
        for path in ([data_path] if isinstance(data_path, str) else data_path):
            rank0_print("#### read from", path)
            with open(path, "r") as f:
                chunk = json.load(f)
            rank0_print("#### len: ", len(chunk))
            for rec in chunk:
                rec["id"] = len(records)
                records.append(rec)
        rank0_print("#### total len:", len(records))
        self.tokenizer = tokenizer
        self.list_data_dict = records
        self.data_args = data_args

    def __len__(self):
        return len(self.list_data_dict)

    @property
    def modality_lengths(self):
        """Whitespace word count of the `conversations` turns, negated for text-only records (dataset.py:52-61,280-289; the preference
        dataset reads the same key, so its records need it for --group_by_modality_length just like in the reference)."""
        out = []
        for rec in self.list_data_dict:
            n = sum(len(turn["value"].split()) for turn in rec["conversations"])
            out.append(n if ("image" in rec or "video" in rec) else -n)
        return out

    # -- pieces of __getitem__ ---------------------------------------------------------------------------------------------------
    def _load_images(self, rec):
        """PIL -> processor tensors; unreadable files become a black 224x224 image (dataset.py:71-92)."""
        from PIL import Image
        args = self.data_args
        proc = args.image_processor
        files = rec["image"] if isinstance(rec["image"], list) else [rec["image"]]
        pils = []
        for name in order_pick_k(files, MAX_IMAGE_LENGTH):
            try:
                pils.append(Image.open(os.path.join(args.image_folder, name)).convert("RGB"))
            except Exception as e:  # noqa: BLE001  (the reference swallows every loader error the same way)
                print(f"Error opening image {name}: {e}, using fallback image.")
                pils.append(Image.new(mode="RGB", size=(224, 224), color=(0, 0, 0)))
        if args.image_aspect_ratio == "pad":
            fill = tuple(int(c * 255) for c in proc.image_mean)
            pils = [expand2square(im, fill) for im in pils]
        return [proc.preprocess(im, return_tensors="pt")["pixel_values"][0] for im in pils]

    def _blank_image(self):
        """Text-only record under a multimodal model: one all-zero image so the tower / projector still run (dataset.py:150-157)."""
        proc = self.data_args.image_processor
        size = proc.crop_size if hasattr(proc, "crop_size") else proc.size
        return [torch.zeros(3, size["height"], size["width"])]

    def _build(self, i):
        rec = self.list_data_dict[i]
        if "video" in rec:
            raise NotImplementedError("video records need the LanguageBind video tower (outside the distillation hot path)")
        has_image = "image" in rec
        images = self._load_images(rec) if has_image else None
        item = {}
        for prefix, key in self.fields.items():
            turns = deep_copy_turns([rec], key)
            if has_image or key != "conversations":        # dataset.py:94,143 vs :382-385: only the text-only SFT branch skips this
                turns = preprocess_multimodal(turns, self.data_args)
            enc = preprocess(turns, self.tokenizer, has_image=has_image)
            item[prefix + "input_ids"] = enc["input_ids"][0]
            item[prefix + "labels"] = enc["labels"][0]
        if has_image:
            item["image"] = images
        elif self.data_args.is_multimodal:
            item["image"] = self._blank_image()
        return item

    def __getitem__(self, i):
        """A record that fails to load is replaced by a random other one, as in the reference (dataset.py:161-163)."""
        try:
            return self._build(i)
        except Exception as e:  # noqa: BLE001
            print(f"Error with {e}")
            return self.__getitem__(random.randint(0, len(self) - 1))


class LazySupervisedDataset(_LazyConversationDataset):
    """Mimic / SFT records: {"image": file | [files], "conversations": [{"from": "human"|"gpt", "value": str}, ...]}."""
    fields = {"": "conversations"}


class LazyDPODataset(_LazyConversationDataset):
    """Preference records: {"image": ..., "chosen": [turns], "rejected": [turns]} (+ "conversations" if length grouping is on)."""
    fields = {"chosen_": "chosen", "rejected_": "rejected"}


def _flatten_images(instances):
    """[[img], [img, img], ...] -> flat list, sample order kept (dataset.py:213-225): the splice consumes them in this order."""
    flat = []
    for inst in instances:
        im = inst["image"]
        flat.extend(im if type(im) is list else [im])
    return flat


@dataclass
class _Collator:
    tokenizer: object
    prefixes = ("",)
    truncate = True

    def __call__(self, instances: Sequence[Dict]) -> Dict[str, torch.Tensor]:
        pad = self.tokenizer.pad_token_id
        batch = {}
        for p in self.prefixes:
            ids = torch.nn.utils.rnn.pad_sequence([x[p + "input_ids"] for x in instances], batch_first=True, padding_value=pad)
            lab = torch.nn.utils.rnn.pad_sequence([x[p + "labels"] for x in instances], batch_first=True, padding_value=IGNORE_INDEX)
            if self.truncate:                               # only the supervised collator truncates (dataset.py:185-186 vs :462)
                ids = ids[:, : self.tokenizer.model_max_length]
                lab = lab[:, : self.tokenizer.model_max_length]
            batch[p + "input_ids"] = ids
            batch[p + "labels"] = lab
            batch[p + "attention_mask"] = ids.ne(pad)       # NB: a pad id that also occurs in the text is masked too, as in the reference
        if "image" not in instances[0]:
            raise ValueError(f"pretrain, {instances}")
        batch["images"] = _flatten_images(instances)
        return batch


@dataclass
class DataCollatorForSupervisedDataset(_Collator):
    prefixes = ("",)
    truncate = True


@dataclass
class DataCollatorForDPODataset(_Collator):
    prefixes = ("chosen_", "rejected_")
    truncate = False


def make_supervised_data_module(tokenizer, data_args) -> Dict:
    return dict(train_dataset=LazySupervisedDataset(tokenizer=tokenizer, data_path=data_args.data_path, data_args=data_args),
                eval_dataset=None, data_collator=DataCollatorForSupervisedDataset(tokenizer=tokenizer))


def make_dpo_data_module(tokenizer, data_args) -> Dict:
    return dict(train_dataset=LazyDPODataset(tokenizer=tokenizer, data_path=data_args.data_path, data_args=data_args),
                eval_dataset=None, data_collator=DataCollatorForDPODataset(tokenizer=tokenizer))
