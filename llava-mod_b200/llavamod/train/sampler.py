"""Length-grouped sampling (`--group_by_modality_length`): samples of similar length land in the same global batch so the padded
micro-batches the CUDA path sees waste little.  Reference: llavamod/train/align_trainer.py:68-163 (same code in dpo_trainer.py and
llava_trainer.py).  All randomness goes through torch (`generator`), like the reference, so a seeded run draws the same order."""
from typing import List, Optional

import torch
from torch.utils.data import Sampler


def split_to_even_chunks(indices, lengths, num_chunks):
    """Deal a megabatch to `num_chunks` ranks.  Divisible case: longest-first greedy onto the currently lightest rank until a rank
    is full; otherwise plain striding (align_trainer.py:68-87)."""
    if len(indices) % num_chunks:
        return [indices[r::num_chunks] for r in range(num_chunks)]
    quota = len(indices) // num_chunks
    chunks = [[] for _ in range(num_chunks)]
    load = [0] * num_chunks
    for idx in indices:
        r = load.index(min(load))
        chunks[r].append(idx)
        load[r] += lengths[idx]
        if len(chunks[r]) == quota:
            load[r] = float("inf")
    return chunks


def get_length_grouped_indices(lengths, batch_size, world_size, generator=None, merge=True):
    """Shuffle, cut into megabatches of world_size*batch_size, sort each by length (descending), balance across ranks (:115-124)."""
    order = torch.randperm(len(lengths), generator=generator)
    mega = world_size * batch_size
    out = []
    for start in range(0, len(lengths), mega):
        block = sorted(order[start: start + mega].tolist(), key=lambda i: lengths[i], reverse=True)
        for chunk in split_to_even_chunks(block, lengths, world_size):
            out.extend(chunk)
    return out


def get_modality_length_grouped_indices(lengths, batch_size, world_size, generator=None):
    """lengths > 0: multimodal samples, < 0: text-only.  Megabatches are single-modality; their order is shuffled; the two ragged
    tails are merged into one final (sorted) megabatch (:90-112).  As in the reference the per-modality shuffles use the global RNG
    (generator=None) and only the megabatch permutation uses `generator`."""
    assert all(l != 0 for l in lengths), "Should not have zero length."
    if all(l > 0 for l in lengths) or all(l < 0 for l in lengths):
        return get_length_grouped_indices(lengths, batch_size, world_size, generator=generator)
    mm = [(i, l) for i, l in enumerate(lengths) if l > 0]
    tx = [(i, -l) for i, l in enumerate(lengths) if l < 0]
    mega = world_size * batch_size

    def blocks(pairs):
        idx, lens = zip(*pairs)
        order = [idx[j] for j in get_length_grouped_indices(lens, batch_size, world_size, generator=None)]
        return [order[s: s + mega] for s in range(0, len(order), mega)]

    mm_blocks, tx_blocks = blocks(mm), blocks(tx)
    tail = mm_blocks[-1] + tx_blocks[-1]
    body = mm_blocks[:-1] + tx_blocks[:-1]
    body = [body[j] for j in torch.randperm(len(body), generator=generator)]
    if tail:
        body.append(sorted(tail))
    return [i for block in body for i in block]


class LengthGroupedSampler(Sampler):
    """Yields one pass over the dataset in length-grouped order (align_trainer.py:127-163)."""

    def __init__(self, batch_size: int, world_size: int, lengths: Optional[List[int]] = None, generator=None, group_by_modality: bool = False):
        if lengths is None:
            raise ValueError("Lengths must be provided.")
        self.batch_size, self.world_size, self.lengths = batch_size, world_size, lengths
        self.generator, self.group_by_modality = generator, group_by_modality

    def __len__(self):
        return len(self.lengths)

    def __iter__(self):
        fn = get_modality_length_grouped_indices if self.group_by_modality else get_length_grouped_indices
        return iter(fn(self.lengths, self.batch_size, self.world_size, generator=self.generator))


class RankShard(Sampler):
    """This rank's share of a GLOBAL sample order: the order is cut into per-device batches and rank r keeps batches r, r+W, r+2W...
    (what accelerate's BatchSamplerShard does to the reference's dataloader); a ragged final round is dropped so every rank runs the
    same number of micro-batches (the gradient all-reduce needs that)."""

    def __init__(self, sampler, batch_size, rank, world_size):
        self.sampler, self.batch_size, self.rank, self.world_size = sampler, batch_size, rank, world_size

    def _rounds(self):
        return len(self.sampler) // (self.batch_size * self.world_size)

    def __len__(self):
        return self._rounds() * self.batch_size

    def __iter__(self):
        order = list(self.sampler)
        b, w = self.batch_size, self.world_size
        for rnd in range(self._rounds()):
            start = (rnd * w + self.rank) * b
            yield from order[start: start + b]


class EpochSeededRandomSampler(torch.utils.data.Sampler):
    """Single-process counterpart of DistributedSampler(shuffle=True, seed): the permutation of epoch e is a function of (seed, e)
    alone, so a run resumed from checkpoint-N walks the same batches the uninterrupted run would (HF Trainer reaches the same goal by
    re-seeding and skipping)."""

    def __init__(self, data_source, seed=42):
        self.n, self.seed, self.epoch = len(data_source), int(seed), 0

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def __iter__(self):
        g = torch.Generator()
        g.manual_seed(self.seed + self.epoch)
        return iter(torch.randperm(self.n, generator=g).tolist())

    def __len__(self):
        return self.n
