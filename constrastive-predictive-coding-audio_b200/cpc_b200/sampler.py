"""Batch composition: which dataset items share a batch, i.e. who the in-batch negatives are.

``FileBatchSampler`` reproduces the index stream of the reference (audio_dataset.py:202-263) bit for bit:
same use of Python's ``random`` (global Mersenne Twister state when ``seed`` is None), same chunking and
drop-last rules, same ``__len__`` quirk.  ``SyntheticAudioDataset`` is the in-memory stand-in used by the
benchmark and tests (disk decoding is out of scope).
"""
import itertools
import math
import random

import numpy as np
import torch
import torch.utils.data


def _fixed_chunks(seq, size, drop_last):
    for start in range(0, len(seq), size):
        if drop_last and start + size > len(seq):
            return
        yield seq[start:start + size]


class FileBatchSampler(torch.utils.data.Sampler):
    def __init__(self, index_count_per_file, batch_size, file_batch_size=1, drop_last=True, seed=None, verbose=False):
        self.index_count_per_file = index_count_per_file
        self.indices_in_file = []
        first = 0
        for count in index_count_per_file:
            self.indices_in_file.append(list(range(first, first + count)))
            first += count
        self.batch_size = batch_size
        self.file_batch_size = file_batch_size
        self.drop_last = drop_last
        self.seed = seed
        rounding = math.floor if drop_last else math.ceil
        self.batches_per_file = [rounding(n / file_batch_size) for n in index_count_per_file]
        if verbose:
            print("minimum batches per file:", min(self.batches_per_file),
                  "maximum batches per file:", max(self.batches_per_file))

    def __iter__(self):
        if self.file_batch_size == 1:
            order = list(range(len(self)))
            if self.seed is not None:
                random.seed(self.seed)
            random.shuffle(order)
            return iter(_fixed_chunks(order, self.batch_size, self.drop_last))
        for i, per_file in enumerate(self.indices_in_file):
            if self.seed is not None:
                random.seed(self.seed + i)
            random.shuffle(per_file)                     # in place: state carries over between epochs
        groups = []
        for per_file in self.indices_in_file:
            groups.extend(_fixed_chunks(per_file, self.file_batch_size, self.drop_last))
        if self.seed is not None:
            random.seed(self.seed)
        random.shuffle(groups)
        groups_per_batch = self.batch_size // self.file_batch_size
        if groups_per_batch > 1:
            return iter(list(itertools.chain(*chunk))
                        for chunk in _fixed_chunks(groups, groups_per_batch, self.drop_last))
        return iter(groups)

    def __len__(self):
        return int(np.sum(self.batches_per_file))


class SyntheticAudioDataset(torch.utils.data.Dataset):
    """``n_items`` items of ``item_length`` samples of seeded white noise, organised as ``files`` pseudo files.
    Item i is ``amplitude * randn`` drawn from a generator seeded with ``seed + i`` (reproducible, order
    independent)."""

    def __init__(self, item_length, n_items, files=1, amplitude=0.1, seed=1234):
        self.item_length = int(item_length)
        self.n_items = int(n_items)
        self.files = int(files)
        self.amplitude = amplitude
        self.seed = seed

    def __len__(self):
        return self.n_items

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed + int(idx))
        return self.amplitude * torch.randn(self.item_length, generator=g)

    def get_example_count_per_file(self):
        base, extra = divmod(self.n_items, self.files)
        return [base + (1 if i < extra else 0) for i in range(self.files)]
