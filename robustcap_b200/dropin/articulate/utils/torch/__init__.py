r"""``from articulate.utils.torch import *`` (articulate/utils/torch/__init__.py): RNN blocks + the dataset helper
evaluate.py needs (RNNDataset.collate_fn, rnn.py:15-60).  Training utilities are out of scope."""
import torch
from robustcap_b200.rnn import RNN, RNNWithInit  # noqa: F401


class RNNDataset(torch.utils.data.Dataset):
    r"""List-of-sequences dataset (rnn.py:15-60): items are (data[T, in], label[T, out])."""

    def __init__(self, data: list, label: list, split_size=-1, augment_fn=None, device=None):
        assert len(data) == len(label) and len(data) != 0
        if split_size > 0:
            self.data, self.label = [], []
            for td, tl in zip(data, label):
                self.data.extend(td.split(split_size))
                self.label.extend(tl.split(split_size))
        else:
            self.data, self.label = data, label
        self.augment_fn = augment_fn
        self.device = data[0].device if device is None else device       # rnn.py:60: default = where data[0] lives

    def __getitem__(self, i):
        item = self.data[i]
        if self.augment_fn is not None:
            item = self.augment_fn(item)
        return item.to(self.device), self.label[i].to(self.device)

    def __len__(self):
        return len(self.data)

    @staticmethod
    def collate_fn(x):
        return list(zip(*x))


__all__ = ['RNN', 'RNNWithInit', 'RNNDataset']
