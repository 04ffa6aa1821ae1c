r"""``RNN`` / ``RNNWithInit`` parameter containers (reference ``articulate/utils/torch/rnn.py:92-133, 174-219``).

``Net.forward_online`` never calls ``RNN.forward``; it drives the sub-modules ``linear1 / rnn / linear2 / init_net``
directly (net/sig_mp.py:126-129, 182).  Here the modules only carry the parameters under the reference's
``state_dict`` names, so checkpoints load unchanged; the arithmetic happens in the CUDA library after
``Net`` packs these tensors (``rc_net_set_tensor``).  ``forward`` over lists of sequences (the training-time API)
is kept for compatibility on top of torch's own LSTM and is not part of the hot path.
"""
import torch
from torch.nn.functional import relu
from torch.nn.utils.rnn import pad_sequence, pack_padded_sequence, pad_packed_sequence

__all__ = ['RNN', 'RNNWithInit']


class RNN(torch.nn.Module):
    def __init__(self, input_size: int, output_size: int, hidden_size: int, num_rnn_layer: int,
                 rnn_type='lstm', bidirectional=False, dropout=0., load_weight_file: str = None):
        super().__init__()
        assert rnn_type == 'lstm' and not bidirectional, 'the B200 hot path implements unidirectional LSTM stacks'
        self.rnn = torch.nn.LSTM(hidden_size, hidden_size, num_rnn_layer, bidirectional=False, dropout=dropout)
        self.linear1 = torch.nn.Linear(input_size, hidden_size)
        self.linear2 = torch.nn.Linear(hidden_size, output_size)
        self.dropout = torch.nn.Dropout(dropout) if dropout > 0 else torch.nn.Identity()
        if load_weight_file:
            import os
            if os.path.exists(load_weight_file):
                self.load_state_dict(torch.load(load_weight_file, map_location=torch.device('cpu')))
                self.eval()

    def forward(self, x, init=None):
        r"""List of [num_frames, input_size] -> list of [num_frames, output_size]. rnn.py:120-133."""
        length = [_.shape[0] for _ in x]
        x = self.dropout(relu(self.linear1(pad_sequence(x))))
        x = self.rnn(pack_padded_sequence(x, length, enforce_sorted=False), init)[0]
        x = self.linear2(pad_packed_sequence(x)[0])
        return [x[:l, i].clone() for i, l in enumerate(length)]


class RNNWithInit(RNN):
    def __init__(self, input_size: int, output_size: int, hidden_size: int, num_rnn_layer: int,
                 rnn_type='lstm', bidirectional=False, dropout=0., load_weight_file: str = None):
        super().__init__(input_size, output_size, hidden_size, num_rnn_layer, rnn_type, bidirectional, dropout)
        self.init_net = torch.nn.Sequential(
            torch.nn.Linear(output_size, hidden_size),
            torch.nn.ReLU(),
            torch.nn.Linear(hidden_size, hidden_size * num_rnn_layer),
            torch.nn.ReLU(),
            torch.nn.Linear(hidden_size * num_rnn_layer, 2 * num_rnn_layer * hidden_size),
        )
        if load_weight_file:
            import os
            if os.path.exists(load_weight_file):
                self.load_state_dict(torch.load(load_weight_file, map_location=torch.device('cpu')))
                self.eval()

    def forward(self, x, _=None):
        r"""rnn.py:207-219."""
        x, x_init = list(zip(*x))
        nd, nh = self.rnn.num_layers, self.rnn.hidden_size
        h, c = self.init_net(torch.stack(x_init)).view(-1, 2, nd, nh).permute(1, 2, 0, 3)
        return super(RNNWithInit, self).forward(x, (h.contiguous(), c.contiguous()))
