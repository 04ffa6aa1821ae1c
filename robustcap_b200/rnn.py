r"""``RNN`` / ``RNNWithInit`` parameter containers (reference ``articulate/utils/torch/rnn.py:92-133, 174-219``).

``Net.forward_online`` never calls ``RNN.forward``; it drives the sub-modules ``linear1 / rnn / linear2 / init_net``
directly (net/sig_mp.py:126-129, 182).  Here the modules only carry the parameters under the reference's
``state_dict`` names, so checkpoints load unchanged; the arithmetic happens in the CUDA library after
``Net`` packs these tensors (``rc_net_set_tensor``).  ``forward`` over lists of padded sequences is the reference's
TRAINING-time API (net/sig_mp.py:301-857, out of scope): it raises and points at the reference.
"""
import torch

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
        raise RuntimeError('robustcap_b200.rnn.RNN only carries parameters for Net (inference hot path); the padded-sequence '
                           'training forward is articulate/utils/torch/rnn.py:120-133 of the reference and is out of scope here')


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

