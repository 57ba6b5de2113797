"""``MLP`` with the reference's constructor and ``state_dict`` layout, evaluated by the
library's Linear(+ReLU) kernel.  reference: models/mlp.py:4-28"""
from torch import nn

from .. import ops


class MLP(nn.Module):
    """Linear -> ReLU per entry of ``fc_dims``; a layer of width 1 gets no ReLU
    (reference: models/mlp.py:14-21).  Parameters live in ``fc_layers`` at the same
    Sequential slots as the reference, so its checkpoints load unchanged."""

    def __init__(self, input_dim, fc_dims, dropout_p=0.4, use_batchnorm=False):
        super().__init__()
        assert isinstance(fc_dims, (list, tuple)), \
            'fc_dims must be either a list or a tuple, but got {}'.format(type(fc_dims))
        if use_batchnorm:
            raise NotImplementedError('use_batchnorm=True is not built into the CUDA path '
                                      '(every shipped config has use_batchnorm: False)')
        self.dropout_p = dropout_p
        layers = []
        for dim in fc_dims:
            layers.append(nn.Linear(input_dim, dim))
            if dim != 1:
                layers.append(nn.ReLU(inplace=True))
            if dropout_p != 0 and dim != 1:
                layers.append(nn.Dropout(p=dropout_p))
            input_dim = dim
        self.fc_layers = nn.Sequential(*layers)

    def linears(self):
        return [m for m in self.fc_layers if isinstance(m, nn.Linear)]

    def forward(self, input):
        if self.dropout_p != 0 and self.training:
            raise NotImplementedError('dropout in training mode is not built into the CUDA path')
        h = input
        for lin in self.linears():
            h = ops.linear(h, lin.weight, lin.bias, relu=lin.out_features != 1)
        return h
