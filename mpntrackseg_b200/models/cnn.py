"""Convolutional sub-modules of the mask branch, with the reference's constructor signatures and
``state_dict`` layout (``layers.{2i}.weight/bias``).  The dense convolutions run on cuDNN through
``torch.nn`` -- SURVEY.md section 2 (#3) scopes hand-written kernels out for them.
reference: src/mot_neural_solver/models/cnn.py (CNN :4-44, MaskRCNNPredictor :47-84)"""
from torch import nn


def _check_lists(dims, kernel_sizes, strides, paddings):
    for name, v in (('dims', dims), ('kernel_sizes', kernel_sizes), ('strides', strides), ('paddings', paddings)):
        assert isinstance(v, (list, tuple)), '{} must be either a list or a tuple, but got {}'.format(name, type(v))
    assert len(dims) == len(kernel_sizes) == len(strides) == len(paddings), \
        'Number of elements mismatch between dims, kernel_sizes and strides'


class CNN(nn.Module):
    """Conv2d + ReLU per entry of ``dims`` (a ReLU follows EVERY conv, reference: models/cnn.py:33-34)."""

    def __init__(self, input_dim, dims, kernel_sizes, strides, paddings, dropout_p=0.4, use_batchnorm=False):
        super().__init__()
        _check_lists(dims, kernel_sizes, strides, paddings)
        if use_batchnorm:
            raise NotImplementedError('use_batchnorm=True is not supported (every shipped config has it off)')
        mods = []
        for d, k, s, p in zip(dims, kernel_sizes, strides, paddings):
            mods.append(nn.Conv2d(input_dim, d, kernel_size=k, stride=s, padding=p))
            mods.append(nn.ReLU(inplace=True))
            if dropout_p != 0 and d != 1:
                mods.append(nn.Dropout2d(p=dropout_p))
            input_dim = d
        self.layers = nn.Sequential(*mods)

    def forward(self, input):
        return self.layers(input)


class MaskRCNNPredictor(nn.Module):
    """(Transposed) convolutions with a ReLU between them, none after the last.
    reference: models/cnn.py:47-84"""

    def __init__(self, input_dim, dims, kernel_sizes, strides, paddings, transposed):
        super().__init__()
        _check_lists(dims, kernel_sizes, strides, paddings)
        mods = []
        for i, (d, k, s, p) in enumerate(zip(dims, kernel_sizes, strides, paddings)):
            conv = nn.ConvTranspose2d if transposed[i] else nn.Conv2d
            mods.append(conv(input_dim, d, kernel_size=k, stride=s, padding=p))
            if i < len(dims) - 1:
                mods.append(nn.ReLU(inplace=True))
            input_dim = d
        self.layers = nn.Sequential(*mods)

    def forward(self, input):
        return self.layers(input)
