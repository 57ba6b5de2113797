from .mpn import (MOTMPNet, MetaLayer, EdgeModel, TimeAwareNodeModel, TimeAwareAttentionModel,  # noqa: F401
                  MLPGraphIndependent, MaskModel, BatchOutput)
from .mlp import MLP  # noqa: F401
from .cnn import CNN, MaskRCNNPredictor  # noqa: F401
