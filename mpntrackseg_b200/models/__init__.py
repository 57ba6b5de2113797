from .mpn import (MOTMPNet, MetaLayer, EdgeModel, TimeAwareNodeModel, MLPGraphIndependent)  # noqa: F401
from .mlp import MLP  # noqa: F401
