"""Import-path shim: ``from dstagnn_drought_b200.model.DSTAGNN_my import make_model`` mirrors the
reference's ``from model.DSTAGNN_my import make_model`` (train_DSTAGNN_my.py:14)."""
from ..modules import (DSTAGNN_block, DSTAGNN_submodule, Embedding, GTU, MultiHeadAttention,  # noqa: F401
                       SMultiHeadAttention, cheb_conv_withSAt, make_model)
