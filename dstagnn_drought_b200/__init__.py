"""dstagnn_drought_b200 -- B200-native (sm_100a) DSTAGNN block forward/backward behind the
reference's module API.  See DESIGN.md."""
from .contraction import contract, contract_raw, get_precision, set_precision  # noqa: F401
from .modules import (DSTAGNN_block, DSTAGNN_submodule, Embedding, GTU, MultiHeadAttention,  # noqa: F401
                      SMultiHeadAttention, cheb_conv_withSAt, make_model)

__all__ = ["make_model", "DSTAGNN_submodule", "DSTAGNN_block", "contract", "contract_raw",
           "set_precision", "get_precision"]
