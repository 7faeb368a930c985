"""The language-context producer in front of the fusion path on the path's own kernels."""
from .minilm import XfLinear, bert_encoder_forward, SBertTokensXf  # noqa: F401
