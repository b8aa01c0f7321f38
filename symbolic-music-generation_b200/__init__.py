"""symbolic-music-generation_b200 — B200-native Transformer-XL hot path of StefanHeng/Symbolic-Music-Generation.

Public surface mirrors reference `musicnlp/models/transformer_xl.py`:
`MyTransfoXLConfig`, `MyTransfoXLLMHeadModel`, `TransfoXLLMHeadModelOutput`.
The directory name carries a hyphen (it is the graft's package name); import it with
`importlib.import_module("symbolic-music-generation_b200")` or through the root alias module `smg_b200`.
"""
from .model import MyTransfoXLConfig, MyTransfoXLLMHeadModel, TransfoXLLMHeadModelOutput, TxlMems, PT_LOSS_PAD
from ._lib import TxlError, LIB_PATH

__all__ = ['MyTransfoXLConfig', 'MyTransfoXLLMHeadModel', 'TransfoXLLMHeadModelOutput', 'TxlMems', 'TxlError', 'PT_LOSS_PAD', 'LIB_PATH']
