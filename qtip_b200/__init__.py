"""qtip_b200: B200-native tandem-read simulator for Qtip (drop-in for the reference's
simplesim/input_model/fasta path).  See DESIGN.md and include/qtipsim.h."""
from ._abi import (CAT_B, CAT_C, CAT_D, CAT_U, FUNC_CONST, FUNC_LINEAR, FUNC_SQRT, NOUT, OUT_NAMES, RNG_PHILOX,
                   RNG_REPLAY)
from .templates import TemplatePaired, TemplateUnpaired

__all__ = ["TemplateUnpaired", "TemplatePaired", "CAT_U", "CAT_B", "CAT_C", "CAT_D", "FUNC_LINEAR", "FUNC_SQRT",
           "FUNC_CONST", "RNG_REPLAY", "RNG_PHILOX", "NOUT", "OUT_NAMES"]
