"""spml_b200: the pixel-to-segment contrastive hot path of SPML (twke18/SPML) as
hand-written sm_100a CUDA behind a C ABI (include/spml_b200.h), with the
reference's own Python operator API on top.

    from spml_b200 import segsort_common, segsort_loss, segsort_eval, model_utils
    from spml_b200.predictions import segsort            # nn.Module factory
    spml_b200.install()                                  # rebind inside `spml.*`

The modules mirror the reference's: segsort_common <- spml/utils/segsort/common.py,
segsort_loss <- .../loss.py, segsort_eval <- .../eval.py, general_common <-
spml/utils/general/common.py, model_utils <- spml/models/utils.py, predictions <-
spml/models/predictions/segsort{,_softmax}.py.  CUDA only; no CPU fallback.
"""

from . import _lib
from . import general_common, model_utils, ops, predictions, segsort_common, segsort_eval
from . import segsort_loss, synth
from .install import install, uninstall
from .head import ContrastiveHead

__all__ = ['general_common', 'model_utils', 'ops', 'predictions', 'segsort_common',
           'segsort_eval', 'segsort_loss', 'synth', 'install', 'uninstall', 'ContrastiveHead']
__version__ = '0.1.0'
