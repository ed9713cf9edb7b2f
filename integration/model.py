"""Drop this file next to main_disentangled.py in place of the reference's model.py (or put the
repository root on PYTHONPATH and keep this name): `from model import Disentangle`
(main_disentangled.py:14) then resolves to the B200 implementation.  See INTEGRATION.md."""
from disenlink_b200.model import (Dec, Dec2, Disentangle, Disentangle_layer,  # noqa: F401
                                  Disentangle_out_layer, Factor, Factor2, LinkScorer)
