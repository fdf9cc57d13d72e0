"""`GCBC` — drop-in for `hulc.models.gcbc.GCBC` (reference hulc/models/gcbc.py:11-181): goal-conditioned behaviour cloning
ablation.  The decoder receives an empty plan (`plan_features = 0`, gcbc.py:16-48), there is no prior and no KL; the
posterior transformer still runs because its `seq_feat` feeds the CLIP auxiliary loss (gcbc.py:50-181)."""
from .hulc import Hulc


class GCBC(Hulc):
    MODEL = "gcbc"
