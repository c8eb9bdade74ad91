"""biospectra_b200 -- B200-native (sm_100a) k-mer index construction + read classification,
a drop-in for that one hot path of iychoi/biospectra.  See DESIGN.md / INTEGRATION.md."""
from .config import Configuration  # noqa: F401
from .engine import Classifier, Indexer  # noqa: F401
from ._lib import BspError, LABEL_NAMES  # noqa: F401
