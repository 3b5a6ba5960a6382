"""audio-training_b200: B200-native (sm_100a) audio front-end behind the reference's feature-function API.

Import as `audio_training_b200` (the importable alias of this directory).  Sub-modules mirror the reference's own
module names: custommel, tfdataset, predict_utils, tfpcen, badwinner2.
"""
from . import _lib  # noqa: F401
from ._lib import CacfeError, build  # noqa: F401
from ._runtime import FrontendConfig, HostPipe, Plan, clear_plans, get_plan, pcen_params  # noqa: F401
from . import audiodataset, badwinner2, custommel, distributed, predict_utils, tfdataset, tfpcen  # noqa: F401
from .badwinner2 import MagTransform  # noqa: F401
from .custommel import hz_to_mel, mel_f, mel_frequencies, mel_spec  # noqa: F401
from .predict_utils import get_spect, load_samples, normalize_data  # noqa: F401
from .tfdataset import normalize, normalize_std, power_to_db, raw_to_mel  # noqa: F401
from .tfpcen import PCEN, ExponentialMovingAverage, normalize_minmax  # noqa: F401

__version__ = "0.1.0"
