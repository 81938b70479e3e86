"""ustrun -- host side of the B200-native UST-RUN SSL train step."""
from . import _lib  # noqa: F401  (fails loudly when libustrun_sm100.so is missing)
from .engine import get_precision, set_force_simt, set_precision  # noqa: F401
