"""ebc — host side of the B200-native EB-CADRL hot path (C ABI: include/ebcadrl.h)."""
from .abi import EbcError, EVENT_NAMES  # noqa: F401
from .config import SimConfig, read_ini  # noqa: F401
from .actions import build_action_space  # noqa: F401
