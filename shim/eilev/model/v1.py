"""``eilev.model.v1`` (eilev/model/v1.py:14-119) on the B200 kernels."""
from eilev_b200.model.v1 import VideoBlipForConditionalGeneration  # noqa: F401
from eilev_b200.model.v2 import VideoBlipVisionModel  # noqa: F401
