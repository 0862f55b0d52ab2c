"""``eilev.model.v2`` (eilev/model/v2.py:24-501) on the B200 kernels."""
from eilev_b200.model.v2 import VideoBlipForConditionalGeneration, VideoBlipVisionModel  # noqa: F401
