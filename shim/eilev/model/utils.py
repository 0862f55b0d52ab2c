"""``eilev.model.utils`` (eilev/model/utils.py:5-26)."""
from eilev_b200.model.utils import process  # noqa: F401
