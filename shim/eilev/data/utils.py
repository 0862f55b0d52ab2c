"""``eilev.data.utils`` (eilev/data/utils.py): the interleaved tokeniser and collators of the hot path
(bit-exact restatements, eilev_b200/data/utils.py) under the reference's names.  Names this repo does not
implement (``NarratedActionClipSampler``: a pytorchvideo clip sampler, outside the hot path) fall through
to the reference's own module when an EILEV checkout sits behind this directory on ``sys.path``."""
import importlib.util as _ilu
import os as _os

from eilev_b200.data.utils import (  # noqa: F401
    DataCollatorForInterleavedVideoSeq2Seq,
    DataCollatorForVideoSeq2Seq,
    clean_narration_text,
    generate_chunks,
    generate_input_ids_and_labels,
    generate_input_ids_and_labels_from_interleaved,
    parse_timestamp,
)
from eilev_b200.data.utils import _RE_CAMERA_WEARER as C_REGEX  # noqa: F401
from eilev_b200.data.utils import _RE_EOS as EOS_REGEX  # noqa: F401
from eilev_b200.data.utils import _RE_UNSURE as UNSURE_MIDDLE_REGEX  # noqa: F401
from eilev_b200.data.utils import _RE_UNSURE_TAIL as UNSURE_END_REGEX  # noqa: F401

_reference = None


def _reference_module():
    global _reference
    if _reference is None:
        import eilev.data as pkg

        here = _os.path.dirname(_os.path.abspath(__file__))
        for d in pkg.__path__:
            cand = _os.path.join(d, "utils.py")
            if _os.path.abspath(d) != here and _os.path.isfile(cand):
                spec = _ilu.spec_from_file_location("eilev.data._reference_utils", cand)
                mod = _ilu.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _reference = mod
                break
        else:
            raise ImportError("no reference eilev/data/utils.py behind the shim on sys.path")
    return _reference


def __getattr__(name):  # PEP 562: only reached for names not defined above
    try:
        return getattr(_reference_module(), name)
    except ImportError as exc:
        raise AttributeError(f"eilev.data.utils.{name} is outside the B200 hot path and no EILEV checkout "
                             f"is on sys.path behind the shim ({exc})") from None
