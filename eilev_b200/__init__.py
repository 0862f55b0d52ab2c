"""eilev_b200 — Blackwell (sm_100a) implementation of the EILEV / VideoBLIP hot path.

Public surface mirrors ``eilev.model.v2`` / ``eilev.model.v1`` / ``eilev.model.utils`` /
``eilev.data.utils`` of the reference (yukw777/EILEV):

    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration, VideoBlipVisionModel
    from eilev_b200.model.v1 import VideoBlipForConditionalGeneration   # HF 4.33.1 Blip2 signatures
    from eilev_b200.model.utils import process, process_on_device
    from eilev_b200.data.utils import DataCollatorForInterleavedVideoSeq2Seq, \
        generate_input_ids_and_labels_from_interleaved
"""
__version__ = "0.1.0"
