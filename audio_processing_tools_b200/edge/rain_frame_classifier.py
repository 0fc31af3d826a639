"""Frame classes of the rain detector (reference: edge/rain_frame_classifier.py:18-23)."""
from enum import IntEnum


class FrameClass(IntEnum):
    NOISE = 0
    UNCERTAIN = 1
    RAIN = 2
