"""Mirror of igneous/types.py:4-12 (enum values are serialised into task JSON)."""
import enum
from typing import Tuple

ShapeType = Tuple[int, int, int]


class DownsampleMethods(enum.IntEnum):
  AVERAGE_POOLING = 1
  MODE_POOLING = 2
  MIN_POOLING = 3
  MAX_POOLING = 4
  STRIDING = 5
  AUTO = 6
